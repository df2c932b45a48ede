"""CPU: the zarr (format 2) directory store of the sensitivity matrix (simpeg_b200/zarr_store.py;
the layout the reference's dask path keeps, simpeg/dask/potential_fields/base.py:46-66)."""
import gzip
import json
import os
import zlib

import numpy as np
import pytest

from simpeg_b200.zarr_store import ZarrMatrixReader, ZarrMatrixWriter, is_zarr_store, wants_zarr


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,chunk_rows", [((37, 11), 8), ((5, 3), None), ((16, 4), 16), ((1, 1), 1), ((0, 4), None)])
def test_write_read_roundtrip_and_header(tmp_path, shape, chunk_rows, dtype):
    rng = np.random.default_rng(0)
    G = rng.normal(size=shape).astype(dtype)
    path = tmp_path / "sensitivity.zarr"
    assert wants_zarr(path) and not wants_zarr(tmp_path / "sensitivity")
    w = ZarrMatrixWriter(path, shape, dtype, chunk_rows=chunk_rows)
    # two "ranks" write interleaved row sets (rows of one chunk come from both)
    rows = np.arange(shape[0])
    for part in (rows[rows % 3 != 1], rows[rows % 3 == 1]):
        ZarrMatrixWriter(path, shape, dtype, chunk_rows=chunk_rows, create=False).write_rows(part, G[part])
    assert is_zarr_store(path)
    meta = json.load(open(path / ".zarray"))
    assert meta["zarr_format"] == 2 and meta["shape"] == list(shape) and meta["compressor"] is None
    assert meta["order"] == "C" and np.dtype(meta["dtype"]) == np.dtype(dtype)
    assert meta["chunks"] == [w.chunk_rows, max(shape[1], 1)]
    # every chunk file is a full C-order chunk, as the zarr specification stores edge chunks
    for k in range(w.n_chunks):
        assert os.path.getsize(path / f"{k}.0") == w.chunk_rows * shape[1] * np.dtype(dtype).itemsize
    r = ZarrMatrixReader(path)
    assert r.shape == shape and r.dtype == np.dtype(dtype)
    assert np.array_equal(r.read_rows(rows), G)
    pick = rows[::-1][::2]
    assert np.array_equal(r[pick], G[pick])
    assert np.array_equal(r[1:4], G[1:4])


@pytest.mark.parametrize("codec", [None, "zlib", "gzip"])
def test_reads_column_chunked_and_compressed_stores(tmp_path, codec):
    """A store as zarr / dask write it: 2-D chunk grid, optional zlib / gzip codec, missing chunks
    filled with fill_value, '/' as dimension separator."""
    rng = np.random.default_rng(1)
    G = rng.normal(size=(10, 7))
    cr, cc = 4, 3
    path = tmp_path / "s.zarr"
    os.makedirs(path)
    comp = None if codec is None else {"id": codec, "level": 1}
    json.dump({"zarr_format": 2, "shape": [10, 7], "chunks": [cr, cc], "dtype": "<f8", "compressor": comp,
               "fill_value": 0.0, "order": "C", "filters": None, "dimension_separator": "/"},
              open(path / ".zarray", "w"))
    expect = G.copy()
    for i in range(3):
        os.makedirs(path / str(i))
        for j in range(3):
            if (i, j) == (1, 1):          # a chunk that was never written
                expect[4:8, 3:6] = 0.0
                continue
            block = np.zeros((cr, cc))
            sub = G[i * cr:(i + 1) * cr, j * cc:(j + 1) * cc]
            block[: sub.shape[0], : sub.shape[1]] = sub
            raw = block.tobytes()
            raw = zlib.compress(raw) if codec == "zlib" else gzip.compress(raw) if codec == "gzip" else raw
            open(path / str(i) / str(j), "wb").write(raw)
    r = ZarrMatrixReader(path)
    assert np.array_equal(r.read_rows(np.arange(10)), expect)
    assert np.array_equal(r.read_rows(np.array([9, 0, 5])), expect[[9, 0, 5]])


def test_unsupported_stores_fail_loudly(tmp_path):
    path = tmp_path / "b.zarr"
    os.makedirs(path)
    base = {"zarr_format": 2, "shape": [4, 4], "chunks": [4, 4], "dtype": "<f8", "fill_value": 0.0,
            "order": "C", "filters": None}
    json.dump(dict(base, compressor={"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 1}),
              open(path / ".zarray", "w"))
    try:
        import numcodecs  # noqa: F401
    except ImportError:
        with pytest.raises(NotImplementedError, match="numcodecs"):
            ZarrMatrixReader(path)
    json.dump(dict(base, compressor=None, zarr_format=3), open(path / ".zarray", "w"))
    with pytest.raises(NotImplementedError, match="format 2"):
        ZarrMatrixReader(path)
    json.dump(dict(base, compressor=None, order="F"), open(path / ".zarray", "w"))
    with pytest.raises(NotImplementedError, match="C-order"):
        ZarrMatrixReader(path)
    json.dump(dict(base, compressor=None), open(path / ".zarray", "w"))
    with pytest.raises(IndexError):
        ZarrMatrixReader(path).read_rows(np.array([4]))
    with pytest.raises(ValueError, match="block must be"):
        ZarrMatrixWriter(tmp_path / "c.zarr", (4, 4), np.float64).write_rows(np.arange(2), np.zeros((3, 4)))
