"""
Minimal zarr (format 2) directory-store access for the sensitivity matrix.

The reference's dask path keeps ``store_sensitivities="disk"`` matrices as
``<sensitivity_path>/sensitivity.zarr`` (``simpeg/dask/potential_fields/base.py:46-66``:
``dask.array.to_zarr`` / ``from_zarr``).  Neither ``zarr`` nor ``dask`` is a dependency of this
package, and the store is simple enough to read and write directly: a directory with a ``.zarray``
JSON header and one file per chunk, named by the chunk's grid index (``"3.0"``), holding the
chunk's C-order bytes (edge chunks are stored full-size).

* :class:`ZarrMatrixWriter` writes an UNCOMPRESSED 2-D array in row chunks that span all columns
  (the layout ``chunk_format="row"`` of the reference produces).  Chunk files are pre-sized, so
  several ranks can write disjoint rows of one chunk.  ``zarr.open`` / ``dask.array.from_zarr``
  read the result as is.
* :class:`ZarrMatrixReader` reads 2-D C-order arrays with no compressor or with the ``zlib`` /
  ``gzip`` codecs (python's standard library decodes both); a Blosc-compressed store -- zarr's
  default compressor -- needs ``numcodecs`` and raises ``NotImplementedError`` with that message
  when it is not importable.
"""
import gzip
import json
import os
import zlib

import numpy as np

ZARRAY = ".zarray"


def is_zarr_store(path):
    """True for a directory that holds a zarr (format 2) array header."""
    return os.path.isdir(path) and os.path.isfile(os.path.join(path, ZARRAY))


def wants_zarr(path):
    """``sensitivity_path`` names a zarr store (``*.zarr``) rather than the raw memmap file."""
    return str(path).rstrip("/\\").endswith(".zarr")


class ZarrMatrixWriter:
    """Uncompressed (rows, cols) zarr array in chunks of ``chunk_rows`` full rows."""

    def __init__(self, path, shape, dtype, chunk_rows=None, create=True, target_chunk_bytes=128 << 20):
        self.path = str(path)
        self.shape = (int(shape[0]), int(shape[1]))
        self.dtype = np.dtype(dtype)
        row_bytes = max(1, self.shape[1] * self.dtype.itemsize)
        if chunk_rows is None:
            # the reference's default chunk budget is 128 MiB (max_chunk_size, dask path)
            chunk_rows = max(1, min(max(self.shape[0], 1), target_chunk_bytes // row_bytes))
        self.chunk_rows = int(chunk_rows)
        self.n_chunks = -(-self.shape[0] // self.chunk_rows) if self.shape[0] else 0
        if create:
            self._create()

    def _chunk_file(self, k):
        return os.path.join(self.path, f"{k}.0")

    def _create(self):
        os.makedirs(self.path, exist_ok=True)
        for name in os.listdir(self.path):           # a stale store of another shape must not survive
            full = os.path.join(self.path, name)
            if os.path.isfile(full):
                os.remove(full)
        header = {
            "zarr_format": 2,
            "shape": list(self.shape),
            "chunks": [self.chunk_rows, max(self.shape[1], 1)],
            "dtype": self.dtype.newbyteorder("<").str if self.dtype.itemsize > 1 else self.dtype.str,
            "compressor": None,
            "fill_value": 0.0 if self.dtype.kind == "f" else 0,
            "order": "C",
            "filters": None,
            "dimension_separator": ".",
        }
        with open(os.path.join(self.path, ZARRAY), "w") as fh:
            json.dump(header, fh, indent=4)
        chunk_bytes = self.chunk_rows * self.shape[1] * self.dtype.itemsize
        for k in range(self.n_chunks):
            with open(self._chunk_file(k), "wb") as fh:
                fh.truncate(chunk_bytes)

    def write_rows(self, rows, block):
        """Store ``block[i]`` as row ``rows[i]`` of the array (``rows`` ascending)."""
        rows = np.asarray(rows, dtype=np.int64)
        block = np.asarray(block)
        if block.shape != (rows.size, self.shape[1]):
            raise ValueError(f"block must be {(rows.size, self.shape[1])}, got {block.shape}")
        if rows.size == 0 or self.shape[1] == 0:
            return
        chunk_of = rows // self.chunk_rows
        starts = np.r_[0, np.nonzero(np.diff(chunk_of))[0] + 1, rows.size]
        for a, b in zip(starts[:-1], starts[1:]):
            mm = np.memmap(self._chunk_file(int(chunk_of[a])), dtype=self.dtype.newbyteorder("<"), mode="r+",
                           shape=(self.chunk_rows, self.shape[1]), order="C")
            mm[rows[a:b] % self.chunk_rows] = block[a:b]
            mm.flush()
            del mm


class ZarrMatrixReader:
    """Row access to a 2-D zarr (format 2) array stored in a directory."""

    def __init__(self, path):
        self.path = str(path)
        with open(os.path.join(self.path, ZARRAY)) as fh:
            meta = json.load(fh)
        if meta.get("zarr_format") != 2:
            raise NotImplementedError(f"zarr_format {meta.get('zarr_format')} is not supported (format 2 only)")
        if len(meta["shape"]) != 2:
            raise ValueError(f"expected a 2-D array, found shape {meta['shape']}")
        if meta.get("order", "C") != "C":
            raise NotImplementedError("only C-order zarr arrays are supported")
        if meta.get("filters"):
            raise NotImplementedError("zarr filters are not supported")
        self.shape = tuple(int(v) for v in meta["shape"])
        self.chunks = tuple(int(v) for v in meta["chunks"])
        self.dtype = np.dtype(meta["dtype"])
        self.fill_value = meta.get("fill_value") or 0
        self.sep = meta.get("dimension_separator", ".")
        comp = meta.get("compressor")
        self.codec = None if comp is None else comp.get("id")
        if self.codec not in (None, "zlib", "gzip"):
            try:
                import numcodecs  # noqa: F401
            except ImportError:
                raise NotImplementedError(
                    f"zarr store compressed with '{self.codec}': decoding needs the numcodecs package; "
                    "re-save the store with compressor=None (or zlib / gzip)") from None
            self._numcodecs = numcodecs.get_codec(comp)

    def _decode(self, raw):
        if self.codec is None:
            return raw
        if self.codec == "zlib":
            return zlib.decompress(raw)
        if self.codec == "gzip":
            return gzip.decompress(raw)
        return self._numcodecs.decode(raw)

    def _chunk(self, i, j):
        name = os.path.join(self.path, f"{i}{self.sep}{j}".replace("/", os.sep))
        if not os.path.isfile(name):
            return np.full(self.chunks, self.fill_value, dtype=self.dtype)
        with open(name, "rb") as fh:
            raw = self._decode(fh.read())
        return np.frombuffer(raw, dtype=self.dtype).reshape(self.chunks)

    def read_rows(self, rows):
        """The rows ``rows`` (ascending index array) as a native-endian (len(rows), cols) array."""
        rows = np.asarray(rows, dtype=np.int64)
        out = np.empty((rows.size, self.shape[1]), dtype=self.dtype.newbyteorder("="))
        if rows.size == 0 or self.shape[1] == 0:
            return out
        if rows.min() < 0 or rows.max() >= self.shape[0]:
            raise IndexError("row index outside the stored array")
        cr, cc = self.chunks
        chunk_of = rows // cr
        starts = np.r_[0, np.nonzero(np.diff(chunk_of))[0] + 1, rows.size]
        for a, b in zip(starts[:-1], starts[1:]):
            i = int(chunk_of[a])
            for j in range(-(-self.shape[1] // cc)):
                c0, c1 = j * cc, min((j + 1) * cc, self.shape[1])
                out[a:b, c0:c1] = self._chunk(i, j)[rows[a:b] % cr, : c1 - c0]
        return out

    def __getitem__(self, rows):
        """``reader[index_array]`` / ``reader[a:b]``: what ``engine.load_G_rows`` asks of its source."""
        if isinstance(rows, slice):
            rows = np.arange(*rows.indices(self.shape[0]))
        return self.read_rows(rows)
