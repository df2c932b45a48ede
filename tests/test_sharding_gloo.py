"""
CPU, world_size 2, gloo: the receiver-row sharding and the two exchanges of
the path (all-gather of rows, all-reduce of column partials).  The oracle
stands in for the local device compute -- the plumbing under test is exactly
what the CUDA engine uses on NCCL.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import OracleSimulation
        from simpeg_b200.sharding import Collectives, ReceiverSharding

        case = load_golden(name)
        coll = Collectives()
        assert (coll.rank, coll.world_size) == (rank, world)
        sh = ReceiverSharding(case["rlist"], coll.rank, coll.world_size)
        local = OracleSimulation(case["physics"], case["nodes"], case["cell_nodes"],
                                 sh.local_receiver_list, b0=case["b0"],
                                 model_type=case["model_type"], parallel=False)
        G_local = local.sensitivity_matrix()
        assert np.array_equal(G_local, case["G"][sh.local_rows])           # row ownership
        fwd = coll.all_gather_rows(torch.from_numpy(local.forward(case["model"])), sh).numpy()
        gtv = coll.all_reduce_sum(torch.from_numpy(local.gt_dot_v(case["v"][sh.local_rows]))).numpy()
        # replicated host vector uploaded in rank slices and all-gathered (odd length: padded tail)
        model = case["model"]
        full = coll.all_gather_slices(lambda lo, hi: torch.from_numpy(model[lo:hi].copy()), model.size)
        assert np.array_equal(full.numpy(), model)
        odd = np.arange(7.0)
        assert np.array_equal(coll.all_gather_slices(lambda lo, hi: torch.from_numpy(odd[lo:hi].copy()), 7).numpy(), odd)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), fwd=fwd, gtv=gtv)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["grav_two_objects", "mag_tmi_bx_vector"])
def test_two_rank_sharding(name, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    case = load_golden(name)
    for rank in range(world):
        z = np.load(tmp_path / f"r{rank}.npz")
        assert np.array_equal(z["fwd"], case["fwd"])                      # all-gather is exact
        assert np.max(np.abs(z["gtv"] - case["gtv"])) <= 1e-13 * np.max(np.abs(case["gtv"]))
