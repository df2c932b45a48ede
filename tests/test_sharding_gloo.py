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
        # the slab-overlapped Jtvec (ShardedDeviceOps.gemv_t_allreduce_dev): column slabs, one async
        # all-reduce per slab, waits at the end -- with a numpy stand-in for the device kernel
        from simpeg_b200.potential_fields import base

        base.JTVEC_SLABS, base.JTVEC_SLAB_ALIGN, base.JTVEC_SLAB_MIN = 3, 4, 8
        G_t = torch.from_numpy(G_local)

        class Engine:
            nP, device = G_local.shape[1], torch.device("cpu")

            def gemv_t_dev(self, y, out=None, cols=None):
                c0, c1 = (0, self.nP) if cols is None else cols
                z = torch.empty(self.nP, dtype=torch.float64) if out is None else out
                z[c0:c1] = G_t[:, c0:c1].T @ y
                return z

        Engine.torch = torch
        ops = base.ShardedDeviceOps(Engine(), coll, sh)
        assert len(base.column_slabs(Engine.nP, base.JTVEC_SLABS)) == 3
        gtv_slabs = ops.gemv_t_allreduce_dev(torch.from_numpy(case["v"][sh.local_rows])).numpy()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), fwd=fwd, gtv=gtv, gtv_slabs=gtv_slabs)
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
        assert np.max(np.abs(z["gtv_slabs"] - case["gtv"])) <= 1e-13 * np.max(np.abs(case["gtv"]))
