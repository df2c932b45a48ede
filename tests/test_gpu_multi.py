"""
GPU, 2 ranks, NCCL: the sharded simulation (receiver rows split over the
ranks, all-gather for dpred/Jvec, all-reduce for Jtvec/getJtJdiag) gives the
same answers as the single-process oracle.  Skipped with fewer than 2 GPUs.
"""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from sim_fixtures import magnetic_survey, make_mapping, small_mesh, susceptibilities
        from simpeg_b200.potential_fields import base, magnetics

        base.SLICED_UPLOAD_MIN = 0      # exercise the rank-sliced upload + all-gather of model vectors
        # ... and the slab-overlapped Jtvec (column slabs, async all-reduce per slab, D2H on a copy stream)
        base.JTVEC_SLABS, base.JTVEC_SLAB_ALIGN, base.JTVEC_SLAB_MIN = 3, 16, 32

        mesh = small_mesh()
        out = {}
        for store in ("ram", "forward_only"):
            sim = magnetics.Simulation3DIntegral(
                mesh, survey=magnetic_survey(["tmi", "bx"], two_objects=True),
                chiMap=make_mapping("identity_map", 3 * mesh.n_cells), model_type="vector",
                store_sensitivities=store, sensitivity_dtype=np.float64, engine="cuda")
            chi = susceptibilities(mesh, "vector")
            rng = np.random.default_rng(11)
            v = rng.uniform(size=chi.size)
            w = rng.uniform(size=sim.survey.nD)
            wd = rng.uniform(0.5, 1.5, size=sim.survey.nD)
            assert sim._coll.world_size == world
            assert sim._dev.n_rows < sim.survey.nD          # really sharded
            out[f"{store}_dpred"] = sim.dpred(chi)
            out[f"{store}_jvec"] = sim.Jvec(chi, v)
            out[f"{store}_jtvec"] = sim.Jtvec(chi, w)
            out[f"{store}_diag"] = sim.getJtJdiag(chi, W=sp.diags(wd))
            if store == "ram":
                # device-resident Gauss-Newton CG on the sharded G (one all-reduce per iteration)
                from simpeg_b200.gauss_newton import DeviceGaussNewtonCG

                scale = 1.0 / np.abs(out["ram_jtvec"]).max()
                solver = DeviceGaussNewtonCG(sim, data_weights=wd * scale, reg_hessian=np.ones(chi.size),
                                             beta=1e-2, cg_maxiter=6, cg_rtol=1e-8)
                solver.set_model(chi)
                out["ram_hvp"] = solver.hessian_vector(v)
                out["ram_step"] = solver.find_search_direction(v - 0.5)
                out["ram_cg_count"] = np.array(solver.cg_count)
        # store_sensitivities="disk": both ranks write their rows into ONE reference-format file
        sim = magnetics.Simulation3DIntegral(
            mesh, survey=magnetic_survey(["tmi", "bx"], two_objects=True),
            chiMap=make_mapping("identity_map", 3 * mesh.n_cells), model_type="vector",
            store_sensitivities="disk", sensitivity_dtype=np.float64, engine="cuda",
            sensitivity_path=os.path.join(out_dir, "sens.bin"))
        out["disk_dpred"] = sim.dpred(chi)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), **out)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_simulation(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from helpers import RTOL_NORMWISE, vec_rel
    from sim_fixtures import magnetic_survey, make_mapping, oracle_for, small_mesh, susceptibilities
    from simpeg_b200.potential_fields import magnetics

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    mesh = small_mesh()
    sim = magnetics.Simulation3DIntegral(
        mesh, survey=magnetic_survey(["tmi", "bx"], two_objects=True),
        chiMap=make_mapping("identity_map", 3 * mesh.n_cells), model_type="vector")
    G = oracle_for(sim).sensitivity_matrix()
    chi = susceptibilities(mesh, "vector")
    rng = np.random.default_rng(11)
    v = rng.uniform(size=chi.size)
    w = rng.uniform(size=sim.survey.nD)
    wd = rng.uniform(0.5, 1.5, size=sim.survey.nD)
    expected = {"dpred": G @ chi, "jvec": G @ v, "jtvec": G.T @ w,
                "diag": np.einsum("i,ij,ij->j", wd**2, G, G)}
    from helpers import rowwise_rel

    on_disk = np.memmap(tmp_path / "sens.bin", dtype=np.float64, mode="r", shape=G.shape, order="C")
    assert rowwise_rel(np.asarray(on_disk), G).max() <= RTOL_NORMWISE
    scale = 1.0 / np.abs(G.T @ w).max()
    H = 2.0 * G.T @ (((wd * scale) ** 2)[:, None] * G) + 1e-2 * np.eye(chi.size)
    from test_gpu_gauss_newton import reference_search_direction

    step, count, _ = reference_search_direction(H, v - 0.5, chi, np.ones(chi.size), -np.inf, np.inf,
                                                6, 1e-8, 0.0)
    for rank in range(world):
        z = np.load(tmp_path / f"r{rank}.npz")
        assert vec_rel(z["ram_hvp"], H @ v) <= RTOL_NORMWISE
        assert vec_rel(z["disk_dpred"], G @ chi) <= RTOL_NORMWISE
        assert int(z["ram_cg_count"]) == count
        assert vec_rel(z["ram_step"], step) <= 1e-7
        for store in ("ram", "forward_only"):
            for key, ref in expected.items():
                assert vec_rel(z[f"{store}_{key}"], ref) <= RTOL_NORMWISE, (rank, store, key)
