"""cProfile of the host-side cost of Jvec / Jtvec through the public API."""
import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
wl = bench.make_workload("cfg4-shard", 1, 1000)
sim = bench.make_simulation(wl, bench.make_survey(wl, wl["locations"]))
nP = sim._n_columns(); nD = sim.survey.nD
rng = np.random.default_rng(0)
m, v, w = rng.uniform(size=nP), rng.uniform(size=nP), rng.uniform(size=nD)
t = time.perf_counter(); sim.G; torch.cuda.synchronize(); print("G", time.perf_counter() - t)
for _ in range(2):
    sim.Jvec(m, v); sim.Jtvec(m, w)
torch.cuda.synchronize()
t = time.perf_counter(); sim.Jvec(m, v); t1 = time.perf_counter(); sim.Jtvec(m, w); t2 = time.perf_counter()
print("Jvec ms", 1e3 * (t1 - t), "Jtvec ms", 1e3 * (t2 - t1))
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    sim.Jvec(m, v); sim.Jtvec(m, w)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
