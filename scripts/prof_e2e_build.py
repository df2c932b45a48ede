"""cProfile of the host side of one end-to-end build step (new simulation -> sim.G), the `e2e` leg of bench.py."""
import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
wl = bench.make_workload("cfg4-shard", 1, 2500)
def step():
    sim = bench.make_simulation(wl, bench.make_survey(wl, wl["locations"]))
    sim.G
    torch.cuda.synchronize()
    return sim
s = step(); del s
for _ in range(2):
    t = time.perf_counter(); s = step(); dt = time.perf_counter() - t; del s
    print("step ms", 1e3 * dt)
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    s = step(); del s
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
