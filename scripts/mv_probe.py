"""Timing breakdown of the order-exact mean velocity (device summaries + copies vs host walk) on app-sized lattices.
   gpurun -- python scripts/mv_probe.py"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import lgca_b200

for model, dx, dy, bc in (("FHP_I", 1400, 700, "pipe"), ("FHP_III", 1480, 740, "pipe"), ("FHP_III", 4400, 2200, "karman")):
    e = lgca_b200.Engine(model, dx, dy, cg_radius=10, bf_dir="x")
    e.apply_bc_device(bc)
    e.init_random_device(7)
    e.step(50)
    e.snapshot()
    e.mean_velocity_exact()
    s0 = e.mean_velocity_stats()
    t = time.perf_counter()
    n = 20
    for _ in range(n):
        e.step(5)
        e.snapshot()
        sums, fluid = e.mean_velocity_exact()
    dt = (time.perf_counter() - t) / n
    s1 = e.mean_velocity_stats()
    d = {k: (s1[k] - s0[k]) / n for k in s1}
    print("%-8s %5dx%-5d %.3f ms per (5 steps + snapshot + mean velocity): device+copies %.3f ms, walk %.3f ms, segments fast %.0f walked %.0f  v=(%.6f, %.6f)" % (
        model, dx, dy, dt * 1e3, d["device_ns"] / 1e6, d["walk_ns"] / 1e6, d["segments_fast"], d["segments_walked"], sums[0] / fluid, sums[1] / fluid))
    e.close()
