"""Timing of the body force on a Karman-default lattice: whole lattice vs 2 strips, device vs host path (development helper)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import lgca_b200
rng = np.random.default_rng(1)
for strips in (1, 2):
    for flags in (0, 16):
        g = lgca_b200.Group("FHP_III", 4400, 2200, n_gpus=strips, dev_ids=[0] * strips, cg_radius=20, bf_dir="x", flags=flags)
        g.apply_bc_device("karman")
        g.init_random_device(3)
        g.step(10)
        g.snapshot()
        draws = rng.integers(0, 2**31 - 1, 1_000_000).astype(np.int32)
        g.sync()
        ts = []
        for _ in range(5):
            t = time.perf_counter()
            used, rev = g.body_force(96800, draws)
            g.sync()
            ts.append(time.perf_counter() - t)
            g.step(5)
            g.snapshot()
            g.sync()
        print("strips %d flags %2d: body_force(96800) used %d reverted %d: %s ms" % (strips, flags, used, rev, ["%.2f" % (x * 1e3) for x in ts]), flush=True)
        g.close()
