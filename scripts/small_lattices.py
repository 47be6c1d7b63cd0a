"""SM-resident kernel vs. the HBM-streaming wavefront kernel on the lattices that fit on chip (development helper):
device time per update for calls of 5 / 100 / 1000 steps (5 = one tick of the reference's viewers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgca_b200
from lgca_b200.capi import FLAG_NO_CELL_FIELDS, FLAG_NO_RESIDENT, FLAG_FORCE_RESIDENT

CONFIGS = [
    ("C1 lgca-pipe FHP-I 1400x700", "FHP_I", 1400, 700, "pipe"),
    ("lgca-pipe default FHP-III 1480x740", "FHP_III", 1480, 740, "pipe"),
    ("lgca-karman default FHP-III 4400x2200", "FHP_III", 4400, 2200, "karman"),
    ("C2 lgca-diffusion HPP 4096x4096", "HPP", 4096, 4096, "periodic"),
    ("HPP 4096x4096 bounce-back frame", "HPP", 4096, 4096, "reflecting_back"),
    ("FHP-III 2048x2048 periodic", "FHP_III", 2048, 2048, "periodic"),
    ("HPP 2048x2048 periodic", "HPP", 2048, 2048, "periodic"),
    ("HPP 1024x1024 periodic", "HPP", 1024, 1024, "periodic"),
    ("FHP-II 4096x2048 box", "FHP_II", 4096, 2048, "reflecting_back"),
    ("FHP-III 256x256 periodic (app 'periodic' default)", "FHP_III", 256, 256, "periodic"),
]
if os.environ.get("SMALL_ONLY_RESIDENT"):
    pass
KS = [int(a) for a in sys.argv[1:]] or [0]
print("| lattice | kernel | k | us/update @5 | us/update @100 | us/update @1000 | site updates/s @1000 |")
print("|---|---|---:|---:|---:|---:|---:|")
for name, model, dx, dy, bc in CONFIGS:
    kernels = [("wave", FLAG_NO_RESIDENT), ("resident", FLAG_FORCE_RESIDENT)]
    if os.environ.get("SMALL_WPT"):  # A-B of the 1-word / 4-word variants of the resident kernel
        kernels = [("resident static", FLAG_FORCE_RESIDENT), ("resident dynamic", FLAG_FORCE_RESIDENT | 32)]
    elif os.environ.get("SMALL_ONLY_RESIDENT"):
        kernels = kernels[1:]
    for kernel, flags in kernels:
        for k in (KS if kernel.startswith("resident") else [0]):
            e = lgca_b200.Engine(model, dx, dy, k_fuse=k, flags=flags | FLAG_NO_CELL_FIELDS)
            e.apply_bc_device(bc)
            e.init_random_device(1)
            n0 = e.count_particles()
            out = []
            for n, reps in ((5, 200), (100, 20), (1000, 3)):
                e.timed_steps(n)
                best = 1e30
                for _ in range(3):
                    t = 0.0
                    for _ in range(reps):
                        t += e.timed_steps(n)
                    best = min(best, t / reps)
                out.append(best * 1e3 / n)
            assert e.count_particles() == n0
            print("| %s | %s | %d | %.3f | %.3f | %.3f | %.3g |" % (name, kernel, k, out[0], out[1], out[2], dx * dy / (out[2] * 1e-6)), flush=True)
            e.close()
