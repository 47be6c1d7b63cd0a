"""Chained vs serial launches of the wavefront kernel over lattice sizes (development helper: where does chaining pay?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgca_b200
from lgca_b200.capi import FLAG_NO_CELL_FIELDS, FLAG_NO_RESIDENT, FLAG_NO_CHAIN

SIZES = [(1400, 700), (2048, 2048), (4400, 2200), (4096, 4096), (8192, 4096), (8192, 8192), (16384, 8192), (16384, 16384)]
for model, bc in (("FHP_III", "periodic"), ("FHP_III", "karman"), ("HPP", "periodic"), ("FHP_I", "pipe")):
    for dx, dy in SIZES:
        res = {}
        for name, fl in (("serial", FLAG_NO_CHAIN), ("chained", 0)):
            e = lgca_b200.Engine(model, dx, dy, flags=FLAG_NO_CELL_FIELDS | FLAG_NO_RESIDENT | fl)
            e.apply_bc_device(bc)
            e.init_random_device(1)
            k = e.info().k_fuse
            steps = max(4 * k, min(1200, int(4e10 / (dx * dy)) // k * k))
            e.timed_steps(steps)
            best = min(e.timed_steps(steps) for _ in range(3))
            res[name] = best / steps * 1e3
            e.close()
        print(f"{model} {bc} {dx}x{dy} k={k}: serial {res['serial']:.2f} us/update  chained {res['chained']:.2f}  x{res['serial'] / res['chained']:.3f}", flush=True)
