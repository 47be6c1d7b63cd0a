"""A-B timing of chained vs strictly serial launches of the wavefront kernel (development helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgca_b200
from lgca_b200.capi import FLAG_NO_CELL_FIELDS, FLAG_NO_RESIDENT, FLAG_NO_CHAIN

CONFIGS = [("FHP_III", 16384, 8192, "karman", 0, 240), ("FHP_III", 32768, 32768, "periodic", 0, 120),
           ("FHP_II", 16384, 8192, "reflecting_back", 0, 240), ("HPP", 4096, 4096, "periodic", 0, 480),
           ("FHP_III", 4400, 2200, "karman", 0, 500), ("FHP_I", 1400, 700, "pipe", 0, 1000),
           ("FHP_III", 16384, 8192, "karman", 4, 240), ("FHP_III", 16384, 8192, "karman", 5, 240)]

if __name__ == "__main__":
    sel = sys.argv[1:] 
    for model, dx, dy, bc, k, steps in CONFIGS:
        res = {}
        for name, fl in (("serial", FLAG_NO_CHAIN), ("chained", 0), ("serial2", FLAG_NO_CHAIN), ("chained2", 0)):
            e = lgca_b200.Engine(model, dx, dy, k_fuse=k, flags=FLAG_NO_CELL_FIELDS | FLAG_NO_RESIDENT | fl)
            e.apply_bc_device(bc)
            e.init_random_device(1)
            n0 = e.count_particles()
            e.timed_steps(steps)
            best = min(e.timed_steps(steps) for _ in range(3))
            assert e.count_particles() == n0
            res[name] = dx * dy * steps / (best * 1e-3)
            kk = e.info().k_fuse
            e.close()
        print(f"{model} {dx}x{dy} {bc} k={kk}: serial {res['serial']:.4g} / {res['serial2']:.4g}  chained {res['chained']:.4g} / "
              f"{res['chained2']:.4g}  x{max(res['chained'], res['chained2']) / max(res['serial'], res['serial2']):.3f}", flush=True)
