import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lgca_b200
from lgca_b200.capi import FLAG_NO_CELL_FIELDS
for model, dx, dy, bc in (("HPP", 4096, 4096, "periodic"), ("FHP_III", 4400, 2200, "karman"), ("FHP_I", 1400, 700, "pipe"), ("HPP", 2048, 2048, "periodic")):
    e = lgca_b200.Engine(model, dx, dy, flags=FLAG_NO_CELL_FIELDS)
    e.apply_bc_device(bc); e.init_random_device(1)
    e.timed_steps(600)
    best = min(e.timed_steps(1200) for _ in range(3))
    print(f"{model} {dx}x{dy} resident/default: {best/1200*1e3:.3f} us/update  {dx*dy*1200/(best*1e-3):.4g} sites/s", flush=True)
    e.close()
