"""Phase timing of the SM-resident kernel (needs a -DLGCA_RES_TIMING build: scripts/build_variant.sh ab_timing.so -DLGCA_RES_TIMING;
   LGCA_B200_LIB=$PWD/ab_timing.so python scripts/res_timing.py)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgca_b200
for name, model, dx, dy, bc in (("C1", "FHP_I", 1400, 700, "pipe"), ("karman", "FHP_III", 4400, 2200, "karman"), ("hpp1024", "HPP", 1024, 1024, "periodic"), ("fhp256", "FHP_III", 256, 256, "periodic")):
    for k, fl in ((0, 8), (2, 8), (4, 8), (0, 8 | 32)):
        e = lgca_b200.Engine(model, dx, dy, k_fuse=k, flags=fl | 1)
        e.apply_bc_device(bc)
        e.init_random_device(1)
        e.step(1000)
        ms = e.timed_steps(1000)
        print(name, "k", k, "flags", fl, "us/update %.3f" % ms, flush=True)
        e.close()
