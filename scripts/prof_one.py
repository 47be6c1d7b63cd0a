"""Run a few launches of one configuration (for ncu captures; development helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgca_b200

model, dx, dy, bc, k, flags = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
launches = int(sys.argv[7]) if len(sys.argv) > 7 else 4
e = lgca_b200.Engine(model, dx, dy, k_fuse=k, flags=flags)
e.apply_bc_device(bc)
e.init_random_device(1)
e.step(max(k, 1) * launches)
e.sync()
print("done", e.launch_count())
