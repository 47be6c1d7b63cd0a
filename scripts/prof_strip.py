"""A few ring blocks of a 2-strip lattice on ONE device (for ncu captures of the EDGE tiles: in-kernel epoch wait +
coherent loads; development helper).  python scripts/prof_strip.py FHP_III 16384 4096 periodic 10"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgca_b200

model, dx, dy, bc, blocks = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
g = lgca_b200.Group(model, dx, dy, n_gpus=2, dev_ids=[0, 0], flags=1)
g.apply_bc_device(bc)
g.init_random_device(1)
k = g.info().k_fuse
g.step(k * blocks)
g.sync()
print("done", k, g.launch_count())
