import os, sys
sys.path.insert(0, "/root/repo")
import lgca_b200
for dims, model, bc in (((1400, 700), "FHP_I", "pipe"), ((1480, 740), "FHP_III", "pipe"), ((4400, 2200), "FHP_III", "karman")):
    for k in (1, 2, 3, 4, 5, 6):
        e = lgca_b200.Engine(model, dims[0], dims[1], k_fuse=k); e.apply_bc_device(bc); e.init_random_device(1)
        e.timed_kernel(50); ms = min(e.timed_kernel(400) for _ in range(3))
        print(model, dims, "k=%d" % k, "%.2f us/launch  %.2f us/update" % (ms * 1e3, ms * 1e3 / k), flush=True)
        e.close()
    e = lgca_b200.Engine(model, dims[0], dims[1], flags=2); e.apply_bc_device(bc); e.init_random_device(1)
    e.timed_steps(50); ms = min(e.timed_steps(400) for _ in range(3)) / 400
    print(model, dims, "simple %.2f us/update" % (ms * 1e3)); e.close()
