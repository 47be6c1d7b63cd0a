"""Debug: group on distinct GPUs vs oracle -- where do the states differ?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import lgca_b200
from cpu_checkers import Oracle, OracleRng


def diff_rows(a, b, dx):
    bad = np.nonzero(a != b)[0]
    if bad.size == 0:
        return "equal"
    rows = np.unique(bad // dx)
    return "%d cells differ, rows %s" % (bad.size, rows[:24].tolist())


def run(model, dims, bc, cg, k, devs, steps, snapshot=False, upload=True):
    o = Oracle(model, dims=dims, cg=cg, bf_dir=b"x", rng=OracleRng(17))
    o.apply_bc(bc)
    o.init("random")
    g = lgca_b200.Group(o.model, o.dim_x, o.dim_y, n_gpus=len(devs), dev_ids=devs, cg_radius=cg, bf_dir="x", k_fuse=k)
    g.upload(o.state, o.cell_type, o.rnd)
    out = []
    for n in steps:
        g.step(n)
        o.step(n)
        if snapshot:
            g.snapshot()
        out.append("+%d: %s" % (n, diff_rows(g.download(), o.state, o.dim_x)))
    g.close()
    print(model, dims, bc, "k=%d" % k, devs, "snap" if snapshot else "", " | ".join(out), flush=True)


for devs in ([0, 0], [0, 1], [1, 0]):
    for k in (1, 2, 3, 5):
        run("FHP_III", (256, 96), "karman", 4, k, devs, (1, 6, 13))
    run("FHP_III", (256, 96), "periodic", 4, 5, devs, (1, 6, 13))
    run("FHP_I", (256, 96), "periodic", 4, 5, devs, (1, 6, 13))
    run("HPP", (256, 96), "periodic", 4, 5, devs, (1, 6, 13))
    run("HPP", (640, 96), "pipe", 4, 0, devs, (1, 6, 13))
    run("FHP_III", (256, 96), "karman", 4, 5, devs, (1, 6, 13), snapshot=True)
    run("FHP_III", (2048, 512), "karman", 16, 5, devs, (1, 6, 13))
