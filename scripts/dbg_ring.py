import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import lgca_b200
from lgca_b200.ring import partition_rows
from cpu_checkers import Oracle, OracleRng
from test_gpu_strips import LocalRing
model, dims, bc, nstrips, k = "FHP_II", (512, 128), "reflecting_back", 4, int(sys.argv[1]) if len(sys.argv) > 1 else 3
o = Oracle(model, dims=dims, cg=1, rng=OracleRng(6)); o.apply_bc(bc); o.init("random")
parts = partition_rows(dims[1], nstrips, 2)
engines = []
for y0, rows in parts:
    e = lgca_b200.Engine(model, dims[0], dims[1], k_fuse=k, y_begin=y0, y_rows=rows)
    sl = slice(y0 * dims[0], (y0 + rows) * dims[0]); e.upload(o.state[sl], o.cell_type[sl], o.rnd); engines.append(e)
flags = [e.wall_flags() for e in engines]
for e in engines: e.set_wall_flags(any(f[0] for f in flags), any(f[1] for f in flags))
LocalRing(engines).exchange(1)
desc = [e.ring_export() for e in engines]; n = len(engines)
for r, e in enumerate(engines): e.ring_connect(desc[(r - 1) % n], desc[(r + 1) % n])
for e in engines: e.ring_start()
done = 0
for steps in [1] * 40:
    for e in engines: e.ring_step(steps)
    o.step(steps); done += steps
    got = np.concatenate([e.download() for e in engines])
    if not np.array_equal(got, o.state):
        bad = np.nonzero(got != o.state)[0]
        ys = sorted(set((bad // dims[0]).tolist()))
        print("MISMATCH after", done, "steps:", bad.size, "cells; rows", ys[:20], "x range", (bad % dims[0]).min(), (bad % dims[0]).max())
        break
else:
    print("ok", done)
