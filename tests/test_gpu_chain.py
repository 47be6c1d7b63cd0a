"""Chained launches of the wavefront kernel (lgca_step_wave.cu: consecutive launches of one lgca_b200_step call overlap,
ordered by per-chunk completion counters): the result must not depend on whether launches overlap.  Checked against the
oracle on lattices with many chunks, against the strictly serial launch order (LGCA_B200_FLAG_NO_CHAIN) at config size,
and across snapshots (three rotating plane sets) and plan changes (mask upload between calls)."""
import numpy as np
import pytest

from cpu_checkers import Oracle, OracleRng, fnv1a64

pytestmark = pytest.mark.gpu

NO_RESIDENT = 4
NO_CHAIN = 64
FORCE_CHAIN = 128   # the library chains on its own only where one launch fills the machine (>= ~20 M sites)


def _engine(o, k_fuse, flags):
    import lgca_b200
    e = lgca_b200.Engine(o.model, o.dim_x, o.dim_y, k_fuse=k_fuse, flags=flags)
    e.upload(o.state, o.cell_type, o.rnd)
    return e


# lattices of 0.3 - 1 M cells: tens of chunks per launch, several launches per call in flight
CASES = [
    ("FHP_III", (2048, 512), "periodic", 6, 61),
    ("FHP_III", (3000, 300), "karman", 6, 43),
    ("FHP_III", (1480, 740), "pipe", 5, 37),
    ("FHP_II", (2048, 256), "reflecting_back", 4, 41),
    ("FHP_II", (1000, 250), "reflecting_forward", 3, 31),
    ("FHP_I", (1400, 700), "pipe", 5, 26),
    ("HPP", (2048, 501), "periodic", 6, 50),     # odd height: last chunk one row short
    ("HPP", (997, 333), "reflecting_back", 8, 49),
    ("FHP_III", (4096, 24), "periodic", 6, 37),  # one or two chunks: a tile waits for its own chunk and wraps onto it
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%dx%d-%s-k%d" % (c[0], c[1][0], c[1][1], c[2], c[3]))
def test_chained_launches_match_oracle(case):
    model, dims, bc, k, steps = case
    o = Oracle(model, dims=dims, cg=1, rng=OracleRng(11))
    o.apply_bc(bc)
    o.init("random")
    e = _engine(o, k, NO_RESIDENT | FORCE_CHAIN)
    s = _engine(o, k, NO_RESIDENT | NO_CHAIN)
    for n in (steps, 2 * k, steps):
        e.step(n)
        s.step(n)
        o.step(n)
        got = e.download()
        assert np.array_equal(s.download(), o.state), "serial launches, +%d" % n
        if not np.array_equal(got, o.state):
            bad = np.nonzero(got != o.state)[0]
            raise AssertionError("chained launches differ after +%d steps: %d cells, first at (x=%d,y=%d)" % (
                n, bad.size, bad[0] % o.dim_x, bad[0] // o.dim_x))
    assert e.count_particles() == o.n_particles()
    e.close()
    s.close()


def test_chain_across_snapshots_and_mask_upload():
    """Snapshots rotate three plane sets under the chain; a mask upload re-plans (counters start over)."""
    o = Oracle("FHP_III", dims=(1024, 512), cg=4, rng=OracleRng(3))
    o.apply_bc("pipe")
    o.init("random")
    e = _engine(o, 4, NO_RESIDENT | FORCE_CHAIN)
    for n in (17, 9, 30):
        e.step(n); o.step(n)
        e.snapshot(); o.snapshot()
        e.step(n + 1); o.step(n + 1)
        o.post_process()
        f = e.post_process(cell=True, mean=False, exact=True)
        assert np.array_equal(f["cell_density"], o.cell_density)
        assert np.array_equal(e.download(), o.state)
    # new walls: the plan (and its counters) are rebuilt
    o.apply_bc("karman")
    e.upload(cell_type=o.cell_type)
    e.step(29); o.step(29)
    assert np.array_equal(e.download(), o.state)
    e.close()


@pytest.mark.parametrize("cfg", [("FHP_III", 16384, 8192, "karman", 6, 126), ("FHP_III", 4400, 2200, "karman", 5, 100),
                                 ("HPP", 4096, 4096, "periodic", 6, 126), ("FHP_II", 8192, 4096, "reflecting_back", 6, 67)],
                         ids=lambda c: "%s-%dx%d" % (c[0], c[1], c[2]))
def test_chain_equals_serial_at_config_size(cfg):
    """BASELINE configs C3 / C2 and the Karman default at full size: chained == strictly serial launches, repeatedly
    (a race between overlapping launches would not hit the same cells twice)."""
    import lgca_b200
    model, dx, dy, bc, k, steps = cfg
    hashes = []
    n0 = None
    for flags in (NO_RESIDENT | FORCE_CHAIN, NO_RESIDENT | NO_CHAIN, NO_RESIDENT):
        e = lgca_b200.Engine(model, dx, dy, k_fuse=k, flags=flags)
        e.apply_bc_device(bc)
        e.init_random_device(seed=5)
        n0 = e.count_particles()
        hs = []
        for _ in range(3):
            e.step(steps)
            hs.append(fnv1a64(e.download()))
        assert e.count_particles() == n0
        hashes.append(hs)
        e.close()
    assert hashes[0] == hashes[1] == hashes[2]


# Row strips: chained blocks of the native ring (the write-after-read wait for my own ghost-row push moves from the stream
# into the edge tiles).  Strips share device 0 here (plain peer pointers); bench.py checks the same invariance across real
# GPUs in every multi-GPU run (multi_gpu_parity).
STRIP_CASES = [
    ("FHP_III", (2048, 768), "periodic", 6, [0, 0], 50),
    ("FHP_III", (1024, 960), "karman", 4, [0, 0, 0], 45),
    ("FHP_II", (2048, 512), "reflecting_back", 6, [0, 0], 37),
    ("HPP", (1024, 600), "reflecting_forward", 6, [0, 0, 0], 61),
    ("FHP_I", (1400, 700), "pipe", 5, [0, 0], 26),
]


@pytest.mark.parametrize("case", STRIP_CASES, ids=lambda c: "%s-%dx%d-%s-k%d-%dstrips" % (c[0], c[1][0], c[1][1], c[2], c[3], len(c[4])))
def test_chained_strips_match_oracle(case):
    import lgca_b200
    model, dims, bc, k, devs, steps = case
    o = Oracle(model, dims=dims, cg=1, rng=OracleRng(13))
    o.apply_bc(bc)
    o.init("random")
    g = lgca_b200.Group(model, dims[0], dims[1], n_gpus=len(devs), dev_ids=devs, k_fuse=k, flags=FORCE_CHAIN)
    g.upload(o.state, o.cell_type, o.rnd)
    for n in (steps, 1, 3 * k, steps):
        g.step(n)
        o.step(n)
        got = g.download()
        if not np.array_equal(got, o.state):
            bad = np.nonzero(got != o.state)[0]
            raise AssertionError("chained strips differ after +%d steps: %d cells, first at (x=%d,y=%d)" % (
                n, bad.size, bad[0] % o.dim_x, bad[0] // o.dim_x))
        g.snapshot()   # rotates the three plane sets on every strip
    assert g.count_particles() == o.n_particles()
    g.close()


def test_chained_strips_equal_whole_lattice_at_size():
    """8192 x 8192 FHP-III with walls as 2 strips (chained by the library's own choice) vs one whole lattice, serial."""
    import lgca_b200
    dx, dy = 8192, 8192
    hashes = []
    for kind in ("strips", "whole"):
        if kind == "strips":
            e = lgca_b200.Group("FHP_III", dx, dy, n_gpus=2, dev_ids=[0, 0])
        else:
            e = lgca_b200.Engine("FHP_III", dx, dy, flags=NO_CHAIN)
        e.apply_bc_device("karman")
        e.init_random_device(seed=9)
        hs = []
        for _ in range(2):
            e.step(97)
            hs.append(fnv1a64(e.download()))
        hashes.append(hs)
        e.close()
    assert hashes[0] == hashes[1]
