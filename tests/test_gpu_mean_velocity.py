"""Order-exact mean velocity on the device (run with -m gpu): lgca_b200_mean_velocity_exact / lgca_b200_group_mean_velocity_exact
must return the digits of the reference's one-thread get_mean_velocity (src/omp_lattice.cpp:508-557) -- sequential float32 sums
over the FLUID cells of the output buffer -- on whole lattices and on row strips, for every model, with walls of both kinds
inside the lattice (solid cells hold particles while they bounce and must not be counted)."""
import os

import numpy as np
import pytest

from cpu_checkers import Oracle, OracleRng, Ref, ref_available

pytestmark = pytest.mark.gpu

CASES = [
    ("FHP_III", (4400, 440), "karman", 3),     # ragged last segment (4400 = 4*1024 + 304), Karman obstacle
    ("FHP_I", (1400, 700), "pipe", 5),         # BASELINE config C1
    ("FHP_II", (1031, 96), "reflecting_forward", 2),  # dim_x % 32 != 0, slip walls
    ("HPP", (2048, 512), "reflecting_back", 4),
    ("HPP", (21, 10), "pipe", 1),              # one short segment per row
    ("FHP_III", (96, 48), "periodic", 6),
]


def oracle_case(model, dims, bc, steps, seed=23):
    o = Oracle(model, dims=dims, cg=4, bf_dir=b"x", rng=OracleRng(seed))
    o.apply_bc(bc)
    o.init("random")
    o.step(steps)
    o.snapshot()
    o.post_process()
    return o


@pytest.mark.parametrize("model,dims,bc,steps", CASES, ids=lambda c: str(c))
def test_exact_mean_velocity_whole_lattice(model, dims, bc, steps):
    import lgca_b200
    o = oracle_case(model, dims, bc, 0)
    e = lgca_b200.Engine(model, dims[0], dims[1], cg_radius=0, bf_dir="x")
    e.upload(o.state, o.cell_type, o.rnd)
    for n in (0, steps, 1):
        if n:
            e.step(n)
            o.step(n)
        e.snapshot()
        e.step(2)  # computed from the snapshot, not from the live state
        o.snapshot()
        o.post_process()
        o.step(2)
        sums, fluid = e.mean_velocity_exact()
        assert fluid == int((o.cell_type == 0).sum())
        got = (sums / np.float32(fluid)).astype(np.float32)
        want = np.asarray(o.mean_velocity(), np.float32)
        assert got.tobytes() == want.tobytes(), (n, got, want)
    e.close()


@pytest.mark.parametrize("nstrips", [2, 3])
@pytest.mark.parametrize("model,dims,bc,steps", CASES[:4], ids=lambda c: str(c))
def test_exact_mean_velocity_on_strips(model, dims, bc, steps, nstrips):
    import lgca_b200
    o = oracle_case(model, dims, bc, 0)
    g = lgca_b200.Group(model, dims[0], dims[1], n_gpus=nstrips, dev_ids=[0] * nstrips, cg_radius=0, bf_dir="x")
    g.upload(o.state, o.cell_type, o.rnd)
    g.step(steps)
    o.step(steps)
    g.snapshot()
    o.snapshot()
    o.post_process()
    got = g.mean_velocity_exact()
    want = np.asarray(o.mean_velocity(), np.float32)
    assert got.tobytes() == want.tobytes(), (got, want)
    g.close()


@pytest.mark.skipif(not ref_available(), reason="prebuilt oracle/_ref did not travel")
def test_exact_mean_velocity_vs_reference_karman_default():
    """App size: the unmodified reference's own get_mean_velocity on the Karman default lattice (4400 x 2200; the x sum runs
    up to ~2^21 where one ulp is 1/8 .. 1/4 -- every rounding matters)."""
    import lgca_b200
    r = Ref("FHP_III", "karman", 80, 0.3, 20)
    r.set_threads(1)
    r.apply_bc("karman")
    r.init("random")
    e = lgca_b200.Engine("FHP_III", r.dim_x, r.dim_y, cg_radius=20, bf_dir="x")
    e.upload(r.state, r.cell_type, r.rnd)
    for n in (0, 5):
        if n:
            r.set_threads(os.cpu_count() or 1)
            r.step(n)
            r.set_threads(1)
            e.step(n)
        r.snapshot()
        r.post_process()
        e.snapshot()
        sums, fluid = e.mean_velocity_exact()
        got = (sums / np.float32(fluid)).astype(np.float32)
        want = np.asarray(r.mean_velocity(), np.float32)
        assert got.tobytes() == want.tobytes(), (n, got, want)
    r.close()
    e.close()
