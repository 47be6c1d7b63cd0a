"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C-ABI, must reproduce
the packed occupation state of the reference bit-exactly.  Checkers: the pinned oracle
(oracle/lgca_oracle.c), the known answers of SURVEY.md Appendix B, and -- where the prebuilt
oracle/_ref travelled along -- the unmodified reference itself."""
import json
import os

import numpy as np
import pytest

from cpu_checkers import Oracle, OracleRng, Ref, fnv1a64, ref_available

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_b.json")))


def engine_from(o, k_fuse=0, flags=0, **kw):
    import lgca_b200
    bf = o.p.bf_dir if isinstance(o.p.bf_dir, int) else (o.p.bf_dir[0] if o.p.bf_dir not in (b"", b"\0") else 0)
    cg = int(o.p.cg_radius)
    if cg and (o.dim_x % (2 * cg) or o.dim_y % (2 * cg) or o.dim_x < 4 * cg):
        cg = 0
    e = lgca_b200.Engine(o.model, o.dim_x, o.dim_y, cg_radius=cg, bf_dir=bf, k_fuse=k_fuse, flags=flags, **kw)
    e.upload(o.state, o.cell_type, o.rnd)
    return e


# flags: 2 = generic one-word-per-thread kernel; 4 = never use the SM-resident kernel, i.e. the HBM-streaming wavefront
# kernel with k fused steps; 0 = library default: lattices that fit on chip run the SM-resident kernel (k = steps per
# ghost-row exchange between CTAs) whenever a call advances >= 2 steps; 8 = force it wherever the lattice fits
VARIANTS = [pytest.param(dict(flags=2), id="simple"), pytest.param(dict(k_fuse=1, flags=4), id="k1"),
            pytest.param(dict(k_fuse=2, flags=4), id="k2"), pytest.param(dict(k_fuse=3, flags=4), id="k3"),
            pytest.param(dict(k_fuse=4, flags=4), id="k4"), pytest.param(dict(k_fuse=6, flags=4), id="k6"),
            pytest.param(dict(flags=4), id="default"),
            pytest.param(dict(k_fuse=1, flags=8), id="res1"), pytest.param(dict(k_fuse=2, flags=8), id="res2"),
            pytest.param(dict(k_fuse=3, flags=8), id="res3"), pytest.param(dict(k_fuse=5, flags=8), id="res5"),
            pytest.param(dict(flags=8), id="res"), pytest.param(dict(), id="auto"),
            # 32: the resident kernel's dynamic four-word-group mapping instead of static word ownership
            pytest.param(dict(k_fuse=1, flags=8 | 32), id="res1-dyn"), pytest.param(dict(k_fuse=4, flags=8 | 32), id="res4-dyn"),
            pytest.param(dict(flags=8 | 32), id="res-dyn")]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("case", GOLD["b2_pure_stepping"], ids=lambda c: "%s-%s" % (c["model"], c["bc"]))
def test_appendix_b_hashes(case, variant):
    o = Oracle(case["model"], *case["ctor"])
    o.apply_bc(case["bc"])
    o.init(case["init"])
    e = engine_from(o, **variant)
    assert fnv1a64(e.download()) == case["hashes"]["0"]  # pack/unpack round trip
    done = 0
    for s in sorted(int(k) for k in case["hashes"]):
        e.step(s - done)
        done = s
        assert fnv1a64(e.download()) == case["hashes"][str(s)], "step %d" % s
    assert e.count_particles() == case["particles"]
    e.close()


SHAPES = [
    # model, dims, bc, init
    ("HPP", (64, 32), "periodic", "random"),
    ("HPP", (37, 24), "reflecting_back", "random"),
    ("HPP", (1000, 31), "reflecting_forward", "random"),
    ("FHP_I", (45, 26), "reflecting_forward", "random"),
    ("FHP_I", (1400, 40), "pipe", "random"),
    ("FHP_II", (33, 18), "reflecting_forward", "random"),
    ("FHP_II", (960, 64), "reflecting_back", "random"),
    ("FHP_II", (2048, 70), "karman", "random"),
    ("FHP_III", (97, 40), "periodic", "random"),
    ("FHP_III", (21, 10), "pipe", "random"),
    ("FHP_III", (1, 2), "periodic", "random"),
    ("FHP_III", (31, 6), "periodic", "random"),
    ("FHP_III", (32, 4), "periodic", "random"),
    ("FHP_III", (3000, 130), "karman", "random"),
    ("FHP_III", (4096, 256), "periodic", "random"),
]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "%s-%dx%d-%s" % (s[0], s[1][0], s[1][1], s[2]))
def test_oracle_parity_shapes(shape, variant):
    model, dims, bc, init = shape
    o = Oracle(model, dims=dims, cg=1, rng=OracleRng(7))
    o.apply_bc(bc)
    o.init(init)
    e = engine_from(o, **variant)
    assert np.array_equal(e.download(), o.state)
    for n in (1, 1, 2, 3, 5, 8, 13):
        e.step(n)
        o.step(n)
        got = e.download()
        if not np.array_equal(got, o.state):
            bad = np.nonzero(got != o.state)[0]
            raise AssertionError("mismatch after +%d steps: %d cells, first at (x=%d,y=%d): got %02x want %02x" % (
                n, bad.size, bad[0] % o.dim_x, bad[0] // o.dim_x, got[bad[0]], o.state[bad[0]]))
    assert e.count_particles() == o.n_particles()
    e.close()


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not present")
def test_reference_parity_karman_small():
    r = Ref("FHP_III", "karman", 10, 0.2, 16)
    r.apply_bc("karman")
    r.init("random")
    import lgca_b200
    e = lgca_b200.Engine("FHP_III", r.dim_x, r.dim_y, cg_radius=16, bf_dir="x")
    e.upload(r.state, r.cell_type, r.rnd)
    for n in (1, 9, 90):
        r.step(n)
        e.step(n)
        assert np.array_equal(e.download(), r.state)
    r.close()
    e.close()


@pytest.mark.parametrize("model,tc,Re,cg,bc", [("FHP_III", "karman", 10, 16, "karman"), ("FHP_I", "pipe", 20, 4, "pipe"),
                                               ("HPP", "pipe", 20, 4, "pipe"), ("FHP_II", "box", 127, 8, "reflecting_back")])
def test_post_process_parity(model, tc, Re, cg, bc):
    o = Oracle(model, tc, Re, 0.2, cg)
    o.apply_bc(bc)
    o.init("random")
    e = engine_from(o)
    e.step(7)
    o.step(7)
    e.snapshot()
    e.step(3)  # the snapshot must be insulated from further stepping
    o.snapshot()
    o.post_process()
    f = e.post_process(cell=True, mean=True, exact=True)
    assert np.array_equal(f["cell_density"], o.cell_density)
    assert np.array_equal(f["cell_momentum"], o.cell_momentum)
    assert np.array_equal(f["mean_density"], o.mean_density)
    assert np.array_equal(f["mean_momentum"], o.mean_momentum)
    fast = e.post_process(cell=False, mean=True, exact=False)
    assert np.array_equal(fast["mean_density"], o.mean_density)
    assert np.array_equal(fast["mean_momentum"][0::2], o.mean_momentum[0::2])
    # popcount path: momentum-y differs from the reference's sequential float32 sum only by rounding
    np.testing.assert_allclose(fast["mean_momentum"][1::2], o.mean_momentum[1::2], rtol=0, atol=2e-6)
    mv = e.mean_velocity()
    np.testing.assert_allclose(mv, o.mean_velocity(), rtol=0, atol=1e-6)
    e.close()


@pytest.mark.parametrize("model,bf", [("HPP", b"x"), ("HPP", b"y"), ("FHP_I", b"x"), ("FHP_II", b"y"), ("FHP_III", b"x")])
def test_body_force_parity(model, bf):
    o = Oracle(model, dims=(120, 48), cg=4, bf_dir=bf, rng=OracleRng(3))
    o.apply_bc("pipe")
    o.init("random")
    e = engine_from(o)
    o.rng = OracleRng(99)
    total_used = 0
    for forcing in (0, 1, 5, 40, 400):
        used_o, rev_o = o.body_force(forcing)
        # the engine gets the same stream from the same position, plus surplus draws it must not consume
        draw_rng = OracleRng(99)
        for _ in range(total_used):
            draw_rng.rand()
        draws = np.array([draw_rng.rand() for _ in range(used_o + 50)], np.int32)
        used_e, rev_e = e.body_force(forcing, draws)
        assert (used_e, rev_e) == (used_o, rev_o)
        assert np.array_equal(e.download(), o.state)
        total_used += used_o
        e.step(2)
        o.step(2)
    e.close()


@pytest.mark.parametrize("case", GOLD["b3_pipe_schedule"], ids=lambda c: c["model"])
def test_canonical_pipe_1000_steps(case):
    """BASELINE config C1: the app schedule (mean velocity -> body force -> 5 steps -> snapshot ->
    post-process) for 1000 steps; state hash must equal the reference's."""
    o = Oracle(case["model"], *case["ctor"])
    o.apply_bc("pipe")
    o.init("random")
    assert o.hash() == case["hashes"]["0"]
    e = engine_from(o)
    rng = o.rng  # continues the reference's rand() stream after ctor + init_random
    fifo = []
    ct = o.cell_type
    fluid = ct == 0
    u = np.float32(o.u)
    e.snapshot()
    f = e.post_process(cell=True, mean=True)
    forcing = o.initial_forcing()
    for tick in range(1, 201):
        # order-exact mean velocity on the host fields, like the reference (src/omp_lattice.cpp:508-557)
        o.cell_density[:] = f["cell_density"]
        o.cell_momentum[:] = f["cell_momentum"]
        mv = o.mean_velocity()
        step = tick * 5
        if str(step) in case["mv_x_at_tick_start"]:
            assert abs(float(mv[0]) - case["mv_x_at_tick_start"][str(step)]) < 5e-7
        if mv[0] < u:
            if float(mv[0]) > 0.9 * float(u):
                forcing = o.equilibrium_forcing()
            remaining, it_max, used_total = forcing, 2 * o.num_cells, 0
            first = True
            while (first or remaining > 0) and used_total < it_max:
                need = max(4096, remaining * 12)
                while len(fifo) < need:
                    fifo.append(rng.rand())
                used, rev = e.body_force(remaining, np.array(fifo[:need], np.int32))
                del fifo[:used]
                used_total += used
                remaining -= rev
                first = False
        e.step(5)
        e.snapshot()
        f = e.post_process(cell=True, mean=True)
        if str(step) in case["hashes"]:
            assert fnv1a64(e.download()) == case["hashes"][str(step)], "step %d" % step
    assert e.count_particles() == case["particles"]
    e.close()


def test_device_init_and_conservation():
    import lgca_b200
    e = lgca_b200.Engine("FHP_III", 2048, 1024, cg_radius=16, bf_dir="x")
    e.apply_bc_device("karman")
    e.init_random_device(seed=5)
    s0 = e.download()
    n0 = e.count_particles()
    assert n0 == int(np.unpackbits(s0).sum())
    dens = n0 / (2048 * 1024 * 7)
    assert 0.12 < dens < 0.16
    # the device-painted BC equals the oracle's painter
    o = Oracle("FHP_III", dims=(2048, 1024), cg=16)
    o.apply_bc("karman")
    solid = o.cell_type != 0
    assert not s0[solid].any()
    e.step(50)
    assert e.count_particles() == n0
    # decomposition-free check against the oracle from the downloaded initial state
    o.state[:] = s0
    e2 = lgca_b200.Engine("FHP_III", 2048, 1024, flags=2)
    e2.apply_bc_device("karman")
    e2.init_random_device(seed=5)
    e2.step(50)
    assert np.array_equal(e.download(), e2.download())
    e.close()
    e2.close()


def test_two_host_threads_step_and_post_process():
    """The reference's viewers drive one lattice from two host threads (apps/pipe/pipe_viewer.cpp:105,150):
    stepping + snapshot on one, post_process on the other.  The snapshot must insulate the two."""
    import threading
    import lgca_b200
    o = Oracle("FHP_III", "karman", 10, 0.2, 16)
    o.apply_bc("karman")
    o.init("random")
    e = engine_from(o)
    n0 = e.count_particles()
    errors, densities = [], []
    e.snapshot()

    def poster():
        try:
            for _ in range(30):
                f = e.post_process(cell=True, mean=True, exact=True)
                densities.append(float(f["cell_density"].sum()))
        except Exception as ex:  # pragma: no cover
            errors.append(ex)

    t = threading.Thread(target=poster)
    t.start()
    for _ in range(30):
        e.step(5)
        e.snapshot()
    t.join()
    assert not errors
    # every post-processed snapshot is a consistent state: its per-cell density sums to the particle count
    assert all(abs(d - n0) < 0.5 for d in densities), (n0, densities[:5])
    o.step(150)
    assert np.array_equal(e.download(), o.state)
    e.close()


@pytest.mark.parametrize("model", ["HPP", "FHP_I", "FHP_II", "FHP_III"])
@pytest.mark.parametrize("fill", ["empty", "full", "dense"])
def test_extreme_occupancies(model, fill):
    """Empty lattice, completely full lattice and a dense random state (every collision-table entry is hit
    many times, unlike the 1/NUM_DIR density of init_random)."""
    o = Oracle(model, dims=(200, 66), cg=1, rng=OracleRng(21))
    o.apply_bc("reflecting_back")
    nd = o.num_dir
    fluid = o.cell_type == 0
    rs = np.random.RandomState(4)
    if fill == "full":
        o.state[fluid] = (1 << nd) - 1
    elif fill == "dense":
        o.state[fluid] = rs.randint(0, 1 << nd, size=int(fluid.sum())).astype(np.uint8)
    # solid cells may hold particles too (they stream in and bounce): exercise that as well
    if fill == "dense":
        o.state[~fluid] = rs.randint(0, 1 << nd, size=int((~fluid).sum())).astype(np.uint8)
    for variant in (dict(flags=2), dict(k_fuse=3, flags=4), dict(flags=4), dict(k_fuse=3, flags=8), dict()):
        e = engine_from(o, **variant)
        oo = Oracle(model, dims=(200, 66), cg=1)
        oo.state[:], oo.cell_type[:], oo.rnd[:] = o.state, o.cell_type, o.rnd
        for n in (1, 7, 12):
            e.step(n)
            oo.step(n)
            assert np.array_equal(e.download(), oo.state), (variant, n)
        assert e.count_particles() == oo.n_particles()
        e.close()


def test_full_size_karman_parity_and_conservation():
    """BASELINE config C3 at its FULL size (FHP-III 16384 x 8192, walls + cylinder): the fused kernel against the
    oracle for 12 steps (bit-exact on all 134 M cells), then size-independent properties over 1000 more steps:
    particle conservation and agreement of the two independent CUDA kernels."""
    import lgca_b200
    dx, dy = 16384, 8192
    o = Oracle("FHP_III", dims=(dx, dy), cg=16, bf_dir=b"x")
    o.apply_bc("karman")
    e = lgca_b200.Engine("FHP_III", dx, dy, cg_radius=16, bf_dir="x")
    e.apply_bc_device("karman")
    e.init_random_device(seed=3)
    o.state[:] = e.download()          # device-generated occupancy (host rand() would take minutes)
    assert not o.state[o.cell_type != 0].any()   # device painter == oracle painter: no particles seeded in solids
    e.upload(cell_type=o.cell_type, rnd_bits=o.rnd)   # the oracle's chirality field and cell types
    n0 = e.count_particles()
    assert n0 == o.n_particles()
    e.step(12)
    o.step(12)
    got = e.download()
    assert np.array_equal(got, o.state)
    e2 = lgca_b200.Engine("FHP_III", dx, dy, flags=2)   # generic one-word-per-thread kernel
    e2.upload(got, o.cell_type, o.rnd)
    e.step(1000)
    assert e.count_particles() == n0
    e2.step(60)
    e3 = lgca_b200.Engine("FHP_III", dx, dy, k_fuse=4)
    e3.upload(got, o.cell_type, o.rnd)
    e3.step(60)
    assert np.array_equal(e2.download(), e3.download())
    for x in (e, e2, e3):
        x.close()


@pytest.mark.gpu
@pytest.mark.parametrize("writer", ["step1", "step2", "body_force", "upload", "init", "none"])
def test_snapshot_is_insulated_from_every_writer(writer):
    """copy_data_to_output_buffer (src/lattice.cpp:437-441) freezes the state.  Whole lattices snapshot without a copy
    (the live buffer becomes the snapshot, copy-on-write for in-place writers): whatever touches the live state
    afterwards, post_process must still see the state of the snapshot, and the live state must be the oracle's."""
    o = Oracle("FHP_III", dims=(192, 64), cg=4, bf_dir=b"x", rng=OracleRng(5))
    o.apply_bc("pipe")
    o.init("random")
    e = engine_from(o)
    e.step(4)
    o.step(4)
    e.snapshot()
    e.snapshot()           # twice in a row: still the same state
    o.snapshot()
    o.post_process()
    if writer == "step1":
        e.step(1); o.step(1)
    elif writer == "step2":
        e.step(1); e.step(1); e.step(5); o.step(7)
    elif writer == "body_force":
        o.rng = OracleRng(77)
        used_o, rev_o = o.body_force(25)
        rng = OracleRng(77)
        draws = np.array([rng.rand() for _ in range(used_o + 10)], np.int32)
        assert e.body_force(25, draws) == (used_o, rev_o)
    elif writer == "upload":
        o.step(3)
        e.upload(state=o.state)
    elif writer == "init":
        e.init_random_device(9)
    f = e.post_process(cell=True, mean=True, exact=True)
    assert np.array_equal(f["cell_density"], o.cell_density)
    assert np.array_equal(f["cell_momentum"], o.cell_momentum)
    assert np.array_equal(f["mean_density"], o.mean_density)
    if writer != "init":
        assert np.array_equal(e.download(), o.state)
        # and the next snapshot follows the live state again
        e.step(2); o.step(2)
        e.snapshot(); o.snapshot(); o.post_process()
        f = e.post_process(cell=True, mean=False, exact=True)
        assert np.array_equal(f["cell_density"], o.cell_density)
        assert np.array_equal(e.download(), o.state)
    e.close()
