"""lgca_b200_group_*: one lattice on several GPUs of one box driven by ONE host process -- the handle behind the C++
B200_Lattice and the apps' --gpus switch (run with -m gpu).  Results must be identical to the oracle's single-lattice
run (decomposition invariance) for every operation of the Lattice interface.  On a one-GPU box the strips share the
device (dev_ids repeats it: same code path, plain peer pointers); tests marked `multi` need >= 2 physical GPUs and
exercise the NVLink peer stores for real."""
import json
import os

import numpy as np
import pytest

from cpu_checkers import Oracle, OracleRng, fnv1a64

pytestmark = pytest.mark.gpu

REF_RUNS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_runs.json")))


def n_devices():
    try:
        import lgca_b200
        return lgca_b200.load_library().lgca_b200_device_count()
    except Exception:
        return 0


def device_sets():
    """(id, dev_ids) pairs: strips sharing device 0, and -- when the box has them -- strips on distinct GPUs."""
    out = [pytest.param([0], id="1strip"), pytest.param([0, 0], id="2strips-1gpu"), pytest.param([0, 0, 0], id="3strips-1gpu")]
    return out


def group_from(o, devs, k_fuse=0, cg=0, **kw):
    import lgca_b200
    bf = o.p.bf_dir if isinstance(o.p.bf_dir, int) else (o.p.bf_dir[0] if o.p.bf_dir not in (b"", b"\0") else 0)
    g = lgca_b200.Group(o.model, o.dim_x, o.dim_y, n_gpus=len(devs), dev_ids=devs, cg_radius=cg, bf_dir=bf, k_fuse=k_fuse, **kw)
    g.upload(o.state, o.cell_type, o.rnd)
    return g


CASES = [
    ("FHP_III", (256, 96), "karman", 4, 0),
    ("FHP_II", (512, 144), "reflecting_back", 4, 3),
    ("FHP_I", (304, 72), "reflecting_forward", 4, 2),
    ("HPP", (640, 96), "pipe", 4, 0),
    ("FHP_III", (48, 48), "pipe", 4, 0),          # dim_x < 64: generic kernel on the strips
]


@pytest.mark.parametrize("host_bf", [False, True], ids=["device-bf", "host-bf"])
@pytest.mark.parametrize("devs", device_sets())
@pytest.mark.parametrize("model,dims,bc,cg,k", CASES, ids=lambda c: str(c))
def test_group_equals_oracle(model, dims, bc, cg, k, devs, host_bf):
    o = Oracle(model, dims=dims, cg=cg, bf_dir=b"x", rng=OracleRng(17))
    o.apply_bc(bc)
    o.init("random")
    # body force: the device-side path (classification per strip, gains summed over peer copies, prefix + stop rule, scatter)
    # and the gather -> ordered host replay -> apply path (flag 16) must both reproduce the reference's sequential loop
    g = group_from(o, devs, k_fuse=k, cg=cg, flags=16 if host_bf else 0)
    assert np.array_equal(g.download(), o.state)
    assert g.count_particles() == o.n_particles()
    o.rng = OracleRng(5)
    draw_rng, stream, pos = OracleRng(5), [], 0   # the same rand() stream, handed to the group in order
    for rounds, n in enumerate((1, 6, 13, 5)):
        g.step(n)
        o.step(n)
        g.snapshot()
        o.snapshot()
        o.post_process()
        f = g.post_process(cell=True, mean=True, exact=True)
        for name in ("cell_density", "cell_momentum", "mean_density", "mean_momentum"):
            assert np.array_equal(f[name], getattr(o, name)), (name, rounds)
        np.testing.assert_allclose(g.mean_velocity(), o.mean_velocity(), rtol=0, atol=1e-6)
        assert g.mean_velocity_exact().tobytes() == np.asarray(o.mean_velocity(), np.float32).tobytes()  # the one-thread digits
        # body force on the live state right after the snapshot (the canonical schedule), then step on
        forcing = (0, 7, 120, 900)[rounds]
        used_o, rev_o = o.body_force(forcing)
        # surplus draws must be left unconsumed; the reference's cap of 2*num_cells draws is the caller's to enforce
        give = min(used_o + 40, 2 * o.num_cells)
        while len(stream) < pos + give:
            stream.append(draw_rng.rand())
        used_g, rev_g = g.body_force(forcing, np.array(stream[pos:pos + give], np.int32))
        assert (used_g, rev_g) == (used_o, rev_o)
        pos += used_o
        assert np.array_equal(g.download(), o.state), "after force %d" % forcing
    g.step(9)
    o.step(9)
    assert np.array_equal(g.download(), o.state)
    assert g.count_particles() == o.n_particles()
    g.close()


def test_group_device_init_matches_single_handle():
    """Device-side BC painting + counter-hash init are keyed on the GLOBAL cell: a group of strips starts from, and
    evolves to, exactly the state of a single whole-lattice handle."""
    import lgca_b200
    dims = (2048, 512)
    one = lgca_b200.Group("FHP_III", *dims, n_gpus=1, dev_ids=[0], cg_radius=16)
    four = lgca_b200.Group("FHP_III", *dims, n_gpus=4, dev_ids=[0, 0, 0, 0], cg_radius=16)
    for g in (one, four):
        g.apply_bc_device("karman")
        g.init_random_device(seed=9)
    assert np.array_equal(one.download(), four.download())
    for n in (6, 31):
        one.step(n)
        four.step(n)
        one.snapshot()
        four.snapshot()
        a = one.post_process(cell=False, mean=True, exact=False)
        b = four.post_process(cell=False, mean=True, exact=False)
        assert np.array_equal(a["mean_density"], b["mean_density"])
        assert np.array_equal(a["mean_momentum"], b["mean_momentum"])
    assert np.array_equal(one.download(), four.download())
    assert one.count_particles() == four.count_particles()
    t = four.timed_steps(12)
    assert t > 0
    one.step(12)
    assert np.array_equal(one.download(), four.download())
    one.close()
    four.close()


def run_canonical_schedule(g, o, steps, marks, golden):
    """The viewers' serialised tick (apps/karman/karman_viewer.cpp:100-184): mean velocity -> body force -> 5 steps ->
    snapshot -> post-process.  The mean velocity is the reference's sequential float32 loop over the per-cell host
    fields (oracle restatement, as B200_Lattice does in C++)."""
    rng = o.rng  # continues the reference's rand() stream after ctor + init_random
    fifo = []
    u = np.float32(o.u)
    g.snapshot()
    f = g.post_process(cell=True, mean=True)
    forcing = o.initial_forcing()
    done = 0
    while done < steps:
        o.cell_density[:] = f["cell_density"]
        o.cell_momentum[:] = f["cell_momentum"]
        mv = o.mean_velocity()
        tick_end = done + 5
        if str(tick_end) in golden["mv_at_tick_start"]:
            np.testing.assert_allclose(mv, golden["mv_at_tick_start"][str(tick_end)], rtol=0, atol=5e-7)
        if mv[0] < u:
            if float(mv[0]) > 0.9 * float(u):
                forcing = o.equilibrium_forcing()
            if str(tick_end) in golden["forcing_at_tick"]:
                assert forcing == golden["forcing_at_tick"][str(tick_end)]
            remaining, it_max, used_total, first = forcing, 2 * o.num_cells, 0, True
            while (first or remaining > 0) and used_total < it_max:
                need = max(4096, remaining * 12)
                while len(fifo) < need:
                    fifo.append(rng.rand())
                used, rev = g.body_force(remaining, np.array(fifo[:need], np.int32))
                del fifo[:used]
                used_total += used
                remaining -= rev
                first = False
        g.step(5)
        done += 5
        g.snapshot()
        f = g.post_process(cell=True, mean=True)
        if done in marks:
            assert fnv1a64(g.download()) == golden["hashes"][str(done)], "step %d" % done


@pytest.mark.parametrize("devs", [pytest.param([0], id="1gpu"), pytest.param([0, 0], id="2strips")])
def test_karman_default_1000_steps_bit_exact(devs):
    """THE north-star target: lgca-karman at the app's defaults (FHP-III, Re 80, Ma 0.3, cg 20 -> 4400 x 2200, walls +
    cylinder) on the canonical schedule with body force and mean-velocity feedback for 1000 steps, state hashes equal
    to the UNMODIFIED reference's at steps 0/5/100/500/1000 (tests/golden/reference_runs.json, generated from
    oracle/_ref by scripts/gen_golden_reference.py).  Also as two row strips (halo ring, strip body force)."""
    gold = REF_RUNS["karman_default"]
    o = Oracle(gold["model"], *gold["ctor"])
    assert [o.dim_x, o.dim_y] == gold["dims"]
    o.apply_bc("karman")
    o.init("random")
    assert o.hash() == gold["hashes"]["0"]
    assert fnv1a64(o.rnd) == gold["chirality_hash"] and fnv1a64(o.cell_type) == gold["cell_type_hash"]
    g = group_from(o, devs, cg=gold["ctor"][3])
    run_canonical_schedule(g, o, 1000, {5, 100, 500, 1000}, gold)
    assert g.count_particles() == gold["particles"]
    g.close()


def test_pipe_default_fhp3_1000_steps_bit_exact():
    """lgca-pipe at ITS defaults (FHP-III 1480 x 740) as three strips, same schedule, 1000 steps vs the reference."""
    gold = REF_RUNS["pipe_default_fhp3"]
    o = Oracle(gold["model"], *gold["ctor"])
    o.apply_bc("pipe")
    o.init("random")
    assert o.hash() == gold["hashes"]["0"]
    g = group_from(o, [0, 0, 0], cg=gold["ctor"][3])
    run_canonical_schedule(g, o, 1000, {5, 100, 500, 1000}, gold)
    assert g.count_particles() == gold["particles"]
    g.close()


def test_config_width_spot_checks():
    """BASELINE configs C2 and C5 at config width against the reference's own stepping: HPP 4096 x 4096 (diffusion disc,
    periodic) and one FHP-III strip of the 32768-wide lattice (32768 x 512), 13 steps each."""
    import lgca_b200
    gold = REF_RUNS["hpp_4096"]
    o = Oracle(gold["model"], *gold["ctor"])
    o.apply_bc(gold["bc"])
    o.init(gold["init"])
    assert [o.dim_x, o.dim_y] == gold["dims"] and o.hash() == gold["hashes"]["0"]
    e = lgca_b200.Engine(o.model, o.dim_x, o.dim_y, cg_radius=16)
    e.upload(o.state, o.cell_type, o.rnd)
    e.step(1)
    assert fnv1a64(e.download()) == gold["hashes"]["1"]
    e.step(12)
    assert fnv1a64(e.download()) == gold["hashes"]["13"]
    assert e.count_particles() == gold["particles"]
    e.close()

    gold = REF_RUNS["fhp3_32768x512"]
    o = Oracle(gold["model"], dims=tuple(gold["dims"]), cg=gold["cg"])
    o.apply_bc(gold["bc"])
    o.init(gold["init"])
    assert o.hash() == gold["hashes"]["0"] and fnv1a64(o.rnd) == gold["chirality_hash"]
    for devs in ([0], [0, 0]):
        g = group_from(o, devs, cg=gold["cg"])
        g.step(1)
        assert fnv1a64(g.download()) == gold["hashes"]["1"]
        g.step(12)
        assert fnv1a64(g.download()) == gold["hashes"]["13"]
        g.close()


# ---- two or more PHYSICAL GPUs: the peer stores cross NVLink for real -------------------------------------------------
multi = pytest.mark.skipif(n_devices() < 2, reason="needs >= 2 GPUs")


@multi
@pytest.mark.parametrize("model,dims,bc,cg,k", CASES[:4], ids=lambda c: str(c))
def test_group_on_distinct_gpus(model, dims, bc, cg, k):
    devs = list(range(min(n_devices(), 4)))
    if dims[1] // max(2 * cg, 2) < len(devs):
        devs = devs[:2]
    test_group_equals_oracle(model, dims, bc, cg, k, devs)


@multi
def test_big_strips_on_distinct_gpus_equal_one_gpu():
    """Edge tiles waiting in-kernel (ld.acquire.sys) on flags a PEER GPU publishes, ghost rows arriving over NVLink during
    the kernel, coherent loads of those rows: 8192-wide FHP-III strips with walls, 79 steps, against one GPU."""
    import lgca_b200
    n = min(n_devices(), 4)
    dims = (8192, 512 * n)
    one = lgca_b200.Group("FHP_III", *dims, n_gpus=1, dev_ids=[0], cg_radius=16)
    many = lgca_b200.Group("FHP_III", *dims, n_gpus=n, dev_ids=list(range(n)), cg_radius=16)
    for g in (one, many):
        g.apply_bc_device("karman")
        g.init_random_device(seed=4)
    for steps in (1, 36, 42):
        one.step(steps)
        many.step(steps)
        many.snapshot()
        one.snapshot()
        assert fnv1a64(one.download()) == fnv1a64(many.download())
    a = one.post_process(cell=False, mean=True, exact=False)
    b = many.post_process(cell=False, mean=True, exact=False)
    assert np.array_equal(a["mean_density"], b["mean_density"]) and np.array_equal(a["mean_momentum"], b["mean_momentum"])
    one.close()
    many.close()
