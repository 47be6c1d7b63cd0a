"""Pins the CPU oracle (oracle/lgca_oracle.c) -- CPU only, no GPU.

1. against the known answers of SURVEY.md Appendix B (tests/golden/appendix_b.json), which were
   produced by the unmodified reference;
2. against the reference itself (oracle/_ref/liblgca_ref.so) on further seeded configurations,
   including odd widths, slip walls, every model, body force, post-processing and mean velocity.
"""
import json
import os

import numpy as np
import pytest

from cpu_checkers import (MODELS, Oracle, OracleRng, Ref, collide_table, fnv1a64, ref_available)

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_b.json")))
needs_ref = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")


def test_rng_matches_glibc_known_values():
    g = OracleRng(1)
    assert [g.rand() for _ in range(3)] == GOLD["rand_first"]


def test_rng_matches_libc_stream():
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 2, 12345):
        libc.srand(seed)
        g = OracleRng(seed)
        assert [g.rand() for _ in range(2000)] == [libc.rand() for _ in range(2000)]


@pytest.mark.parametrize("case", GOLD["b4_default_sizes"], ids=lambda c: "%s-%s" % (c["model"], c["ctor"][0]))
def test_sizing(case):
    tc, Re, Ma, cg = case["ctor"]
    from cpu_checkers import _Params, _oracle_lib
    import ctypes as C
    p = _Params()
    assert _oracle_lib().lgca_oracle_params_init(C.byref(p), MODELS[case["model"]], tc.encode(), Re, Ma, cg) == 0
    assert [p.dim_x, p.dim_y] == case["dims"]
    if "coarse" in case:
        assert [p.coarse_dim_x, p.coarse_dim_y] == case["coarse"]
    if "initial_forcing" in case:
        assert _oracle_lib().lgca_oracle_initial_forcing(C.byref(p)) == case["initial_forcing"]


@pytest.mark.parametrize("model", ["HPP", "FHP_I", "FHP_II", "FHP_III"])
def test_collision_truth_tables(model):
    key = "FHP_II" if model == "FHP_III" else model  # FHP-III as coded == FHP-II
    changed = {int(k): v for k, v in GOLD["collision_changed_states"][key].items()}
    table = collide_table(model)
    nd = 4 if model == "HPP" else (6 if model == "FHP_I" else 7)
    for s in range(1 << nd):
        for p in (0, 1):
            expect = changed[s][p] if s in changed else s
            assert table[(s, p)] == expect, (model, s, p)
            assert bin(table[(s, p)]).count("1") == bin(s).count("1")  # mass conservation


def test_single_collision_trace():
    c = GOLD["b1_single_collision"]
    o = Oracle(c["model"], *c["ctor"])
    assert [o.dim_x, o.dim_y] == c["dims"]
    bits = "".join(str((o.rnd[i >> 3] >> (i & 7)) & 1) for i in range(16))
    assert bits == c["first16_chirality"]
    o.apply_bc(c["bc"])
    o.init(c["init"])
    for step, expect in enumerate(c["trace"]):
        occ = sorted([int(cell % o.dim_x), int(cell // o.dim_x), d]
                     for cell in np.nonzero(o.state)[0] for d in range(7) if (o.state[cell] >> d) & 1)
        assert occ == sorted(expect), step
        o.step(1)


@pytest.mark.parametrize("case", GOLD["b2_pure_stepping"], ids=lambda c: "%s-%s" % (c["model"], c["bc"]))
def test_pure_stepping_hashes(case):
    o = Oracle(case["model"], *case["ctor"])
    assert [o.dim_x, o.dim_y] == case["dims"]
    o.apply_bc(case["bc"])
    o.init(case["init"])
    assert o.n_particles() == case["particles"]
    done = 0
    for s in sorted(int(k) for k in case["hashes"]):
        o.step(s - done)
        done = s
        assert o.hash() == case["hashes"][str(s)], "step %d" % s
    assert o.n_particles() == case["particles"]


def canonical_pipe_schedule(lat, ticks, pp_interval=5, on_tick=None):
    """Canonical headless schedule of one GUI tick (SURVEY 3.3 / apps/pipe/pipe_viewer.cpp:100-184)."""
    lat.snapshot()
    lat.post_process()
    forcing = lat.initial_forcing()
    u = np.float32(lat.u)
    for tick in range(1, ticks + 1):
        mv = lat.mean_velocity()
        if mv[0] < u:
            if float(mv[0]) > 0.9 * float(u):
                forcing = lat.equilibrium_forcing()
            lat.body_force(forcing)
        lat.step(pp_interval)
        lat.snapshot()
        lat.post_process()
        if on_tick:
            on_tick(tick, mv)


@pytest.mark.slow
@pytest.mark.parametrize("case", GOLD["b3_pipe_schedule"], ids=lambda c: c["model"])
def test_canonical_pipe_1000_steps(case):
    o = Oracle(case["model"], *case["ctor"])
    assert [o.dim_x, o.dim_y] == case["dims"]
    assert fnv1a64(o.rnd) == case["chirality_hash"]
    o.apply_bc("pipe")
    assert fnv1a64(o.cell_type) == case["cell_type_hash"]
    o.init("random")
    assert o.n_particles() == case["particles"]
    assert o.initial_forcing() == case["initial_forcing"]
    assert o.hash() == case["hashes"]["0"]
    seen = {}

    def on_tick(tick, mv):
        step = tick * 5
        if str(step) in case["hashes"]:
            seen[step] = o.hash()
        if str(step) in case["mv_x_at_tick_start"]:
            assert abs(float(mv[0]) - case["mv_x_at_tick_start"][str(step)]) < 5e-7, step

    canonical_pipe_schedule(o, 200, on_tick=on_tick)
    for k, v in case["hashes"].items():
        if int(k) > 0:
            assert seen[int(k)] == v, "step " + k
    assert o.n_particles() == case["particles"]


# ------------------------------------------------------------------------------------------------
# Oracle vs. the unmodified reference
# ------------------------------------------------------------------------------------------------
REF_CASES = [
    # model, test_case, Re, Ma, cg, bc, init, dims override
    ("HPP", "periodic", 63, 0.2, 4, "periodic", "random", None),
    ("HPP", "box", 40, 0.2, 4, "reflecting_back", "random", (37, 24)),
    ("FHP_I", "pipe", 10, 0.2, 2, "pipe", "random", None),
    ("FHP_I", "box", 40, 0.2, 4, "reflecting_forward", "random", (45, 26)),
    ("FHP_II", "karman", 4, 0.2, 4, "karman", "random", None),
    ("FHP_II", "box", 40, 0.2, 4, "reflecting_forward", "random", (33, 18)),
    ("FHP_III", "diffusion", 60, 0.2, 2, "reflecting_back", "diffusion", None),
    ("FHP_III", "periodic", 40, 0.2, 4, "periodic", "random", (97, 40)),
    ("FHP_III", "collision", 80, 0.2, 1, "pipe", "random", None),
]


@needs_ref
@pytest.mark.parametrize("case", REF_CASES, ids=lambda c: "%s-%s-%s" % (c[0], c[5], c[7]))
def test_oracle_matches_reference_stepping(case):
    model, tc, Re, Ma, cg, bc, init, dims = case
    r = Ref(model, tc, Re, Ma, cg, dims=dims)
    r.apply_bc(bc)
    r.init(init)
    if dims is None:
        o = Oracle(model, tc, Re, Ma, cg)
    else:
        o = Oracle(model, dims=dims, cg=cg)
    assert (o.dim_x, o.dim_y) == (r.dim_x, r.dim_y)
    # same BC painter result, then take the reference's own initial data
    o.apply_bc(bc)
    assert np.array_equal(o.cell_type, r.cell_type)
    o.state[:] = r.state
    o.rnd[:] = r.rnd
    for n in (1, 1, 3, 20):
        r.step(n)
        o.step(n)
        assert np.array_equal(o.state, r.state)
    r.snapshot(); r.post_process()
    o.snapshot(); o.post_process()
    assert np.array_equal(o.cell_density, r.cell_density)
    assert np.array_equal(o.cell_momentum, r.cell_momentum)
    if dims is None:  # coarse dims need dim % 2cg == 0
        assert np.array_equal(o.mean_density, r.mean_density)
        assert np.array_equal(o.mean_momentum, r.mean_momentum)
    assert np.array_equal(o.mean_velocity(), r.mean_velocity())
    assert o.n_particles() == r.n_particles()
    r.close()


@needs_ref
@pytest.mark.parametrize("model,tc,cg,bc", [("FHP_I", "pipe", 2, "pipe"), ("HPP", "pipe", 2, "pipe"),
                                            ("FHP_III", "karman", 4, "karman")])
def test_oracle_matches_reference_full_pipeline(model, tc, cg, bc):
    """ctor -> BC -> init_random -> canonical ticks incl. body force: identical rand() consumption."""
    Re = 10 if tc == "pipe" else 4
    r = Ref(model, tc, Re, 0.3, cg)
    o = Oracle(model, tc, Re, 0.3, cg)
    assert np.array_equal(o.rnd, r.rnd)
    for lat in (r, o):
        lat.apply_bc(bc)
        lat.init("random")
    assert np.array_equal(o.state, r.state)
    assert o.initial_forcing() == r.initial_forcing()
    assert o.equilibrium_forcing() == r.equilibrium_forcing()
    assert o.u == r.u

    def check(tick, mv):
        pass

    # run both through the schedule tick by tick and compare after each
    for lat in (r, o):
        lat.snapshot(); lat.post_process()
    forcing_r = forcing_o = o.initial_forcing()
    for tick in range(12):
        mv_r, mv_o = r.mean_velocity(), o.mean_velocity()
        assert np.array_equal(mv_r, mv_o)
        if mv_r[0] < np.float32(r.u):
            if float(mv_r[0]) > 0.9 * r.u:
                forcing_r = forcing_o = o.equilibrium_forcing()
            r.body_force(forcing_r)
            o.body_force(forcing_o)
            assert np.array_equal(o.state, r.state), "body force tick %d" % tick
        r.step(5); o.step(5)
        assert np.array_equal(o.state, r.state), "tick %d" % tick
        for lat in (r, o):
            lat.snapshot(); lat.post_process()
    r.close()


@needs_ref
@pytest.mark.parametrize("model,bf", [("HPP", b"y"), ("FHP_II", b"y"), ("FHP_I", b"x"), ("HPP", b"x")])
def test_oracle_body_force_directions(model, bf):
    r = Ref(model, "periodic", 30, 0.2, 4)
    r.set_bf_dir(bf)
    r.apply_bc("reflecting_back")
    r.init("random")
    o = Oracle(model, dims=(r.dim_x, r.dim_y), cg=4, bf_dir=bf)
    o.apply_bc("reflecting_back")
    o.state[:] = r.state
    # continue the reference's libc stream inside the oracle's generator: same seed, same position
    g = OracleRng(1)
    ndraw = r.num_cells + r.num_dir * int((r.cell_type == 0).sum())
    for _ in range(ndraw):
        g.rand()
    o.rng = g
    for forcing in (0, 1, 7, 50):
        r.body_force(forcing)
        o.body_force(forcing)
        assert np.array_equal(o.state, r.state)
    r.close()


# ---- tests/golden/reference_runs.json: known answers generated from the UNMODIFIED reference at app-default and
# config sizes (scripts/gen_golden_reference.py).  The oracle must reproduce them; the GPU tests then compare with them.
REF_RUNS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_runs.json")))


@pytest.mark.parametrize("name,bc,upto", [("pipe_default_fhp3", "pipe", 100), ("karman_default", "karman", 5)])
def test_oracle_reproduces_reference_app_runs(name, bc, upto):
    """Canonical tick schedule at the apps' default sizes: ctor sizing, chirality field, BC painter, init_random, forcing
    formulas, order-exact mean velocity, exact body force, stepping."""
    gold = REF_RUNS[name]
    o = Oracle(gold["model"], *gold["ctor"])
    assert [o.dim_x, o.dim_y] == gold["dims"]
    o.apply_bc(bc)
    o.init("random")
    assert fnv1a64(o.rnd) == gold["chirality_hash"] and fnv1a64(o.cell_type) == gold["cell_type_hash"]
    assert o.hash() == gold["hashes"]["0"] and o.n_particles() == gold["particles"]
    assert o.initial_forcing() == gold["initial_forcing"] and o.equilibrium_forcing() == gold["equilibrium_forcing"]
    o.snapshot(); o.post_process()
    forcing, done = o.initial_forcing(), 0
    while done < upto:
        mv = o.mean_velocity()
        if str(done + 5) in gold["mv_at_tick_start"]:
            assert [float(mv[0]), float(mv[1])] == gold["mv_at_tick_start"][str(done + 5)]
        if mv[0] < np.float32(o.u):
            if float(mv[0]) > 0.9 * o.u:
                forcing = o.equilibrium_forcing()
            o.body_force(forcing)
        o.step(5)
        done += 5
        o.snapshot(); o.post_process()
        if str(done) in gold["hashes"]:
            assert o.hash() == gold["hashes"][str(done)], done


def test_oracle_reproduces_reference_config_width_runs():
    gold = REF_RUNS["hpp_4096"]
    o = Oracle(gold["model"], *gold["ctor"])
    o.apply_bc(gold["bc"]); o.init(gold["init"])
    assert [o.dim_x, o.dim_y] == gold["dims"] and o.hash() == gold["hashes"]["0"]
    o.step(1)
    assert o.hash() == gold["hashes"]["1"]
    o.step(12)
    assert o.hash() == gold["hashes"]["13"] and o.n_particles() == gold["particles"]
    gold = REF_RUNS["fhp3_32768x512"]
    o = Oracle(gold["model"], dims=tuple(gold["dims"]), cg=gold["cg"])
    o.apply_bc(gold["bc"]); o.init(gold["init"])
    assert o.hash() == gold["hashes"]["0"] and fnv1a64(o.rnd) == gold["chirality_hash"]
    o.step(13)
    assert o.hash() == gold["hashes"]["13"]
