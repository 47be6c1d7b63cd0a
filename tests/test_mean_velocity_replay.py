"""Order-exact mean velocity, host side (lgca_b200_mean_velocity_replay, no GPU needed): per-segment integer summaries per
float32 binade + the ordered walk must reproduce the reference's sequential float32 sums of m/rho over the FLUID cells
(OMP_Lattice::get_mean_velocity at one thread, src/omp_lattice.cpp:508-557) bit for bit -- checked against an independent
numpy restatement (np.cumsum accumulates sequentially in float32) and against the pinned oracle."""
import numpy as np
import pytest

from cpu_checkers import Oracle, OracleRng

SIN = np.float32(0.866025388)
VX = {4: [1, 0, -1, 0], 7: [1, 0.5, -0.5, -1, -0.5, 0.5, 0]}
VY = {4: [0, 1, 0, -1], 7: [0, SIN, SIN, 0, -SIN, -SIN, 0]}
ND = {"HPP": 4, "FHP_I": 6, "FHP_II": 7, "FHP_III": 7}


@pytest.fixture(scope="module")
def capi():
    from lgca_b200.build import build_library
    build_library()
    from lgca_b200 import capi
    return capi


def addend_tables(model):
    """v_x, v_y of every state byte with the reference's float32 operations (cell_post_process + the division)."""
    nd = ND[model]
    vx, vy = np.zeros(128, np.float32), np.zeros(128, np.float32)
    tx, ty = VX[4 if nd == 4 else 7], VY[4 if nd == 4 else 7]
    for b in range(1 << nd):
        mx, my, rho = np.float32(0), np.float32(0), 0
        for d in range(nd):
            ns = (b >> d) & 1
            rho += ns
            mx = np.float32(mx + np.float32(ns) * np.float32(tx[d]))
            my = np.float32(my + np.float32(ns) * np.float32(ty[d]))
        if rho:
            vx[b] = mx / np.float32(rho)
            vy[b] = my / np.float32(rho)
    return vx, vy


def sequential_sums(model, cls, start=(0.0, 0.0)):
    vx, vy = addend_tables(model)
    out = []
    for tab, s0 in ((vx, start[0]), (vy, start[1])):
        seq = np.concatenate([[np.float32(s0)], tab[cls]]).astype(np.float32)
        out.append(np.cumsum(seq, dtype=np.float32)[-1])
    return np.array(out, np.float32)


def random_classes(model, n, rng, density, solid_frac=0.05, drift=0.0):
    nd = ND[model]
    bits = rng.random((n, nd)) < density
    if drift:  # more particles along +x than along -x: the x sum drifts through the binades like in a forced pipe
        bits[:, 0] = rng.random(n) < min(1.0, density + drift)
    cls = np.zeros(n, np.uint8)
    for d in range(nd):
        cls |= (bits[:, d].astype(np.uint8) << d)
    cls[rng.random(n) < solid_frac] = 0
    return cls


@pytest.mark.parametrize("model", ["HPP", "FHP_I", "FHP_II", "FHP_III"])
@pytest.mark.parametrize("dim_x,rows,density,drift", [(21, 10, 0.3, 0.0), (1400, 700, 0.25, 0.05), (4400, 600, 0.5, 0.1),
                                                      (1031, 997, 0.14, -0.05), (3000, 1000, 0.9, 0.0)])
def test_replay_equals_sequential_float32(capi, model, dim_x, rows, density, drift):
    rng = np.random.default_rng(dim_x * 7 + rows)
    cls = random_classes(model, dim_x * rows, rng, density, drift=drift)
    want = sequential_sums(model, cls)
    got, fast, walked = capi.mean_velocity_replay(model, cls, dim_x, rows)
    assert got.tobytes() == want.tobytes(), (got, want)
    if dim_x * rows > 500000:
        # whole-segment integer adds carry the work; only segments that cross a binade boundary are walked (the
        # undriven components are random walks around zero and cross boundaries more often)
        assert fast > 2 * walked


def test_replay_continues_running_sums(capi):
    """Strips chain: the second strip starts from the first strip's sums (and a negative accumulator works too)."""
    rng = np.random.default_rng(3)
    cls = random_classes("FHP_III", 2048 * 400, rng, 0.4, drift=-0.1)
    want = sequential_sums("FHP_III", cls)
    half = 2048 * 150
    s1, _, _ = capi.mean_velocity_replay("FHP_III", cls[:half], 2048, 150)
    s2, _, _ = capi.mean_velocity_replay("FHP_III", cls[half:], 2048, 250, sums=s1)
    assert s2.tobytes() == want.tobytes()
    assert want[0] < -1000  # the x sum really ran through negative binades


def test_replay_ties_and_binade_edges(capi):
    """Adversarial: only dyadic addends (rho = 1, 2, 4: exact ties at ulp 1, 1/2, 1/4) from accumulators parked just below
    binade boundaries."""
    model = "FHP_III"
    # state bytes with dyadic v_x: {d0}: 1; {d1}: 0.5; {d0,d1}: 0.75; {d1,d6}: 0.25; {d0,d1,d2,d6}: 0.25; {d3}: -1; {d2}: -0.5
    dyadic = np.array([0x01, 0x02, 0x03, 0x42, 0x47, 0x08, 0x04, 0x0C, 0x48], np.uint8)
    rng = np.random.default_rng(11)
    for start in (2.0 ** 23 - 3, 2.0 ** 24 - 5, 2.0 ** 22 - 1.5, 2.0 ** 21 - 0.75, -(2.0 ** 23) + 2, 8388607.5, 0.0, 2.0 ** 25 - 8):
        cls = dyadic[rng.integers(0, dyadic.size, 1024 * 9)]
        want = sequential_sums(model, cls, start=(start, -start))
        got, _, _ = capi.mean_velocity_replay(model, cls, 1024 * 3, 3, sums=(start, -start))
        assert got.tobytes() == want.tobytes(), (start, got, want)


@pytest.mark.parametrize("model,bc", [("FHP_III", "karman"), ("FHP_I", "pipe"), ("HPP", "pipe")])
def test_replay_matches_oracle_mean_velocity(capi, model, bc):
    o = Oracle(model, dims=(1200, 600), cg=10, bf_dir=b"x", rng=OracleRng(5))
    o.apply_bc(bc)
    o.init("random")
    for _ in range(3):
        o.step()
    o.snapshot()
    o.post_process()
    want = np.asarray(o.mean_velocity(), np.float32)
    cls = o.state.copy()
    cls[o.cell_type != 0] = 0
    sums, _, _ = capi.mean_velocity_replay(model, cls, 1200, 600)
    fluid = int((o.cell_type == 0).sum())
    got = (sums / np.float32(fluid)).astype(np.float32)
    assert got.tobytes() == want.tobytes(), (got, want)
