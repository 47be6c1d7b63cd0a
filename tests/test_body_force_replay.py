"""Host-only part of the exact body force (lgca_b200_body_force_replay, no GPU needed): replaying batches of
draws against gathered cell bytes must reproduce the reference's sequential apply_body_force
(src/omp_lattice.cpp:254-346, via the pinned oracle) -- including duplicates inside a batch, the do-while's
"at least one draw", batch boundaries and the multi-strip combination (element-wise minimum of the strips' bytes)."""
import numpy as np
import pytest

from cpu_checkers import Oracle, OracleRng


@pytest.fixture(scope="module")
def capi():
    from lgca_b200.build import build_library
    build_library()
    from lgca_b200 import capi
    return capi


def host_gather(state, cell_type, cells, lo=0, hi=None):
    """What lgca_b200_body_force_gather returns for a strip owning cells [lo, hi)."""
    hi = state.size if hi is None else hi
    b = state[cells].copy()
    b[cell_type[cells] != 0] |= 0x80
    b[(cells < lo) | (cells >= hi)] = 0xFF
    return b


@pytest.mark.parametrize("model,bf", [("FHP_III", b"x"), ("FHP_I", b"x"), ("HPP", b"x"), ("HPP", b"y"), ("FHP_II", b"y")])
@pytest.mark.parametrize("nstrips", [1, 3])
def test_replay_matches_oracle(capi, model, bf, nstrips):
    o = Oracle(model, dims=(96, 48), cg=4, bf_dir=bf, rng=OracleRng(9))
    o.apply_bc("pipe")
    o.init("random")
    state, ct, n = o.state.copy(), o.cell_type.copy(), o.num_cells
    o.rng = OracleRng(77)
    g = OracleRng(77)
    bounds = np.linspace(0, n, nstrips + 1).astype(int)
    pending = []  # draws taken from the stream but not consumed yet (they belong to the next call)
    for forcing in (0, 1, 3, 50, 400, 2):
        used_o, rev_o = o.body_force(forcing)
        remaining, first, used_total, rev_total = forcing, True, 0, 0
        while first or remaining > 0:
            batch = 37  # small batches: many boundaries, duplicates inside and across batches
            while len(pending) < batch:
                pending.append(g.rand())
            cells = (np.array(pending[:batch], np.int64) % n).astype(np.int32)
            combined = np.full(batch, 0xFF, np.uint8)
            for s in range(nstrips):  # every strip gathers, the driver combines
                combined = np.minimum(combined, host_gather(state, ct, cells, bounds[s], bounds[s + 1]))
            used, rev, ch_cells, ch_bytes = capi.body_force_replay(model, bf, remaining if first else max(remaining, 1),
                                                                   cells, combined)
            state[ch_cells] = ch_bytes
            del pending[:used]
            used_total += used
            rev_total += rev
            remaining -= rev
            first = False
        assert (used_total, rev_total) == (used_o, rev_o), forcing
        assert np.array_equal(state, o.state), forcing
        # unconsumed draws stay queued: the generator is ahead of the oracle's by exactly len(pending)
    assert len(pending) < 37


def test_negative_forcing_runs_to_the_end_like_the_reference():
    """`reverted_particles < forcing` compares unsigned int with int in the reference (src/omp_lattice.cpp:264,346):
    a negative forcing never stops on the count, only on the draw budget."""
    from lgca_b200.capi import body_force_replay
    cells = np.arange(50, dtype=np.int32)
    state = np.full(50, 0x08, np.uint8)      # FHP: direction 3 set, direction 0 clear -> every draw reverts one particle
    used, rev, ch_cells, ch_bytes = body_force_replay("FHP_III", "x", -1, cells, state)
    assert (used, rev) == (50, 50)
    assert np.array_equal(ch_bytes, np.full(50, 0x01, np.uint8))
    used, rev, _, _ = body_force_replay("FHP_III", "x", 3, cells, state)
    assert (used, rev) == (3, 3)
