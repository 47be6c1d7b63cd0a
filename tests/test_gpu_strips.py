"""Row strips on the GPU (run with -m gpu): several strip handles on ONE device, ghost rows moved with the
same Ring-style export/import calls the multi-GPU bench uses (here by plain device pointers inside one
process).  The union of the strips must equal the oracle's single-lattice run bit-exactly
(decomposition invariance), including walls crossing strip borders and the odd/even hex rows."""
import numpy as np
import pytest

from cpu_checkers import Oracle, OracleRng

pytestmark = pytest.mark.gpu


class LocalRing:
    """All strips in one process: exports of every strip first, then imports from the ring neighbours."""

    def __init__(self, engines):
        import torch
        self.t, self.e = torch, engines
        self.n = len(engines)
        self.buf = {}

    def _b(self, what):
        if what not in self.buf:
            mk = lambda e: self.t.empty(e.halo_bytes(what), dtype=self.t.uint8, device="cuda")
            self.buf[what] = [(mk(e), mk(e)) for e in self.e]
        return self.buf[what]

    def exchange(self, what=0):
        b = self._b(what)
        for e, (top, bottom) in zip(self.e, b):
            e.halo_export(what, top.data_ptr(), bottom.data_ptr())
        for e in self.e:
            e.sync()
        for r, e in enumerate(self.e):
            upper, lower = (r + 1) % self.n, (r - 1) % self.n
            # from the upper neighbour: its bottom rows; from the lower neighbour: its top rows
            e.halo_import(what, b[upper][1].data_ptr(), b[lower][0].data_ptr())
        for e in self.e:
            e.sync()


@pytest.mark.parametrize("model,dims,bc,nstrips,k", [
    ("FHP_III", (256, 96), "karman", 2, 2),
    ("FHP_III", (1000, 120), "periodic", 3, 4),
    ("FHP_II", (512, 128), "reflecting_back", 4, 3),
    ("FHP_I", (300, 64), "reflecting_forward", 2, 1),
    ("HPP", (640, 90), "reflecting_forward", 3, 4),
    # dim_x < 64: only the generic kernel applies, which writes owned rows only -> ONE step per exchange
    ("FHP_III", (48, 64), "pipe", 2, 4),
    ("HPP", (33, 30), "reflecting_back", 3, 2),
])
def test_strips_equal_single_lattice(model, dims, bc, nstrips, k):
    import lgca_b200
    from lgca_b200.ring import partition_rows
    o = Oracle(model, dims=dims, cg=1, rng=OracleRng(5))
    o.apply_bc(bc)
    o.init("random")
    parts = partition_rows(dims[1], nstrips, 2)
    engines = []
    for y0, rows in parts:
        e = lgca_b200.Engine(model, dims[0], dims[1], k_fuse=k, y_begin=y0, y_rows=rows)
        sl = slice(y0 * dims[0], (y0 + rows) * dims[0])
        e.upload(o.state[sl], o.cell_type[sl], o.rnd)
        engines.append(e)
    # every strip runs with the union of the wall kinds; ghost rows get the neighbours' masks once
    flags = [e.wall_flags() for e in engines]
    for e in engines:
        e.set_wall_flags(any(f[0] for f in flags), any(f[1] for f in flags))
    ring = LocalRing(engines)
    ring.exchange(1)
    ring.exchange(0)
    halo = engines[0].halo_rows()
    assert halo >= k and halo % 2 == 0
    block = engines[0].steps_per_exchange()
    assert block == (k if dims[0] >= 64 else 1)
    done = 0
    for n in (1, k, 2 * k + 1, 7):
        left = n
        while left > 0:
            b = min(left, block)
            for e in engines:
                e.step(b)
            ring.exchange(0)
            left -= b
        o.step(n)
        done += n
        got = np.concatenate([e.download() for e in engines])
        assert np.array_equal(got, o.state), "after %d steps" % done
    assert sum(e.count_particles() for e in engines) == o.n_particles()
    with pytest.raises(lgca_b200.LgcaError):
        engines[0].step(block + 1)  # a strip may not run past its ghost rows
    for e in engines:
        e.close()


def make_strips(o, nstrips, k, cg=0, unit=2, **kw):
    import lgca_b200
    from lgca_b200.ring import partition_rows
    parts = partition_rows(o.dim_y, nstrips, unit)
    engines = []
    for y0, rows in parts:
        e = lgca_b200.Engine(o.model, o.dim_x, o.dim_y, cg_radius=cg, k_fuse=k, y_begin=y0, y_rows=rows, **kw)
        sl = slice(y0 * o.dim_x, (y0 + rows) * o.dim_x)
        e.upload(o.state[sl], o.cell_type[sl], o.rnd)
        engines.append(e)
    # every strip runs with the union of the wall kinds; ghost rows get the neighbours' masks once
    flags = [e.wall_flags() for e in engines]
    for e in engines:
        e.set_wall_flags(any(f[0] for f in flags), any(f[1] for f in flags))
    LocalRing(engines).exchange(1)  # static masks of the ghost rows: once, through the packed-buffer path
    return engines, parts


def connect_native(engines):
    desc = [e.ring_export() for e in engines]
    n = len(engines)
    for r, e in enumerate(engines):
        e.ring_connect(desc[(r - 1) % n], desc[(r + 1) % n])
    for e in engines:
        e.ring_start()


@pytest.mark.parametrize("model,dims,bc,nstrips,k", [
    ("FHP_III", (256, 96), "karman", 2, 2),
    ("FHP_III", (1024, 120), "periodic", 3, 4),
    ("FHP_II", (512, 128), "reflecting_back", 4, 3),
    ("HPP", (640, 96), "reflecting_forward", 3, 4),
    ("FHP_III", (1024, 100), "pipe", 3, 4),        # strips of unequal height (34 / 34 / 32 rows)
    ("FHP_I", (40, 48), "pipe", 2, 3),             # dim_x < 64: generic kernel, one step per exchange, wait kernel
])
def test_native_ring_equals_single_lattice(model, dims, bc, nstrips, k):
    """Same invariance through the native ring: peer stores into the neighbours' ghost rows + epoch flags,
    everything enqueued by lgca_b200_ring_step (strips of one process share plain device pointers).  Snapshots are
    taken in between: the three plane sets of every strip rotate in lockstep (zero-copy snapshot)."""
    o = Oracle(model, dims=dims, cg=1, rng=OracleRng(6))
    o.apply_bc(bc)
    o.init("random")
    engines, parts = make_strips(o, nstrips, k)
    connect_native(engines)
    done = 0
    for i, steps in enumerate((1, k, 3 * k + 1, 10, 2)):
        for e in engines:
            e.ring_step(steps)
        o.step(steps)
        done += steps
        if i % 2 == 1:
            for e in engines:
                e.snapshot()
        got = np.concatenate([e.download() for e in engines])
        assert np.array_equal(got, o.state), "after %d steps" % done
    for e in engines:
        e.close()


@pytest.mark.parametrize("native", [False, True], ids=["packed", "native"])
def test_strip_snapshots_and_coarse_means(native):
    """Post-processing on strips: the coarse means of a strip's top coarse row reach one row into the upper neighbour
    (reference window, src/omp_lattice.cpp:423-436).  Snapshot on every strip, step on, then post-process: the union of
    the strips' fields equals the oracle's, and the snapshot is insulated from the stepping (zero-copy rotation)."""
    cg = 4
    o = Oracle("FHP_III", dims=(256, 96), cg=cg, rng=OracleRng(11))
    o.apply_bc("karman")
    o.init("random")
    engines, parts = make_strips(o, 3, 4, cg=cg, unit=2 * cg)
    ring = LocalRing(engines)
    if native:
        connect_native(engines)
    else:
        ring.exchange(0)

    def advance(n):
        block = engines[0].steps_per_exchange()
        while n > 0:
            b = min(n, block)
            for e in engines:
                e.ring_step(b) if native else e.step(b)
            if not native:
                ring.exchange(0)
            n -= b

    for rounds in range(3):
        advance(7)
        o.step(7)
        for e in engines:
            e.snapshot()
        o.snapshot()
        o.post_process()
        advance(5)          # the snapshot must not see these
        o.step(5)
        fields = [e.post_process(cell=True, mean=True, exact=True) for e in engines]
        for name in ("cell_density", "cell_momentum", "mean_density", "mean_momentum"):
            got = np.concatenate([f[name] for f in fields])
            assert np.array_equal(got, getattr(o, name)), (name, rounds)
        got = np.concatenate([e.download() for e in engines])
        assert np.array_equal(got, o.state)
    for e in engines:
        e.close()


@pytest.mark.parametrize("native", [False, True], ids=["packed", "native"])
def test_body_force_on_strips(native):
    """Exact body force across strips: gather on every strip, combine (minimum), ordered host replay, apply --
    then the changed edge rows are published again and the lattice steps on: the next steps read the neighbours'
    ghost rows, so a stale copy would show up here (force -> step -> compare)."""
    from lgca_b200.capi import body_force_replay
    model, dims, nstrips = "FHP_III", (256, 96), 3
    o = Oracle(model, dims=dims, cg=1, bf_dir=b"x", rng=OracleRng(8))
    o.apply_bc("karman")
    o.init("random")
    engines, parts = make_strips(o, nstrips, 2, bf_dir="x")
    ring = LocalRing(engines)
    if native:
        connect_native(engines)
    else:
        ring.exchange(0)
    o.rng = OracleRng(123)
    g = OracleRng(123)
    pending = []
    n = o.num_cells
    for forcing in (0, 3, 200, 1500):
        for e in engines:
            e.snapshot()           # the canonical schedule snapshots before the force: copy-on-write on every strip
        used_o, rev_o = o.body_force(forcing)
        remaining, first, used_t, rev_t, changed = forcing, True, 0, 0, False
        while first or remaining > 0:
            want = max(512, remaining * 6)
            while len(pending) < want:
                pending.append(g.rand())
            cells = (np.array(pending[:want], np.int64) % n).astype(np.int32)
            combined = np.full(want, 0xFF, np.uint8)
            for e in engines:
                combined = np.minimum(combined, e.body_force_gather(cells))
            used, rev, cc, cb = body_force_replay(model, "x", remaining if first else max(remaining, 1), cells, combined)
            if cc.size:
                for e in engines:
                    e.body_force_apply(cc, cb)
                changed = True
            del pending[:used]
            used_t += used
            rev_t += rev
            remaining -= rev
            first = False
        assert (used_t, rev_t) == (used_o, rev_o)
        got = np.concatenate([e.download() for e in engines])
        assert np.array_equal(got, o.state), forcing
        if changed:                # mandatory after an in-place write: publish the edge rows again
            if native:
                for e in engines:
                    e.ring_republish()
            else:
                ring.exchange(0)
        for steps in (2, 1):
            for e in engines:
                e.ring_step(steps) if native else e.step(steps)
            if not native:
                ring.exchange(0)
            o.step(steps)
        got = np.concatenate([e.download() for e in engines])
        assert np.array_equal(got, o.state), "stepping after forcing %d" % forcing
    for e in engines:
        e.close()


def test_strip_geometry_is_validated():
    """Strips shorter than their halo, FHP strips on odd rows: rejected at create (LGCA_B200_EINVAL)."""
    import lgca_b200
    with pytest.raises(lgca_b200.LgcaError):
        lgca_b200.Engine("FHP_III", 256, 96, k_fuse=6, y_begin=0, y_rows=4)      # 4 rows < halo 6
    with pytest.raises(lgca_b200.LgcaError):
        lgca_b200.Engine("FHP_III", 256, 96, k_fuse=2, y_begin=3, y_rows=32)     # odd first row on a hex lattice
    with pytest.raises(lgca_b200.LgcaError):
        lgca_b200.Engine("FHP_II", 256, 96, k_fuse=2, y_begin=32, y_rows=33)     # odd height
    e = lgca_b200.Engine("HPP", 256, 96, k_fuse=2, y_begin=3, y_rows=33)         # HPP has no row parity
    assert e.steps_per_exchange() == 1                                           # ... but then only the generic kernel applies
    e.close()
