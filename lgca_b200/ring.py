"""Row-strip domain decomposition over the GPUs of one box: one process (rank) per GPU, one strip per rank.

The reference has no decomposition at all (one shared-memory lattice); its torus is always
periodic (src/omp_lattice.cpp:150-176), so the strips form a periodic ring in y: rank r's upper neighbour is
(r+1) % world, its lower neighbour (r-1) % world.  After every block of at most `halo` fused steps each rank
sends its top rows up and its bottom rows down and receives the matching ghost rows.  The transfer itself is
torch.distributed point-to-point (NCCL over NVLink on GPUs, gloo in the CPU tests); pack/unpack of the rows is
done by the engine on its own stream and everything is stream-ordered (no host synchronisation per block).

This module is host-side plumbing only -- it never touches lattice data itself.
"""
import contextlib


def partition_rows(dim_y, world, multiple=2):
    """Split dim_y rows into `world` contiguous strips whose heights are multiples of `multiple`
    (2 keeps the hexagonal row parity local == global; 2*cg_radius keeps coarse cells strip-local).
    Returns [(y_begin, y_rows)] with the remainder spread over the first strips."""
    if dim_y % multiple:
        raise ValueError("dim_y must be a multiple of %d" % multiple)
    units = dim_y // multiple
    if units < world:
        raise ValueError("lattice too small for %d strips" % world)
    base, extra = divmod(units, world)
    out, y = [], 0
    for r in range(world):
        rows = (base + (1 if r < extra else 0)) * multiple
        out.append((y, rows))
        y += rows
    return out


def ring_neighbours(rank, world):
    """(lower, upper) ranks of the periodic ring."""
    return (rank - 1) % world, (rank + 1) % world


class Ring:
    """Halo exchange driver for one strip.

    `engine` needs: halo_rows(), halo_bytes(what), halo_export(what, top_ptr, bottom_ptr),
    halo_import(what, from_upper_ptr, from_lower_ptr), step(n), compute_stream() (CUDA engines only).
    """

    STATE, MASKS = 0, 1

    def __init__(self, engine, rank, world, device=None, group=None, native=False):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.e, self.rank, self.world, self.group = engine, rank, world, group
        self.lower, self.upper = ring_neighbours(rank, world)
        self.halo = engine.halo_rows()
        # ghost rows are refreshed after every launch of the fused-step kernel: block = k_fuse steps
        self.block = getattr(engine, "steps_per_exchange", lambda: self.halo)()
        self.device = device if device is not None else torch.device("cpu")
        self._bufs = {}
        self._stream = None
        if self.device.type == "cuda":
            # NCCL work is ordered against the engine's own compute stream
            self._stream = torch.cuda.ExternalStream(engine.compute_stream(), device=self.device)
        self.exchanges = 0
        # native mode: ghost rows are stored straight into the neighbours' memory by the library (CUDA IPC /
        # peer stores + device-side epoch flags); torch.distributed only carries the descriptors once
        self.native = bool(native) and self.world > 1 and self.halo > 0
        if self.native:
            mine = engine.ring_export()
            allb = [None] * world
            dist.all_gather_object(allb, mine, group=group)
            engine.ring_connect(allb[self.lower], allb[self.upper])
            dist.barrier(group=group)

    def start(self):
        """Publish the current edge rows (after upload / init, before the first step)."""
        if self.native:
            self.e.ring_start()
        else:
            self.exchange(self.STATE)

    def republish(self):
        """Publish the edge rows again after the live state was changed in place (body force, upload, initialiser).
        Collective: every rank calls it at the same point."""
        if self.world == 1 or self.halo == 0:
            return
        if self.native:
            self.e.ring_republish()
        else:
            self.exchange(self.STATE)

    def _buffers(self, what):
        if what not in self._bufs:
            n = self.e.halo_bytes(what)
            mk = lambda: self.torch.empty(n, dtype=self.torch.uint8, device=self.device)
            self._bufs[what] = dict(top=mk(), bottom=mk(), from_upper=mk(), from_lower=mk())
        return self._bufs[what]

    def _ctx(self):
        return self.torch.cuda.stream(self._stream) if self._stream is not None else contextlib.nullcontext()

    def exchange(self, what=0):
        """Send own top rows up / bottom rows down, receive the ghost rows, install them."""
        if self.world == 1 or self.halo == 0:
            return
        b, dist = self._buffers(what), self.dist
        with self._ctx():
            self.e.halo_export(what, b["top"].data_ptr(), b["bottom"].data_ptr())
            # order matters when lower == upper (world == 2): the peer's first send (its top rows) must meet
            # our first recv (from_lower), its second (bottom rows) our second (from_upper)
            ops = [dist.P2POp(dist.isend, b["top"], self.upper, self.group),
                   dist.P2POp(dist.isend, b["bottom"], self.lower, self.group),
                   dist.P2POp(dist.irecv, b["from_lower"], self.lower, self.group),
                   dist.P2POp(dist.irecv, b["from_upper"], self.upper, self.group)]
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            self.e.halo_import(what, b["from_upper"].data_ptr(), b["from_lower"].data_ptr())
        self.exchanges += 1

    def step(self, n):
        """n lattice updates; ghost rows are refreshed after every block of <= k_fuse steps."""
        if self.world == 1 or self.halo == 0:
            self.e.step(n)
            return
        if self.native:
            self.e.ring_step(n)
            return
        while n > 0:
            k = min(self.block, n)
            self.e.step(k)
            self.exchange(self.STATE)
            n -= k


def ring_body_force(engine, forcing, draws, num_cells, model, bf_dir, combine, batch=None, ring=None):
    """Exact body force on a lattice that is split into strips (reference: OMP_Lattice::apply_body_force,
    src/omp_lattice.cpp:254-346).  Every rank calls this with the SAME draws (the caller's rand() values in stream
    order): each strip gathers the bytes of the drawn cells (0xFF outside the strip), `combine(uint8 array)` returns
    the element-wise minimum over all strips (all-reduce MIN), the batch is replayed in draw order on the host
    (identically on every rank) and each strip applies the changed cells it owns.  The apply may change edge rows
    the neighbours hold copies of: pass `ring` (a Ring) to have them republished, or call ring.republish() yourself
    before the next step.
    Returns (draws consumed, particles reverted)."""
    import numpy as np
    from .capi import body_force_replay
    draws = np.asarray(draws, np.int64)
    pos, remaining, first, reverted, changed = 0, int(forcing), True, 0, False
    while pos < draws.size and (first or remaining > 0):
        n = batch or max(4096, min(1 << 20, remaining * 12))
        cells = (draws[pos:pos + n] % num_cells).astype(np.int32)
        combined = combine(engine.body_force_gather(cells))
        used, rev, ch_cells, ch_bytes = body_force_replay(model, bf_dir, remaining if first else max(remaining, 1), cells, combined)
        if ch_cells.size:
            engine.body_force_apply(ch_cells, ch_bytes)
            changed = True
        pos += used
        remaining -= rev
        reverted += rev
        first = False
    if changed and ring is not None:
        ring.republish()
    return pos, reverted
