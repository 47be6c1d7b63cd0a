"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo processes, each owning a row
strip that is stepped by the ORACLE inside a test double of the engine; halo exchange goes through the very
lgca_b200.ring.Ring code the GPU bench uses.  Result must equal the oracle's single-lattice run
(decomposition invariance)."""
import os
import socket
import sys

import numpy as np
import pytest

from lgca_b200.ring import partition_rows, ring_neighbours


def test_partition_rows():
    assert partition_rows(64, 2, 2) == [(0, 32), (32, 32)]
    assert partition_rows(96, 3, 32) == [(0, 32), (32, 32), (64, 32)]
    assert partition_rows(128, 3, 32) == [(0, 64), (64, 32), (96, 32)]
    p = partition_rows(100, 3, 2)
    assert p[0][0] == 0 and sum(r for _, r in p) == 100 and all(r % 2 == 0 for _, r in p)
    assert all(p[i][0] + p[i][1] == p[i + 1][0] for i in range(2))
    with pytest.raises(ValueError):
        partition_rows(10, 8, 2)
    with pytest.raises(ValueError):
        partition_rows(33, 2, 2)


def test_ring_neighbours():
    assert ring_neighbours(0, 4) == (3, 1)
    assert ring_neighbours(3, 4) == (2, 0)
    assert ring_neighbours(0, 2) == (1, 1)
    assert ring_neighbours(0, 1) == (0, 0)


class HostStripEngine:
    """Test double with the engine's strip interface, stepping its rows (+ghost rows) with the oracle."""

    def __init__(self, model, dim_x, dim_y, y_begin, y_rows, halo, state, cell_type, rnd):
        from cpu_checkers import Oracle
        self.dim_x, self.dim_y, self.y0, self.own, self.halo = dim_x, dim_y, y_begin, y_rows, halo
        self.rows = y_rows + 2 * halo
        # the strip with its ghost rows is stepped as a small periodic lattice; wrap-around garbage stays
        # inside the ghost rows for <= halo steps
        self.o = Oracle(model, dims=(dim_x, self.rows), cg=1)
        gy = (np.arange(self.rows) + y_begin - halo) % dim_y
        self.o.state[:] = state.reshape(dim_y, dim_x)[gy].ravel()
        self.o.cell_type[:] = cell_type.reshape(dim_y, dim_x)[gy].ravel()
        bits = np.unpackbits(rnd, bitorder="little")[: dim_x * dim_y].reshape(dim_y, dim_x)[gy].ravel()
        self.o.rnd[:] = np.packbits(bits, bitorder="little")
        assert (y_begin - halo) % 2 == 0

    def halo_rows(self):
        return self.halo

    def halo_bytes(self, what=0):
        return self.halo * self.dim_x

    def _view(self, ptr, n):
        import ctypes
        return np.frombuffer((ctypes.c_uint8 * n).from_address(ptr), np.uint8)

    def halo_export(self, what, top_ptr, bottom_ptr):
        n, s = self.halo * self.dim_x, self.o.state.reshape(self.rows, self.dim_x)
        self._view(top_ptr, n)[:] = s[self.rows - 2 * self.halo: self.rows - self.halo].ravel()
        self._view(bottom_ptr, n)[:] = s[self.halo: 2 * self.halo].ravel()

    def halo_import(self, what, from_upper_ptr, from_lower_ptr):
        n, s = self.halo * self.dim_x, self.o.state.reshape(self.rows, self.dim_x)
        s[self.rows - self.halo:] = self._view(from_upper_ptr, n).reshape(self.halo, self.dim_x)
        s[: self.halo] = self._view(from_lower_ptr, n).reshape(self.halo, self.dim_x)

    def step(self, n):
        assert n <= self.halo
        self.o.step(n)

    # body force stages of the strip interface (lgca_b200_body_force_gather / _apply)
    def body_force_gather(self, cells):
        s = self.o.state.reshape(self.rows, self.dim_x)
        ct = self.o.cell_type.reshape(self.rows, self.dim_x)
        gy, x = cells // self.dim_x, cells % self.dim_x
        mine = (gy >= self.y0) & (gy < self.y0 + self.own)
        r = np.where(mine, gy - self.y0 + self.halo, 0)
        out = np.where(mine, s[r, x] | np.where(ct[r, x] != 0, 0x80, 0), 0x80).astype(np.uint8)
        return out

    def body_force_apply(self, cells, new_bytes):
        s = self.o.state.reshape(self.rows, self.dim_x)
        gy, x = cells // self.dim_x, cells % self.dim_x
        mine = (gy >= self.y0) & (gy < self.y0 + self.own)
        s[gy[mine] - self.y0 + self.halo, x[mine]] = new_bytes[mine]

    def own_state(self):
        return self.o.state.reshape(self.rows, self.dim_x)[self.halo: self.halo + self.own].ravel().copy()


def _worker(rank, world, port, model, dims, steps, halo, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_checkers import Oracle, OracleRng
    from lgca_b200.ring import Ring
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        o = Oracle(model, dims=dims, cg=1, rng=OracleRng(11))
        o.apply_bc("reflecting_back")
        o.init("random")
        parts = partition_rows(dims[1], world, 2)
        y0, rows = parts[rank]
        e = HostStripEngine(model, dims[0], dims[1], y0, rows, halo, o.state, o.cell_type, o.rnd)
        ring = Ring(e, rank, world)
        ring.step(steps)
        o.step(steps)
        want = o.state.reshape(dims[1], dims[0])[y0: y0 + rows].ravel()
        q.put((rank, bool(np.array_equal(e.own_state(), want)), ring.exchanges))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,model,dims,steps,halo", [(2, "FHP_III", (70, 48), 9, 2), (2, "HPP", (64, 40), 7, 4),
                                                         (3, "FHP_II", (48, 60), 8, 2)])
def test_strips_match_single_lattice(world, model, dims, steps, halo):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, model, dims, steps, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res), res
    expect = -(-steps // halo)
    assert all(r[2] == expect for r in res)


def _force_worker(rank, world, port, q):
    """Exact body force on strips through lgca_b200.ring.ring_body_force (gather -> all-reduce MIN -> host replay ->
    apply -> republish), then stepping on: the steps read the neighbours' ghost rows, so a missing republish shows."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_checkers import Oracle, OracleRng
    from lgca_b200.ring import Ring, ring_body_force
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        model, dims, halo = "FHP_III", (64, 48), 2
        o = Oracle(model, dims=dims, cg=1, bf_dir=b"x", rng=OracleRng(11))
        o.apply_bc("pipe")
        o.init("random")
        y0, rows = partition_rows(dims[1], world, 2)[rank]
        e = HostStripEngine(model, dims[0], dims[1], y0, rows, halo, o.state, o.cell_type, o.rnd)
        ring = Ring(e, rank, world)

        def combine(a):
            t = torch.from_numpy(a.copy())
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return t.numpy()

        o.rng = OracleRng(31)
        draws_rng, ok, pos = OracleRng(31), True, 0
        stream = [draws_rng.rand() for _ in range(200000)]
        for forcing in (0, 5, 300):
            used_o, rev_o = o.body_force(forcing)
            used, rev = ring_body_force(e, forcing, np.array(stream[pos:pos + used_o + 64], np.int64), o.num_cells, model, "x",
                                        combine, ring=ring)
            ok &= (used, rev) == (used_o, rev_o)
            pos += used_o
            ring.step(3)
            o.step(3)
            want = o.state.reshape(dims[1], dims[0])[y0: y0 + rows].ravel()
            ok &= bool(np.array_equal(e.own_state(), want))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_body_force_then_step_on_strips():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_force_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(r[1] for r in res), res
