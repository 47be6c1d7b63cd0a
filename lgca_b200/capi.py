"""ctypes binding of include/lgca_b200.h.  Fails loudly when the CUDA library is missing: there is no
CPU path in this package."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODELS = {"HPP": 0, "FHP_I": 1, "FHP_II": 2, "FHP_III": 3}
NUM_DIR = {0: 4, 1: 6, 2: 7, 3: 7}

FLAG_NO_CELL_FIELDS = 1 << 0
FLAG_SIMPLE_KERNEL = 1 << 1
FLAG_NO_RESIDENT = 1 << 2
FLAG_FORCE_RESIDENT = 1 << 3
FLAG_HOST_BODY_FORCE = 1 << 4
FLAG_RESIDENT_DYNAMIC = 1 << 5
FLAG_NO_CHAIN = 1 << 6
FLAG_FORCE_CHAIN = 1 << 7

# every symbol include/lgca_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "lgca_b200_create", "lgca_b200_destroy", "lgca_b200_last_error", "lgca_b200_version",
    "lgca_b200_device_count", "lgca_b200_host_alloc", "lgca_b200_host_free", "lgca_b200_upload",
    "lgca_b200_download", "lgca_b200_step", "lgca_b200_snapshot", "lgca_b200_post_process",
    "lgca_b200_mean_velocity", "lgca_b200_mean_velocity_exact", "lgca_b200_mean_velocity_replay", "lgca_b200_mean_velocity_stats",
    "lgca_b200_group_mean_velocity_exact", "lgca_b200_body_force", "lgca_b200_body_force_gather", "lgca_b200_body_force_replay",
    "lgca_b200_body_force_apply", "lgca_b200_count_particles",
    "lgca_b200_init_random_device", "lgca_b200_apply_bc_device", "lgca_b200_sync",
    "lgca_b200_compute_stream", "lgca_b200_timed_steps", "lgca_b200_timed_kernel", "lgca_b200_launch_count", "lgca_b200_get_info",
    "lgca_b200_halo_rows", "lgca_b200_halo_bytes", "lgca_b200_halo_export", "lgca_b200_halo_import",
    "lgca_b200_get_wall_flags", "lgca_b200_set_wall_flags",
    "lgca_b200_ring_descriptor_bytes", "lgca_b200_ring_export", "lgca_b200_ring_connect", "lgca_b200_ring_start",
    "lgca_b200_ring_step", "lgca_b200_ring_disconnect", "lgca_b200_ring_republish", "lgca_b200_steps_per_exchange",
    "lgca_b200_group_create", "lgca_b200_group_destroy", "lgca_b200_group_size", "lgca_b200_group_strip",
    "lgca_b200_group_upload", "lgca_b200_group_download", "lgca_b200_group_step", "lgca_b200_group_snapshot",
    "lgca_b200_group_post_process", "lgca_b200_group_mean_velocity", "lgca_b200_group_body_force",
    "lgca_b200_group_count_particles", "lgca_b200_group_init_random_device", "lgca_b200_group_apply_bc_device",
    "lgca_b200_group_sync", "lgca_b200_group_timed_steps", "lgca_b200_group_launch_count", "lgca_b200_group_get_info",
]


class LgcaError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("model", C.c_int32), ("dim_x", C.c_uint32), ("dim_y", C.c_uint32), ("cg_radius", C.c_uint32),
                ("bf_dir", C.c_int32), ("device", C.c_int32), ("k_fuse", C.c_int32), ("y_begin", C.c_uint32),
                ("y_rows", C.c_uint32), ("flags", C.c_uint32)]


class Info(C.Structure):
    _fields_ = [("dim_x", C.c_uint32), ("dim_y", C.c_uint32), ("y_begin", C.c_uint32), ("y_rows", C.c_uint32),
                ("words_per_row", C.c_uint32), ("num_planes", C.c_uint32), ("has_no_slip", C.c_uint32),
                ("has_slip", C.c_uint32), ("k_fuse", C.c_int32), ("bytes_per_site_step_x8", C.c_uint64),
                ("device_bytes", C.c_uint64)]


def library_path():
    # LGCA_B200_LIB: alternative build of the same library (A/B experiments); never a different backend
    return os.environ.get("LGCA_B200_LIB") or os.path.join(HERE, "liblgca_b200.so")


_LIB = None


def load_library():
    """Load liblgca_b200.so; raise (never fall back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise LgcaError("liblgca_b200.so is missing (%s): build it with `python -m lgca_b200.build` "
                        "(__graft_entry__.build()); lgca_b200 has no CPU fallback" % path)
    L = C.CDLL(path)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    L.lgca_b200_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.lgca_b200_destroy.argtypes = [vp]
    L.lgca_b200_last_error.restype = C.c_char_p
    L.lgca_b200_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.lgca_b200_host_free.argtypes = [vp]
    L.lgca_b200_upload.argtypes = [vp, vp, vp, vp]
    L.lgca_b200_download.argtypes = [vp, vp]
    L.lgca_b200_step.argtypes = [vp, i32]
    L.lgca_b200_snapshot.argtypes = [vp]
    L.lgca_b200_post_process.argtypes = [vp, vp, vp, vp, vp, i32]
    L.lgca_b200_mean_velocity.argtypes = [vp, vp]
    L.lgca_b200_mean_velocity_exact.argtypes = [vp, vp, C.POINTER(u64)]
    L.lgca_b200_mean_velocity_replay.argtypes = [i32, vp, C.c_uint32, C.c_uint32, vp, C.POINTER(u64), C.POINTER(u64)]
    L.lgca_b200_group_mean_velocity_exact.argtypes = [vp, vp]
    L.lgca_b200_mean_velocity_stats.argtypes = [vp, vp]
    L.lgca_b200_body_force.argtypes = [vp, i32, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]
    L.lgca_b200_body_force_gather.argtypes = [vp, vp, C.c_size_t, vp]
    L.lgca_b200_body_force_replay.argtypes = [i32, i32, i32, vp, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32),
                                              vp, vp, C.POINTER(C.c_size_t)]
    L.lgca_b200_body_force_apply.argtypes = [vp, vp, vp, C.c_size_t]
    L.lgca_b200_count_particles.argtypes = [vp, C.POINTER(u64)]
    L.lgca_b200_init_random_device.argtypes = [vp, u64]
    L.lgca_b200_apply_bc_device.argtypes = [vp, C.c_char_p]
    L.lgca_b200_sync.argtypes = [vp]
    L.lgca_b200_compute_stream.argtypes = [vp]
    L.lgca_b200_compute_stream.restype = vp
    L.lgca_b200_timed_steps.argtypes = [vp, i32, C.POINTER(C.c_float)]
    L.lgca_b200_timed_kernel.argtypes = [vp, i32, C.POINTER(C.c_float)]
    L.lgca_b200_launch_count.argtypes = [vp, C.POINTER(u64)]
    L.lgca_b200_get_info.argtypes = [vp, C.POINTER(Info)]
    L.lgca_b200_halo_rows.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.lgca_b200_halo_bytes.argtypes = [vp, i32, C.POINTER(C.c_size_t)]
    L.lgca_b200_halo_export.argtypes = [vp, i32, vp, vp]
    L.lgca_b200_halo_import.argtypes = [vp, i32, vp, vp]
    L.lgca_b200_get_wall_flags.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.lgca_b200_set_wall_flags.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.lgca_b200_ring_descriptor_bytes.argtypes = [C.POINTER(C.c_size_t)]
    L.lgca_b200_ring_export.argtypes = [vp, vp, C.c_size_t]
    L.lgca_b200_ring_connect.argtypes = [vp, vp, vp]
    L.lgca_b200_ring_start.argtypes = [vp]
    L.lgca_b200_ring_step.argtypes = [vp, i32]
    L.lgca_b200_ring_disconnect.argtypes = [vp]
    L.lgca_b200_ring_republish.argtypes = [vp]
    L.lgca_b200_steps_per_exchange.argtypes = [vp, C.POINTER(C.c_int)]
    L.lgca_b200_group_create.argtypes = [C.POINTER(Config), i32, vp, C.POINTER(vp)]
    L.lgca_b200_group_destroy.argtypes = [vp]
    L.lgca_b200_group_size.argtypes = [vp, C.POINTER(C.c_int)]
    L.lgca_b200_group_strip.argtypes = [vp, i32, C.POINTER(vp)]
    L.lgca_b200_group_upload.argtypes = [vp, vp, vp, vp]
    L.lgca_b200_group_download.argtypes = [vp, vp]
    L.lgca_b200_group_step.argtypes = [vp, i32]
    L.lgca_b200_group_snapshot.argtypes = [vp]
    L.lgca_b200_group_post_process.argtypes = [vp, vp, vp, vp, vp, i32]
    L.lgca_b200_group_mean_velocity.argtypes = [vp, vp]
    L.lgca_b200_group_body_force.argtypes = [vp, i32, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]
    L.lgca_b200_group_count_particles.argtypes = [vp, C.POINTER(u64)]
    L.lgca_b200_group_init_random_device.argtypes = [vp, u64]
    L.lgca_b200_group_apply_bc_device.argtypes = [vp, C.c_char_p]
    L.lgca_b200_group_sync.argtypes = [vp]
    L.lgca_b200_group_timed_steps.argtypes = [vp, i32, C.POINTER(C.c_float)]
    L.lgca_b200_group_launch_count.argtypes = [vp, C.POINTER(u64)]
    L.lgca_b200_group_get_info.argtypes = [vp, C.POINTER(Info)]
    _LIB = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def mean_velocity_replay(model, class_bytes, dim_x, rows, sums=(0.0, 0.0)):
    """Host-only order-exact mean-velocity sums over a row-major array of class bytes (lgca_b200_mean_velocity_replay).
    Returns (sums float32[2], segments_fast, segments_walked)."""
    L = load_library()
    model = MODELS[model] if isinstance(model, str) else int(model)
    cls = np.ascontiguousarray(class_bytes, np.uint8)
    assert cls.size == dim_x * rows
    out = np.array(sums, np.float32)
    fast, walked = C.c_uint64(0), C.c_uint64(0)
    rc = L.lgca_b200_mean_velocity_replay(model, _ptr(cls), int(dim_x), int(rows), _ptr(out), C.byref(fast), C.byref(walked))
    if rc != 0:
        raise LgcaError(L.lgca_b200_last_error().decode())
    return out, int(fast.value), int(walked.value)


def body_force_replay(model, bf_dir, forcing, cells, cell_bytes):
    """Host-only ordered replay of one batch of body-force draws (see lgca_b200_body_force_replay)."""
    L = load_library()
    model = MODELS[model] if isinstance(model, str) else int(model)
    if isinstance(bf_dir, (bytes, str)):
        bf_dir = ord(bf_dir) if len(bf_dir) and bf_dir not in (b"\0", "\0") else 0
    cells = np.ascontiguousarray(cells, np.int32)
    cell_bytes = np.ascontiguousarray(cell_bytes, np.uint8)
    n = cells.size
    ch_cells, ch_bytes = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.uint8)
    used, rev, nch = C.c_size_t(0), C.c_uint32(0), C.c_size_t(0)
    rc = L.lgca_b200_body_force_replay(model, bf_dir, int(forcing), _ptr(cells), _ptr(cell_bytes), n, C.byref(used), C.byref(rev),
                                       _ptr(ch_cells), _ptr(ch_bytes), C.byref(nch))
    if rc != 0:
        raise LgcaError(L.lgca_b200_last_error().decode())
    return int(used.value), int(rev.value), ch_cells[: nch.value].copy(), ch_bytes[: nch.value].copy()


class PinnedArray:
    """numpy view over cudaHostAlloc'ed memory (host mirrors of the reference-layout arrays)."""

    def __init__(self, shape, dtype):
        L = load_library()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        rc = L.lgca_b200_host_alloc(self.nbytes, C.byref(p))
        if rc != 0:
            raise LgcaError(L.lgca_b200_last_error().decode())
        self._p = p
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            load_library().lgca_b200_host_free(self._p)
            self._p = None


class Engine:
    """One lattice (or one row strip of it) on one GPU -- a 1:1 wrapper of the C-ABI handle."""

    def __init__(self, model, dim_x, dim_y, cg_radius=0, bf_dir=0, device=0, k_fuse=0, y_begin=0, y_rows=0, flags=0):
        self.L = load_library()
        self.model = MODELS[model] if isinstance(model, str) else int(model)
        if isinstance(bf_dir, (bytes, str)):
            bf_dir = ord(bf_dir) if bf_dir not in (b"\0", "\0", "", b"") else 0
        cfg = Config(self.model, dim_x, dim_y, cg_radius, bf_dir, device, k_fuse, y_begin, y_rows, flags)
        h = C.c_void_p()
        self._check(self.L.lgca_b200_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self._flags = flags
        self.dim_x, self.dim_y = dim_x, dim_y
        self.cg = cg_radius
        self.num_dir = NUM_DIR[self.model]
        i = self.info()
        self.y_begin, self.y_rows = i.y_begin, i.y_rows
        self.cells = self.dim_x * self.y_rows

    def _check(self, rc):
        if rc != 0:
            raise LgcaError("lgca_b200 error %d: %s" % (rc, self.L.lgca_b200_last_error().decode()))

    def close(self):
        if getattr(self, "h", None):
            self.L.lgca_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- data movement (reference layouts) -----------------------------------------------------
    def upload(self, state=None, cell_type=None, rnd_bits=None):
        if state is not None:
            state = np.ascontiguousarray(state, np.uint8)
            assert state.size == self.cells
        if cell_type is not None:
            cell_type = np.ascontiguousarray(cell_type, np.int32)
            assert cell_type.size == self.cells
        if rnd_bits is not None:
            rnd_bits = np.ascontiguousarray(rnd_bits, np.uint8)
            assert rnd_bits.size >= (self.dim_x * self.dim_y + 7) // 8
        self._check(self.L.lgca_b200_upload(self.h, _ptr(state), _ptr(cell_type), _ptr(rnd_bits)))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.cells, np.uint8)
        self._check(self.L.lgca_b200_download(self.h, _ptr(out)))
        return out

    # --- the hot path ------------------------------------------------------------------------------
    def step(self, n=1):
        self._check(self.L.lgca_b200_step(self.h, int(n)))

    def timed_steps(self, n):
        ms = C.c_float(0)
        self._check(self.L.lgca_b200_timed_steps(self.h, int(n), C.byref(ms)))
        return float(ms.value)

    def timed_kernel(self, launches):
        ms = C.c_float(0)
        self._check(self.L.lgca_b200_timed_kernel(self.h, int(launches), C.byref(ms)))
        return float(ms.value)

    def sync(self):
        self._check(self.L.lgca_b200_sync(self.h))

    def snapshot(self):
        self._check(self.L.lgca_b200_snapshot(self.h))

    def post_process(self, cell=True, mean=True, exact=True, out=None):
        out = {} if out is None else out
        n = self.cells
        if cell:
            out.setdefault("cell_density", np.empty(n, np.float32))
            out.setdefault("cell_momentum", np.empty(2 * n, np.float32))
        if mean:
            nc = (self.dim_x // (2 * self.cg)) * (self.y_rows // (2 * self.cg))
            out.setdefault("mean_density", np.empty(nc, np.float32))
            out.setdefault("mean_momentum", np.empty(2 * nc, np.float32))
        self._check(self.L.lgca_b200_post_process(self.h, _ptr(out.get("cell_density")), _ptr(out.get("cell_momentum")),
                                                  _ptr(out.get("mean_density")), _ptr(out.get("mean_momentum")),
                                                  1 if exact else 0))
        return out

    def mean_velocity_exact(self, sums=(0.0, 0.0), fluid_cells=0):
        """Continues the reference's sequential float32 sums over this handle's rows; returns (sums, fluid_cells)."""
        out = np.array(sums, np.float32)
        cnt = C.c_uint64(int(fluid_cells))
        self._check(self.L.lgca_b200_mean_velocity_exact(self.h, _ptr(out), C.byref(cnt)))
        return out, int(cnt.value)

    def mean_velocity_stats(self):
        """{segments_fast, segments_walked, device_ns, walk_ns} accumulated over the mean_velocity_exact calls so far."""
        out = np.zeros(4, np.uint64)
        self._check(self.L.lgca_b200_mean_velocity_stats(self.h, _ptr(out)))
        return dict(zip(("segments_fast", "segments_walked", "device_ns", "walk_ns"), (int(v) for v in out)))

    def mean_velocity(self):
        out = np.zeros(2, np.float32)
        self._check(self.L.lgca_b200_mean_velocity(self.h, _ptr(out)))
        return out

    def body_force(self, forcing, draws):
        draws = np.ascontiguousarray(draws, np.int32)
        used, rev = C.c_size_t(0), C.c_uint32(0)
        self._check(self.L.lgca_b200_body_force(self.h, int(forcing), _ptr(draws), draws.size, C.byref(used), C.byref(rev)))
        return int(used.value), int(rev.value)

    def body_force_gather(self, cells):
        cells = np.ascontiguousarray(cells, np.int32)
        out = np.empty(cells.size, np.uint8)
        self._check(self.L.lgca_b200_body_force_gather(self.h, _ptr(cells), cells.size, _ptr(out)))
        return out

    def body_force_apply(self, cells, new_bytes):
        cells = np.ascontiguousarray(cells, np.int32)
        new_bytes = np.ascontiguousarray(new_bytes, np.uint8)
        self._check(self.L.lgca_b200_body_force_apply(self.h, _ptr(cells), _ptr(new_bytes), cells.size))

    def count_particles(self):
        v = C.c_uint64(0)
        self._check(self.L.lgca_b200_count_particles(self.h, C.byref(v)))
        return int(v.value)

    def init_random_device(self, seed=1):
        self._check(self.L.lgca_b200_init_random_device(self.h, int(seed)))

    def apply_bc_device(self, name):
        self._check(self.L.lgca_b200_apply_bc_device(self.h, name.encode()))

    def launch_count(self):
        v = C.c_uint64(0)
        self._check(self.L.lgca_b200_launch_count(self.h, C.byref(v)))
        return int(v.value)

    def info(self):
        i = Info()
        self._check(self.L.lgca_b200_get_info(self.h, C.byref(i)))
        return i

    def compute_stream(self):
        return self.L.lgca_b200_compute_stream(self.h)

    # --- multi-GPU halo plumbing (device pointers; async on the compute stream) -------------------
    HALO_STATE, HALO_MASKS = 0, 1

    def halo_rows(self):
        v = C.c_uint32(0)
        self._check(self.L.lgca_b200_halo_rows(self.h, C.byref(v)))
        return int(v.value)

    def steps_per_exchange(self):
        """Steps a strip may advance between two halo exchanges (one launch of the kernel that will actually run)."""
        v = C.c_int(0)
        self._check(self.L.lgca_b200_steps_per_exchange(self.h, C.byref(v)))
        return int(v.value)

    def halo_bytes(self, what=0):
        v = C.c_size_t(0)
        self._check(self.L.lgca_b200_halo_bytes(self.h, what, C.byref(v)))
        return int(v.value)

    def halo_export(self, what, dev_top, dev_bottom):
        self._check(self.L.lgca_b200_halo_export(self.h, what, C.c_void_p(dev_top), C.c_void_p(dev_bottom)))

    def halo_import(self, what, dev_from_upper, dev_from_lower):
        self._check(self.L.lgca_b200_halo_import(self.h, what, C.c_void_p(dev_from_upper), C.c_void_p(dev_from_lower)))

    # --- native ring (peer stores; descriptors are opaque bytes moved between ranks by the caller) --
    def ring_export(self):
        n = C.c_size_t(0)
        self._check(self.L.lgca_b200_ring_descriptor_bytes(C.byref(n)))
        buf = (C.c_uint8 * n.value)()
        self._check(self.L.lgca_b200_ring_export(self.h, buf, n.value))
        return bytes(buf)

    def ring_connect(self, lower_descriptor, upper_descriptor):
        lo = (C.c_uint8 * len(lower_descriptor)).from_buffer_copy(lower_descriptor)
        up = (C.c_uint8 * len(upper_descriptor)).from_buffer_copy(upper_descriptor)
        self._check(self.L.lgca_b200_ring_connect(self.h, lo, up))

    def ring_start(self):
        self._check(self.L.lgca_b200_ring_start(self.h))

    def ring_step(self, n):
        self._check(self.L.lgca_b200_ring_step(self.h, int(n)))

    def ring_disconnect(self):
        self._check(self.L.lgca_b200_ring_disconnect(self.h))

    def ring_republish(self):
        self._check(self.L.lgca_b200_ring_republish(self.h))

    def wall_flags(self):
        a, b = C.c_uint32(0), C.c_uint32(0)
        self._check(self.L.lgca_b200_get_wall_flags(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_wall_flags(self, has_no_slip, has_slip):
        self._check(self.L.lgca_b200_set_wall_flags(self.h, int(has_no_slip), int(has_slip)))


class Group:
    """One lattice on n GPUs of this box driven by ONE process (lgca_b200_group_*): same calls as Engine, all host
    arrays are GLOBAL reference-layout arrays.  n_gpus == 1 is a plain whole-lattice handle."""

    def __init__(self, model, dim_x, dim_y, n_gpus=1, dev_ids=None, cg_radius=0, bf_dir=0, k_fuse=0, flags=0):
        self.L = load_library()
        self.model = MODELS[model] if isinstance(model, str) else int(model)
        if isinstance(bf_dir, (bytes, str)):
            bf_dir = ord(bf_dir) if bf_dir not in (b"\0", "\0", "", b"") else 0
        cfg = Config(self.model, dim_x, dim_y, cg_radius, bf_dir, 0, k_fuse, 0, 0, flags)
        ids = None
        if dev_ids is not None:
            ids = (C.c_int * n_gpus)(*dev_ids)
        g = C.c_void_p()
        self._check(self.L.lgca_b200_group_create(C.byref(cfg), int(n_gpus), ids, C.byref(g)))
        self.g = g
        self.n_gpus = int(n_gpus)
        self.dim_x, self.dim_y, self.cg = dim_x, dim_y, cg_radius
        self.cells = dim_x * dim_y
        self.num_dir = NUM_DIR[self.model]

    _check = Engine._check

    def close(self):
        if getattr(self, "g", None):
            self.L.lgca_b200_group_destroy(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, state=None, cell_type=None, rnd_bits=None):
        if state is not None:
            state = np.ascontiguousarray(state, np.uint8)
            assert state.size == self.cells
        if cell_type is not None:
            cell_type = np.ascontiguousarray(cell_type, np.int32)
            assert cell_type.size == self.cells
        if rnd_bits is not None:
            rnd_bits = np.ascontiguousarray(rnd_bits, np.uint8)
            assert rnd_bits.size >= (self.cells + 7) // 8
        self._check(self.L.lgca_b200_group_upload(self.g, _ptr(state), _ptr(cell_type), _ptr(rnd_bits)))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.cells, np.uint8)
        self._check(self.L.lgca_b200_group_download(self.g, _ptr(out)))
        return out

    def step(self, n=1):
        self._check(self.L.lgca_b200_group_step(self.g, int(n)))

    def timed_steps(self, n):
        ms = C.c_float(0)
        self._check(self.L.lgca_b200_group_timed_steps(self.g, int(n), C.byref(ms)))
        return float(ms.value)

    def sync(self):
        self._check(self.L.lgca_b200_group_sync(self.g))

    def snapshot(self):
        self._check(self.L.lgca_b200_group_snapshot(self.g))

    def post_process(self, cell=True, mean=True, exact=True, out=None):
        out = {} if out is None else out
        n = self.cells
        if cell:
            out.setdefault("cell_density", np.empty(n, np.float32))
            out.setdefault("cell_momentum", np.empty(2 * n, np.float32))
        if mean:
            nc = (self.dim_x // (2 * self.cg)) * (self.dim_y // (2 * self.cg))
            out.setdefault("mean_density", np.empty(nc, np.float32))
            out.setdefault("mean_momentum", np.empty(2 * nc, np.float32))
        self._check(self.L.lgca_b200_group_post_process(self.g, _ptr(out.get("cell_density")), _ptr(out.get("cell_momentum")),
                                                        _ptr(out.get("mean_density")), _ptr(out.get("mean_momentum")),
                                                        1 if exact else 0))
        return out

    def mean_velocity_exact(self):
        out = np.zeros(2, np.float32)
        self._check(self.L.lgca_b200_group_mean_velocity_exact(self.g, _ptr(out)))
        return out

    def mean_velocity(self):
        out = np.zeros(2, np.float32)
        self._check(self.L.lgca_b200_group_mean_velocity(self.g, _ptr(out)))
        return out

    def body_force(self, forcing, draws):
        draws = np.ascontiguousarray(draws, np.int32)
        used, rev = C.c_size_t(0), C.c_uint32(0)
        self._check(self.L.lgca_b200_group_body_force(self.g, int(forcing), _ptr(draws), draws.size, C.byref(used), C.byref(rev)))
        return int(used.value), int(rev.value)

    def count_particles(self):
        v = C.c_uint64(0)
        self._check(self.L.lgca_b200_group_count_particles(self.g, C.byref(v)))
        return int(v.value)

    def init_random_device(self, seed=1):
        self._check(self.L.lgca_b200_group_init_random_device(self.g, int(seed)))

    def apply_bc_device(self, name):
        self._check(self.L.lgca_b200_group_apply_bc_device(self.g, name.encode()))

    def launch_count(self):
        v = C.c_uint64(0)
        self._check(self.L.lgca_b200_group_launch_count(self.g, C.byref(v)))
        return int(v.value)

    def info(self):
        i = Info()
        self._check(self.L.lgca_b200_group_get_info(self.g, C.byref(i)))
        return i
