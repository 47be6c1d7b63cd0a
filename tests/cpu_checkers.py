"""ctypes bindings for the CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  -- oracle/liblgca_oracle.so, the plain-C restatement (oracle/lgca_oracle.c).
* ``Ref``     -- oracle/_ref/liblgca_ref.so, the UNMODIFIED reference CPU backend behind
                 oracle/ref_driver.cpp (built only where /root/reference exists; travels prebuilt).

Nothing under lgca_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liblgca_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "liblgca_ref.so")

MODELS = {"HPP": 0, "FHP_I": 1, "FHP_II": 2, "FHP_III": 3}
NUM_DIR = {0: 4, 1: 6, 2: 7, 3: 7}


def fnv1a64(buf) -> str:
    """FNV-1a 64 as used by SURVEY.md Appendix B (via the oracle's C implementation)."""
    a = np.ascontiguousarray(buf).view(np.uint8).ravel()
    return "%016x" % _oracle_lib().lgca_oracle_fnv1a64(a.ctypes.data_as(C.c_void_p), a.size)


class _Rng(C.Structure):
    _fields_ = [("r", C.c_int32 * 34), ("f", C.c_int), ("b", C.c_int)]


class _Params(C.Structure):
    _fields_ = [
        ("model", C.c_int), ("num_dir", C.c_int),
        ("dim_x", C.c_uint32), ("dim_y", C.c_uint32), ("num_cells", C.c_uint64),
        ("cg_radius", C.c_uint32), ("coarse_dim_x", C.c_uint32), ("coarse_dim_y", C.c_uint32),
        ("num_coarse_cells", C.c_uint64),
        ("Re", C.c_float), ("Ma_s", C.c_float), ("d", C.c_float), ("nu", C.c_float), ("g", C.c_float),
        ("nu_s", C.c_float), ("c_s", C.c_float), ("u", C.c_float), ("bf_dir", C.c_char),
    ]


_ORACLE = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


def _oracle_lib():
    global _ORACLE
    if _ORACLE is None:
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
                os.path.join(ROOT, "oracle", "lgca_oracle.c")):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        PP = C.POINTER(_Params)
        RP = C.POINTER(_Rng)
        L.lgca_oracle_srand.argtypes = [RP, C.c_uint]
        L.lgca_oracle_rand.argtypes = [RP]
        L.lgca_oracle_rand.restype = i32
        L.lgca_oracle_num_dir.argtypes = [i32]
        L.lgca_oracle_params_init.argtypes = [PP, i32, C.c_char_p, C.c_float, C.c_float, i32]
        L.lgca_oracle_params_dims.argtypes = [PP, i32, u32, u32, i32, C.c_char]
        L.lgca_oracle_initial_forcing.argtypes = [PP]
        L.lgca_oracle_initial_forcing.restype = u64
        L.lgca_oracle_equilibrium_forcing.argtypes = [PP]
        L.lgca_oracle_equilibrium_forcing.restype = u64
        L.lgca_oracle_fill_rnd.argtypes = [PP, vp, RP]
        L.lgca_oracle_apply_bc.argtypes = [PP, C.c_char_p, vp]
        L.lgca_oracle_init.argtypes = [PP, C.c_char_p, vp, vp, RP]
        L.lgca_oracle_collide_cell.argtypes = [i32, vp, vp, i32]
        L.lgca_oracle_step.argtypes = [PP, vp, vp, vp, vp]
        L.lgca_oracle_steps.argtypes = [PP, vp, vp, vp, vp, i32]
        L.lgca_oracle_body_force.argtypes = [PP, vp, vp, i32, RP, C.POINTER(u32)]
        L.lgca_oracle_body_force.restype = u64
        L.lgca_oracle_cell_post_process.argtypes = [PP, vp, vp, vp]
        L.lgca_oracle_mean_post_process.argtypes = [PP, vp, vp, vp, vp]
        L.lgca_oracle_mean_velocity.argtypes = [PP, vp, vp, vp, vp]
        L.lgca_oracle_n_particles.argtypes = [PP, vp]
        L.lgca_oracle_n_particles.restype = u64
        L.lgca_oracle_fnv1a64.argtypes = [vp, u64]
        L.lgca_oracle_fnv1a64.restype = u64
        _ORACLE = L
    return _ORACLE


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleRng:
    """glibc rand() restated (seed 1 = the reference's never-seeded stream)."""

    def __init__(self, seed=1):
        self.L = _oracle_lib()
        self.g = _Rng()
        self.L.lgca_oracle_srand(C.byref(self.g), seed)

    def rand(self):
        return self.L.lgca_oracle_rand(C.byref(self.g))


class Oracle:
    """Host-array lattice driven by the plain-C oracle; mirrors the reference call sequence."""

    def __init__(self, model, test_case=None, Re=80.0, Ma=0.2, cg=16, dims=None, bf_dir=b"\0", rng=None):
        self.L = _oracle_lib()
        self.model = MODELS[model] if isinstance(model, str) else model
        self.p = _Params()
        if dims is not None:
            rc = self.L.lgca_oracle_params_dims(C.byref(self.p), self.model, dims[0], dims[1], cg, bf_dir)
        else:
            rc = self.L.lgca_oracle_params_init(C.byref(self.p), self.model, test_case.encode(), Re, Ma, cg)
        if rc != 0:
            raise ValueError("invalid oracle parameters")
        self.rng = rng if rng is not None else OracleRng(1)
        n = int(self.p.num_cells)
        self.dim_x, self.dim_y, self.num_cells = int(self.p.dim_x), int(self.p.dim_y), n
        self.num_dir = int(self.p.num_dir)
        self.state = np.zeros(n, np.uint8)
        self.scratch = np.zeros(n, np.uint8)
        self.state_out = np.zeros(n, np.uint8)
        self.cell_type = np.zeros(n, np.int32)
        self.rnd = np.zeros((n + 7) // 8, np.uint8)
        nc = int(self.p.num_coarse_cells)
        self.cell_density = np.zeros(n, np.float32)
        self.cell_momentum = np.zeros(2 * n, np.float32)
        self.mean_density = np.zeros(nc, np.float32)
        self.mean_momentum = np.zeros(2 * nc, np.float32)
        # ctor order of the reference: allocate, then fill the chirality bits (src/omp_lattice.cpp:81-84)
        self.L.lgca_oracle_fill_rnd(C.byref(self.p), _p(self.rnd), C.byref(self.rng.g))

    @property
    def u(self):
        return float(self.p.u)

    def apply_bc(self, name):
        if self.L.lgca_oracle_apply_bc(C.byref(self.p), name.encode(), _p(self.cell_type)) != 0:
            raise ValueError(name)

    def init(self, name):
        if self.L.lgca_oracle_init(C.byref(self.p), name.encode(), _p(self.state), _p(self.cell_type),
                                   C.byref(self.rng.g)) != 0:
            raise ValueError(name)

    def step(self, n=1):
        self.L.lgca_oracle_steps(C.byref(self.p), _p(self.state), _p(self.scratch), _p(self.cell_type),
                                 _p(self.rnd), n)

    def body_force(self, forcing):
        rev = C.c_uint32(0)
        used = self.L.lgca_oracle_body_force(C.byref(self.p), _p(self.state), _p(self.cell_type), int(forcing),
                                             C.byref(self.rng.g), C.byref(rev))
        return int(used), int(rev.value)

    def snapshot(self):
        self.state_out[:] = self.state

    def post_process(self):
        self.L.lgca_oracle_cell_post_process(C.byref(self.p), _p(self.state_out), _p(self.cell_density),
                                             _p(self.cell_momentum))
        self.L.lgca_oracle_mean_post_process(C.byref(self.p), _p(self.cell_density), _p(self.cell_momentum),
                                             _p(self.mean_density), _p(self.mean_momentum))

    def mean_velocity(self):
        out = np.zeros(2, np.float32)
        self.L.lgca_oracle_mean_velocity(C.byref(self.p), _p(self.cell_type), _p(self.cell_density),
                                         _p(self.cell_momentum), _p(out))
        return out

    def n_particles(self):
        return int(self.L.lgca_oracle_n_particles(C.byref(self.p), _p(self.state)))

    def initial_forcing(self):
        return int(self.L.lgca_oracle_initial_forcing(C.byref(self.p)))

    def equilibrium_forcing(self):
        return int(self.L.lgca_oracle_equilibrium_forcing(C.byref(self.p)))

    def hash(self):
        return fnv1a64(self.state)


def collide_table(model):
    """Full truth table {(state, p): state'} of ModelDescriptor<M>::collide via the oracle."""
    L = _oracle_lib()
    m = MODELS[model] if isinstance(model, str) else model
    nd = NUM_DIR[m]
    table = {}
    for s in range(1 << nd):
        for p in (0, 1):
            i = np.array([(s >> d) & 1 for d in range(8)], np.uint8)
            o = np.zeros(8, np.uint8)
            L.lgca_oracle_collide_cell(m, _p(i), _p(o), p)
            table[(s, p)] = int(sum(int(o[d]) << d for d in range(nd)))
    return table


# ----------------------------------------------------------------------------------------------
# The unmodified reference (oracle/_ref)
# ----------------------------------------------------------------------------------------------
_REF = None


def ref_available():
    return os.path.exists(REF_SO)


def _ref_lib():
    global _REF
    if _REF is None:
        L = C.CDLL(REF_SO)
        vp = C.c_void_p
        L.lgca_ref_create.restype = vp
        L.lgca_ref_create.argtypes = [C.c_int, C.c_char_p, C.c_float, C.c_float, C.c_int]
        L.lgca_ref_destroy.argtypes = [vp]
        L.lgca_ref_resize.argtypes = [vp, C.c_uint, C.c_uint]
        for f in ("dim_x", "dim_y", "coarse_dim_x", "coarse_dim_y"):
            getattr(L, "lgca_ref_" + f).restype = C.c_uint
            getattr(L, "lgca_ref_" + f).argtypes = [vp]
        L.lgca_ref_num_dir.argtypes = [vp]
        L.lgca_ref_u.restype = C.c_float
        L.lgca_ref_u.argtypes = [vp]
        L.lgca_ref_apply_bc.argtypes = [vp, C.c_char_p]
        L.lgca_ref_init.argtypes = [vp, C.c_char_p]
        for f in ("state", "state_out", "rnd"):
            getattr(L, "lgca_ref_" + f).restype = C.POINTER(C.c_uint8)
            getattr(L, "lgca_ref_" + f).argtypes = [vp]
        L.lgca_ref_cell_type.restype = C.POINTER(C.c_int32)
        L.lgca_ref_cell_type.argtypes = [vp]
        L.lgca_ref_step.argtypes = [vp, C.c_int]
        L.lgca_ref_body_force.argtypes = [vp, C.c_int]
        L.lgca_ref_snapshot.argtypes = [vp]
        L.lgca_ref_post_process.argtypes = [vp]
        L.lgca_ref_mean_velocity.argtypes = [vp, vp]
        L.lgca_ref_n_particles.restype = C.c_ulong
        L.lgca_ref_n_particles.argtypes = [vp]
        for f in ("cell_density", "cell_momentum", "mean_density", "mean_momentum"):
            getattr(L, "lgca_ref_" + f).restype = C.POINTER(C.c_float)
            getattr(L, "lgca_ref_" + f).argtypes = [vp]
        L.lgca_ref_initial_forcing.restype = C.c_size_t
        L.lgca_ref_initial_forcing.argtypes = [vp]
        L.lgca_ref_equilibrium_forcing.restype = C.c_size_t
        L.lgca_ref_equilibrium_forcing.argtypes = [vp]
        L.lgca_ref_set_bf_dir.argtypes = [vp, C.c_char]
        L.lgca_ref_set_threads.argtypes = [C.c_int]
        L.lgca_ref_srand.argtypes = [C.c_uint]
        _REF = L
    return _REF


class Ref:
    """The reference's OMP_Lattice<Model> through oracle/ref_driver.cpp.

    The libc rand() stream is process-global: ``seed`` re-seeds it before construction so that every
    instance sees the never-seeded (= seed 1) stream of a fresh reference process.
    """

    def __init__(self, model, test_case, Re=80.0, Ma=0.2, cg=16, seed=1, threads=1, dims=None):
        self.L = _ref_lib()
        self.L.lgca_ref_set_threads(threads)
        if seed is not None:
            self.L.lgca_ref_srand(seed)
        self.model = MODELS[model] if isinstance(model, str) else model
        self.h = self.L.lgca_ref_create(self.model, test_case.encode(), Re, Ma, cg)
        if dims is not None:
            self.L.lgca_ref_resize(self.h, dims[0], dims[1])
        self.dim_x, self.dim_y = self.L.lgca_ref_dim_x(self.h), self.L.lgca_ref_dim_y(self.h)
        self.num_cells = self.dim_x * self.dim_y
        self.num_dir = self.L.lgca_ref_num_dir(self.h)
        self.num_coarse = self.L.lgca_ref_coarse_dim_x(self.h) * self.L.lgca_ref_coarse_dim_y(self.h)

    def close(self):
        if self.h:
            self.L.lgca_ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def u(self):
        return float(self.L.lgca_ref_u(self.h))

    # live views into the reference's own arrays
    @property
    def state(self):
        return np.ctypeslib.as_array(self.L.lgca_ref_state(self.h), shape=(self.num_cells,))

    @property
    def state_out(self):
        return np.ctypeslib.as_array(self.L.lgca_ref_state_out(self.h), shape=(self.num_cells,))

    @property
    def cell_type(self):
        return np.ctypeslib.as_array(self.L.lgca_ref_cell_type(self.h), shape=(self.num_cells,))

    @property
    def rnd(self):
        return np.ctypeslib.as_array(self.L.lgca_ref_rnd(self.h), shape=((self.num_cells + 7) // 8,))

    def _f(self, name, n):
        return np.ctypeslib.as_array(getattr(self.L, "lgca_ref_" + name)(self.h), shape=(n,))

    @property
    def cell_density(self):
        return self._f("cell_density", self.num_cells)

    @property
    def cell_momentum(self):
        return self._f("cell_momentum", 2 * self.num_cells)

    @property
    def mean_density(self):
        return self._f("mean_density", self.num_coarse)

    @property
    def mean_momentum(self):
        return self._f("mean_momentum", 2 * self.num_coarse)

    def apply_bc(self, name):
        self.L.lgca_ref_apply_bc(self.h, name.encode())

    def init(self, name):
        self.L.lgca_ref_init(self.h, name.encode())

    def step(self, n=1):
        self.L.lgca_ref_step(self.h, n)

    def body_force(self, forcing):
        self.L.lgca_ref_body_force(self.h, int(forcing))

    def snapshot(self):
        self.L.lgca_ref_snapshot(self.h)

    def post_process(self):
        self.L.lgca_ref_post_process(self.h)

    def mean_velocity(self):
        out = np.zeros(2, np.float32)
        self.L.lgca_ref_mean_velocity(self.h, _p(out))
        return out

    def n_particles(self):
        return int(self.L.lgca_ref_n_particles(self.h))

    def initial_forcing(self):
        return int(self.L.lgca_ref_initial_forcing(self.h))

    def equilibrium_forcing(self):
        return int(self.L.lgca_ref_equilibrium_forcing(self.h))

    def set_bf_dir(self, c):
        self.L.lgca_ref_set_bf_dir(self.h, c)

    def set_threads(self, n):
        self.L.lgca_ref_set_threads(n)

    def hash(self):
        return fnv1a64(self.state)
