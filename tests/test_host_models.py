"""Host ModelDescriptor<M> (lgca_b200/host/lgca_models.h): the per-cell rules against the oracle's truth tables, the
generated streaming offset tables against the reference's own (where /root/reference exists) and against the
oracle's streaming (single particles walked over a small torus, every cell and direction).  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cpu_checkers import MODELS, NUM_DIR, Oracle, collide_table

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, "host_shims")
INV = {4: [2, 3, 0, 1], 6: [3, 4, 5, 0, 1, 2], 7: [3, 4, 5, 0, 1, 2, 6]}
MIR_X = {4: [0, 3, 2, 1], 6: [0, 5, 4, 3, 2, 1], 7: [0, 5, 4, 3, 2, 1, 6]}
MIR_Y = {4: [2, 1, 0, 3], 6: [3, 2, 1, 0, 5, 4], 7: [3, 2, 1, 0, 5, 4, 6]}
NAMES = ["neighbor_even", "neighbor_odd", "eastern_even", "eastern_odd", "northern_even", "northern_odd", "western_even",
         "western_odd", "southern_even", "southern_odd"]


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(SHIMS, "libmodels_host.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", out,
                           os.path.join(SHIMS, "models_host.cpp")])
    L = C.CDLL(out)
    L.lgca_host_model_offsets.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_void_p]
    L.lgca_host_model_rule.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    return L


def offsets(fn, model, dx, dy):
    out = np.zeros((10, 7), np.int32)
    fn(MODELS[model], dx, dy, out.ctypes.data_as(C.c_void_p))
    return out


@pytest.mark.parametrize("model", ["HPP", "FHP_I", "FHP_II", "FHP_III"])
def test_rules_match_oracle(lib, model):
    m, nd = MODELS[model], NUM_DIR[MODELS[model]]
    table = collide_table(model)
    for s in range(1 << nd):
        i = np.array([(s >> d) & 1 for d in range(8)], np.uint8)
        for p in (0, 1):
            o = np.zeros(8, np.uint8)
            lib.lgca_host_model_rule(m, 0, i.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), p)
            assert sum(int(o[d]) << d for d in range(nd)) == table[(s, p)], (s, p)
        for what, perm in ((1, INV[nd]), (2, MIR_X[nd]), (3, MIR_Y[nd])):
            o = np.zeros(8, np.uint8)
            lib.lgca_host_model_rule(m, what, i.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), 0)
            assert [int(o[d]) for d in range(nd)] == [int(i[perm[d]]) for d in range(nd)]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
@pytest.mark.parametrize("model", ["HPP", "FHP_I", "FHP_II", "FHP_III"])
def test_offset_tables_equal_the_references(lib, model):
    out = os.path.join(SHIMS, "libref_models.so")
    subprocess.check_call(["g++", "-O1", "-std=c++11", "-fPIC", "-shared", "-w", "-I", os.path.join(HERE, "..", "oracle", "shim"),
                           "-I", "/root/reference/src", "-o", out, os.path.join(SHIMS, "ref_models.cpp")])
    R = C.CDLL(out)
    R.lgca_ref_model_offsets.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_void_p]
    for dx, dy in ((21, 10), (64, 32), (1400, 700)):
        mine, ref = offsets(lib.lgca_host_model_offsets, model, dx, dy), offsets(R.lgca_ref_model_offsets, model, dx, dy)
        for a, name in enumerate(NAMES):
            for d in range(NUM_DIR[MODELS[model]]):
                if model != "HPP" and name == "southern_even" and d == 2:
                    # the reference's stray +1 (src/lgca_models.h:328): unreachable -- even rows are never the top row
                    assert ref[a, d] == mine[a, d] + 1
                    continue
                assert mine[a, d] == ref[a, d], (name, d, dx, dy)


@pytest.mark.parametrize("model,dims", [("HPP", (6, 5)), ("FHP_I", (6, 4)), ("FHP_III", (5, 6))])
def test_offset_tables_reproduce_the_oracles_streaming(lib, model, dims):
    """Pull streaming with the tables, as src/omp_lattice.cpp:150-176 applies them, == one oracle step of a lone particle."""
    dx, dy = dims
    nd = NUM_DIR[MODELS[model]]
    t = offsets(lib.lgca_host_model_offsets, model, dx, dy)
    tab = {n: t[a] for a, n in enumerate(NAMES)}

    def source(cell, d):   # the cell direction d is pulled from
        inv = INV[nd][d]
        par = "odd" if (cell // dx) % 2 else "even"
        off = tab["neighbor_" + par][inv]
        if (cell + 1) % dx == 0:
            off += tab["western_" + par][inv]
        if cell >= dx * dy - dx:
            off += tab["southern_" + par][inv]
        if cell % dx == 0:
            off += tab["eastern_" + par][inv]
        if cell < dx:
            off += tab["northern_" + par][inv]
        return cell + off

    for d in range(nd):
        src = np.array([source(c, d) for c in range(dx * dy)])
        assert sorted(src) == list(range(dx * dy))          # a permutation of the torus
        for start in range(dx * dy):
            o = Oracle(model, dims=dims, cg=1)
            o.state[start] = 1 << d
            o.step(1)
            (dest,) = np.nonzero(o.state)[0]
            assert o.state[dest] == 1 << d and src[dest] == start, (d, start)
