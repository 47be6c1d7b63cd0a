"""CPU check of the bit-sliced boolean networks the CUDA kernels use (lgca_b200/csrc/lgca_collide.cuh),
compiled for the host, against the oracle's per-cell rules -- exhaustively over every (state, chirality,
cell type, edge) combination.  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cpu_checkers import MODELS, NUM_DIR, collide_table

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_shims", "collide_host.cpp")
OUT = os.path.join(HERE, "host_shims", "libcollide_host.so")

INV = {4: [2, 3, 0, 1], 6: [3, 4, 5, 0, 1, 2], 7: [3, 4, 5, 0, 1, 2, 6]}
MIR_X = {4: [0, 3, 2, 1], 6: [0, 5, 4, 3, 2, 1], 7: [0, 5, 4, 3, 2, 1, 6]}
MIR_Y = {4: [2, 1, 0, 3], 6: [3, 2, 1, 0, 5, 4], 7: [3, 2, 1, 0, 5, 4, 6]}


@pytest.fixture(scope="module")
def lib():
    hdr = os.path.join(HERE, "..", "lgca_b200", "csrc", "lgca_collide.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-o", OUT, SRC])
    L = C.CDLL(OUT)
    L.lgca_host_collide_words.argtypes = [C.c_int, C.c_void_p] + [C.c_uint32] * 5
    return L


def permute(s, perm, nd):
    return sum(((s >> perm[d]) & 1) << d for d in range(nd))


@pytest.mark.parametrize("model", ["HPP", "FHP_I", "FHP_II", "FHP_III"])
def test_network_matches_oracle_rules(lib, model):
    m = MODELS[model]
    nd = NUM_DIR[m]
    table = collide_table(model)
    # one site per bit: enumerate all (state, p, type, ew, ns_row) in batches of 32 sites
    cases = [(s, p, t, ew, nsr) for s in range(1 << nd) for p in (0, 1) for t in (0, 1, 2)
             for ew in (0, 1) for nsr in (0, 1)]
    # ns_row is a per-word (per-row) flag: batch sites that share it
    for nsr in (0, 1):
        sub = [c for c in cases if c[4] == nsr]
        for i in range(0, len(sub), 32):
            chunk = sub[i:i + 32]
            n = np.zeros(7, np.uint32)
            pw = nsw = slw = eww = 0
            for b, (s, p, t, ew, _) in enumerate(chunk):
                for d in range(nd):
                    n[d] |= np.uint32(((s >> d) & 1) << b)
                pw |= p << b
                nsw |= (1 if t == 1 else 0) << b
                slw |= (1 if t == 2 else 0) << b
                eww |= ew << b
            lib.lgca_host_collide_words(m, n.ctypes.data_as(C.c_void_p), pw, nsw, slw, eww,
                                        0xFFFFFFFF if nsr else 0)
            for b, (s, p, t, ew, _) in enumerate(chunk):
                got = sum(((int(n[d]) >> b) & 1) << d for d in range(nd))
                if t == 0:
                    want = table[(s, p)]
                elif t == 1:
                    want = permute(s, INV[nd], nd)
                else:
                    want = s
                    if nsr:
                        want = permute(s, MIR_X[nd], nd)
                    if ew:
                        want = permute(s, MIR_Y[nd], nd)
                assert got == want, (model, s, p, t, ew, nsr, got, want)
