"""B200Options::bulk_rand (host layer, no GPU needed): body-force draws taken in bulk from glibc's default generator by
borrowing its state through initstate()/setstate() must be the very values rand() would have returned, leave the generator
exactly where a run of rand() calls would have left it, and interleave freely with plain rand() calls (the reference's
apply_body_force calls rand() once per draw, src/omp_lattice.cpp:269)."""
import ctypes as C
import os

import numpy as np
import pytest

HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lgca_b200", "host", "liblgca_host.so")


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(HOST):
        import subprocess
        from lgca_b200.build import build_library
        build_library()
        subprocess.check_call(["make", "-C", os.path.dirname(HOST), "-j4"], stdout=subprocess.DEVNULL)
    host = C.CDLL(HOST)
    host.lgca_host_bulk_rand.argtypes = [C.c_void_p, C.c_size_t]
    host.lgca_host_bulk_rand_selfcheck.argtypes = [C.c_void_p]
    libc = C.CDLL(None)
    libc.srand.argtypes = [C.c_uint]
    return host, libc


def bulk(host, n):
    out = np.empty(n, np.int32)
    assert host.lgca_host_bulk_rand(out.ctypes.data_as(C.c_void_p), n) == 1
    return out.tolist()


@pytest.mark.parametrize("seed", [1, 123, 2**31 - 1])
def test_bulk_draws_are_the_rand_stream(libs, seed):
    host, libc = libs
    libc.srand(seed)
    want = [libc.rand() for _ in range(6000)]
    libc.srand(seed)
    got = [libc.rand() for _ in range(100)]
    got += bulk(host, 1000)
    got += [libc.rand() for _ in range(50)]          # libc continues where the bulk fill stopped
    got += bulk(host, 3)
    got += bulk(host, 3997)
    got += [libc.rand() for _ in range(850)]
    assert got == want


def test_selfcheck_consumes_the_next_32_values(libs):
    host, libc = libs
    libc.srand(77)
    want = [libc.rand() for _ in range(40)]
    libc.srand(77)
    out = np.empty(32, np.int32)
    assert host.lgca_host_bulk_rand_selfcheck(out.ctypes.data_as(C.c_void_p)) == 1
    assert out.tolist() == want[:32]
    assert [libc.rand() for _ in range(8)] == want[32:]


def test_other_generator_types_are_left_alone(libs):
    """A caller that installed its own (smaller) state array gets plain rand(): the bulk path refuses anything but TYPE_3."""
    host, libc = libs
    buf = C.create_string_buffer(32)                   # TYPE_1: degree 7
    libc.initstate.restype = C.c_void_p
    libc.initstate.argtypes = [C.c_uint, C.c_void_p, C.c_size_t]
    libc.setstate.restype = C.c_void_p
    libc.setstate.argtypes = [C.c_void_p]
    old = libc.initstate(5, buf, 32)
    try:
        a = libc.rand()
        out = np.empty(4, np.int32)
        assert host.lgca_host_bulk_rand(out.ctypes.data_as(C.c_void_p), 4) == 0
        libc.initstate(5, buf, 32)
        assert libc.rand() == a                          # and the caller's generator still works as before
    finally:
        libc.setstate(old)
