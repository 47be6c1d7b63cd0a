"""The dependency-free PNG writer of the host layer (lgca_b200/host/lgca_io_png.cpp): file structure
(signature, IHDR/IDAT/IEND chunks, CRC-32, zlib stream with Adler-32) and the colours of VTK's blue->red rainbow table
as the reference's viewers configure it (apps/pipe/pipe_viewer.cpp:311-315).  CPU only; a GPU test writes one through
the lgca-pipe app (tests/test_host_apps.py)."""
import colorsys
import ctypes as C
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "lgca_b200", "host")


@pytest.fixture(scope="module")
def host_lib():
    from lgca_b200.build import build_library
    build_library()
    subprocess.check_call(["make", "-s", "-C", HOST])
    L = C.CDLL(os.path.join(HOST, "liblgca_host.so"))
    L.lgca_host_write_png.argtypes = [C.c_char_p, C.c_uint, C.c_uint, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_uint]
    return L


def read_png(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        data = raw[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
        assert crc == zlib.crc32(typ + data) & 0xFFFFFFFF, typ
        chunks.append((typ, data))
        pos += 12 + n
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    w, h, depth, ctype, comp, filt, inter = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, ctype, comp, filt, inter) == (8, 2, 0, 0, 0)
    scan = zlib.decompress(chunks[1][1])       # checks the Adler-32 too
    assert len(scan) == h * (3 * w + 1)
    img = np.frombuffer(scan, np.uint8).reshape(h, 3 * w + 1)
    assert not img[:, 0].any()                  # filter type 0 on every scanline
    return img[:, 1:].reshape(h, w, 3)


def vtk_rainbow(values, lo, hi):
    """vtkLookupTable: 256 entries, hue 2/3 -> 0, saturation = value = 1."""
    idx = np.clip(np.floor((values.astype(np.float64) - lo) * (256.0 / (hi - lo))), 0, 255).astype(int)
    table = np.array([[round(c * 255) for c in colorsys.hsv_to_rgb((2.0 / 3.0) * (1 - i / 255.0), 1.0, 1.0)] for i in range(256)], np.uint8)
    return table[idx]


def test_png_scalar_field(host_lib, tmp_path):
    w, h = 37, 21
    v = np.linspace(-1.0, 3.0, w * h, dtype=np.float32).reshape(h, w)
    path = str(tmp_path / "a.png")
    assert host_lib.lgca_host_write_png(path.encode(), w, h, v.ctypes.data_as(C.c_void_p), 1, 1.0, 0.0, 1) == 0  # lo > hi: data range
    img = read_png(path)
    want = vtk_rainbow(v, float(v.min()), float(v.max()))[::-1]     # lattice row 0 is the bottom of the picture
    assert np.array_equal(img, want)
    assert tuple(img[-1, 0]) == (0, 0, 255) and tuple(img[0, -1]) == (255, 0, 0)   # lowest value blue, highest red


def test_png_vector_magnitude_zoom_and_big_image(host_lib, tmp_path):
    w, h, zoom = 300, 200, 2          # 600 x 400 x 3 bytes: several stored deflate blocks
    rs = np.random.RandomState(1)
    v = rs.rand(h, w, 2).astype(np.float32) - 0.5
    path = str(tmp_path / "b.png")
    assert host_lib.lgca_host_write_png(path.encode(), w, h, v.ctypes.data_as(C.c_void_p), 2, 0.0, 0.5, zoom) == 0
    img = read_png(path)
    assert img.shape == (h * zoom, w * zoom, 3)
    mag = np.sqrt(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]).astype(np.float32)
    want = np.repeat(np.repeat(vtk_rainbow(mag, 0.0, 0.5)[::-1], zoom, axis=0), zoom, axis=1)
    assert np.array_equal(img, want)
