"""The C++ host layer (Lattice<Model>, B200_Lattice<Model>, IoVti, headless apps) end to end.

CPU part: the layer builds with plain g++ and the apps fail loudly without a GPU (no CPU fallback).
GPU part (-m gpu): the apps, driven by the process's real glibc rand() stream exactly like the reference's
viewers, must reproduce the reference's state hashes -- incl. BASELINE config C1, the
1000-step FHP-I pipe run with body force."""
import json
import os
import re
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "lgca_b200", "host", "bin")
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_b.json")))
REF_RUNS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_runs.json")))


@pytest.fixture(scope="module")
def apps():
    from lgca_b200.build import build_library
    build_library()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "lgca_b200", "host")])
    return BIN


def run_app(name, *args, check=True):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([os.path.join(BIN, name)] + [str(a) for a in args], capture_output=True, text=True, env=env)
    if check:
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p


def hashes(out):
    return {int(m.group(1)): m.group(2) for m in re.finditer(r"HASH step (\d+) ([0-9a-f]{16})", out)}


def test_apps_build(apps):
    for a in ("pipe", "karman", "diffusion", "single", "box", "periodic"):
        assert os.path.exists(os.path.join(apps, "lgca-" + a))


def test_apps_fail_loudly_without_gpu(apps):
    import lgca_b200
    if lgca_b200.load_library().lgca_b200_device_count() > 0:
        pytest.skip("GPU present")
    p = run_app("lgca-periodic", "--steps", 2, check=False)
    assert p.returncode != 0
    assert "ERROR in B200_Lattice" in p.stdout + p.stderr


@pytest.mark.gpu
def test_c1_pipe_fhp1_1000_steps(apps):
    """BASELINE config C1: lgca-pipe, FHP-I, app default size 1400x700, canonical schedule, 1000 steps."""
    case = [c for c in GOLD["b3_pipe_schedule"] if c["model"] == "FHP_I"][0]
    p = run_app("lgca-pipe", "--model", "FHP_I", "--steps", 1000, "--hash-every", 500, "--quiet")
    h = hashes(p.stdout)
    assert h[0] == case["hashes"]["0"]
    assert h[500] == case["hashes"]["500"]
    assert h[1000] == case["hashes"]["1000"]
    assert "Error check PASSED" in p.stdout


@pytest.mark.gpu
def test_pipe_fhp3_default(apps):
    case = [c for c in GOLD["b3_pipe_schedule"] if c["model"] == "FHP_III"][0]
    p = run_app("lgca-pipe", "--steps", 500, "--hash-every", 100, "--quiet")
    h = hashes(p.stdout)
    for s in (0, 100, 500):
        assert h[s] == case["hashes"][str(s)]


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [None, "0,0"], ids=["1gpu", "2strips"])
def test_karman_app_default_1000_steps(apps, devices):
    """The north-star target through the C++ drop-in: lgca-karman at the app's defaults (FHP-III 4400 x 2200), driven by
    the process's real glibc rand() stream, 1000 steps on the canonical schedule -- state hashes equal to the
    unmodified reference's (tests/golden/reference_runs.json).  `--devices 0,0` runs it as two row strips behind the
    same B200_Lattice (B200Options::n_gpus; both strips on device 0 where the box has one GPU)."""
    gold = REF_RUNS["karman_default"]
    extra = ["--devices", devices] if devices else []
    p = run_app("lgca-karman", "--steps", 1000, "--hash-every", 500, "--quiet", *extra)
    h = hashes(p.stdout)
    for s in (0, 500, 1000):
        assert h[s] == gold["hashes"][str(s)], s
    assert "Error check PASSED" in p.stdout


@pytest.mark.gpu
def test_apps_on_several_strips(apps):
    """--gpus / --devices: every app gives the single-GPU hashes on row strips (C1 pipe incl. strip body force)."""
    case = [c for c in GOLD["b3_pipe_schedule"] if c["model"] == "FHP_I"][0]
    p = run_app("lgca-pipe", "--model", "FHP_I", "--steps", 500, "--hash-every", 500, "--quiet", "--devices", "0,0,0")
    assert hashes(p.stdout)[500] == case["hashes"]["500"]
    case = GOLD["b2_pure_stepping"][1]
    p = run_app("lgca-box", "-r", 255, "-c", 16, "--model", "FHP_II", "--steps", 200, "--pp-interval", 10, "--hash-every", 100,
                "--quiet", "--devices", "0,0,0,0")
    h = hashes(p.stdout)
    for s in (0, 100, 200):
        assert h[s] == case["hashes"][str(s)]
    import lgca_b200
    if lgca_b200.load_library().lgca_b200_device_count() >= 2:
        p = run_app("lgca-box", "-r", 255, "-c", 16, "--model", "FHP_II", "--steps", 200, "--pp-interval", 10, "--hash-every", 100,
                    "--quiet", "--gpus", 2)
        assert hashes(p.stdout)[200] == case["hashes"]["200"]


@pytest.mark.gpu
@pytest.mark.parametrize("app,extra,idx", [("lgca-periodic", ["-r", 255, "-c", 16], 2),
                                           ("lgca-box", ["-r", 255, "-c", 16, "--model", "FHP_II"], 1),
                                           ("lgca-box", ["-r", 127, "-c", 16, "--model", "FHP_I", "--bounce", "forward"], 4),
                                           ("lgca-box", ["-r", 127, "-c", 16, "--model", "HPP", "--bounce", "forward"], 5)])
def test_pure_stepping_apps(apps, app, extra, idx):
    case = GOLD["b2_pure_stepping"][idx]
    p = run_app(app, *extra, "--steps", 200, "--pp-interval", 10, "--hash-every", 100, "--quiet")
    h = hashes(p.stdout)
    for s in (0, 100, 200):
        assert h[s] == case["hashes"][str(s)], (app, s)
    assert "Error check PASSED" in p.stdout


@pytest.mark.gpu
def test_karman_diffusion_single_run(apps):
    for app, extra in (("lgca-karman", ["-r", 10, "-m", 0.2, "-c", 16, "--steps", 40]),
                       ("lgca-diffusion", ["--steps", 20]), ("lgca-single", ["--steps", 6])):
        p = run_app(app, *extra, "--quiet")
        assert "Error check PASSED" in p.stdout, app


def read_vti(path):
    raw = open(path, "rb").read()
    head, _, tail = raw.partition(b"<AppendedData encoding=\"raw\">")
    blob = tail[tail.index(b"_") + 1:]
    arrays = {}
    for m in re.finditer(rb'Name="([^"]+)" NumberOfComponents="(\d+)" format="appended" offset="(\d+)"', head):
        off = int(m.group(3))
        n = struct.unpack_from("<Q", blob, off)[0]
        arrays[m.group(1).decode()] = np.frombuffer(blob, np.float32, n // 4, off + 8)
    ext = re.search(rb'WholeExtent="0 (\d+) 0 (\d+) 0 0"', head)
    return arrays, (int(ext.group(1)), int(ext.group(2)))


@pytest.mark.gpu
def test_vti_output_matches_oracle(apps, tmp_path):
    from cpu_checkers import Oracle
    out = str(tmp_path) + "/"
    run_app("lgca-periodic", "-r", 255, "-c", 16, "--steps", 20, "--pp-interval", 10, "-w", 20, "-o", "vti", "--out-dir", out, "--quiet")
    o = Oracle("FHP_III", "periodic", 255.0, 0.2, 16)
    o.apply_bc("periodic")
    o.init("random")
    o.step(20)
    o.snapshot()
    o.post_process()
    cell, ext = read_vti(out + "cell_res_20.vti")
    assert ext == (256, 256)
    assert np.array_equal(cell["Cell density"], o.cell_density)
    assert np.array_equal(cell["Cell momentum"], o.cell_momentum)
    mean, ext = read_vti(out + "mean_res_20.vti")
    assert ext == (7, 7)
    assert np.array_equal(mean["Mean density"], o.mean_density)
    assert np.array_equal(mean["Mean momentum"], o.mean_momentum)


@pytest.mark.gpu
def test_png_output_matches_oracle_field(apps, tmp_path):
    """-o png (the viewers' OUTPUT_FORMAT, apps/karman/karman_viewer.h:87): res_<step>.png shows |mean momentum| of the
    snapshot over its own range through the blue->red table; compared per coarse cell with the oracle's field."""
    from cpu_checkers import Oracle
    from test_png_writer import read_png, vtk_rainbow
    out = str(tmp_path) + "/"
    run_app("lgca-periodic", "-r", 255, "-c", 16, "--steps", 20, "--pp-interval", 10, "-w", 20, "-o", "png", "--out-dir", out, "--quiet")
    o = Oracle("FHP_III", "periodic", 255.0, 0.2, 16)
    o.apply_bc("periodic")
    o.init("random")
    o.step(20)
    o.snapshot()
    o.post_process()
    img = read_png(out + "res_20.png")
    m = o.mean_momentum.reshape(8, 8, 2)
    mag = np.sqrt(m[..., 0] * m[..., 0] + m[..., 1] * m[..., 1]).astype(np.float32)
    assert np.array_equal(img, vtk_rainbow(mag, float(mag.min()), float(mag.max()))[::-1])
