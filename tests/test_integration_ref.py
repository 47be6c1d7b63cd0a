"""Drop-in proof against the reference's OWN headers: oracle/_ref/ref_b200_app is integration/ref_app.cpp +
integration/b200_lattice.h (the binding INTEGRATION.md shows) compiled by oracle/Makefile against the unmodified
/root/reference/src/{lattice.h,lgca_bitset.h,lgca_models.h} and linked with the reference's lattice.cpp -- the
reference's base class, BC painters, initialisers and forcing formulas drive the B200 backend through the C-ABI.
Built where /root/reference exists (the build container); the binary travels to the GPU box prebuilt."""
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "oracle", "_ref", "ref_b200_app")
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_b.json")))
REF_RUNS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_runs.json")))

needs_app = pytest.mark.skipif(not os.path.exists(APP), reason="oracle/_ref/ref_b200_app not built (needs /root/reference)")


def run(*args, check=True):
    p = subprocess.run([APP] + [str(a) for a in args], capture_output=True, text=True)
    if check:
        assert p.returncode == 0, (p.stdout + p.stderr)[-1500:]
    return p


def hashes(out):
    return {int(m.group(1)): m.group(2) for m in re.finditer(r"HASH step (\d+) ([0-9a-f]{16})", out)}


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_binding_compiles_against_reference_headers():
    """Rebuild from scratch: integration/b200_lattice.h must compile against the reference's unmodified headers."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    assert os.path.exists(APP)
    src = open(os.path.join(ROOT, "integration", "b200_lattice.h")).read()
    assert '#include "lattice.h"' in src and "lgca_b200/host" not in src.split("#ifndef")[1]   # only reference + C-ABI headers


@needs_app
def test_fails_loudly_without_gpu():
    import lgca_b200
    if lgca_b200.load_library().lgca_b200_device_count() > 0:
        pytest.skip("GPU present")
    p = run("collision", 7, 80, 0.2, 1, 3, 1, check=False)
    assert p.returncode != 0 and "ERROR in B200_Lattice" in p.stdout + p.stderr


@needs_app
@pytest.mark.gpu
def test_single_collision_trace_on_reference_base():
    """The 21 x 10 single-collision demo (apps/single): two particles meet head on and scatter by the site's chirality."""
    from cpu_checkers import Oracle
    case = GOLD["b1_single_collision"]
    o = Oracle(case["model"], *case["ctor"])
    o.apply_bc("pipe")
    o.init("single_collision")
    h = hashes(run("collision", 7, 80, 0.2, 1, 7, 1).stdout)
    assert h[0] == o.hash()
    for s in range(1, 8):
        o.step(1)
        assert h[s] == o.hash(), s


@needs_app
@pytest.mark.gpu
def test_c1_pipe_1000_steps_on_reference_base():
    case = [c for c in GOLD["b3_pipe_schedule"] if c["model"] == "FHP_I"][0]
    p = run("pipe", 6, 80, 0.3, 10, 1000, 500)
    h = hashes(p.stdout)
    for s in (0, 500, 1000):
        assert h[s] == case["hashes"][str(s)], s
    assert re.search(r"PARTICLES (\d+) \1\b", p.stdout)


@needs_app
@pytest.mark.gpu
def test_karman_default_on_reference_base():
    gold = REF_RUNS["karman_default"]
    import lgca_b200
    gpus = 2 if lgca_b200.load_library().lgca_b200_device_count() >= 2 else 1
    h = hashes(run("karman", 7, 80, 0.3, 20, 100, 5, gpus).stdout)
    for s in (0, 5, 100):
        assert h[s] == gold["hashes"][str(s)], s
