"""In-tree build of liblgca_b200.so (hand-written CUDA for sm_100a + the C-ABI).

`python -m lgca_b200.build` or `lgca_b200.build.build_library()`; nvcc cross-compiles without a GPU.
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblgca_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-O3",
    "--use_fast_math",  # no float math on the hot path; the post kernels use explicit _rn intrinsics
]
OBJ_DIR = os.path.join(HERE, "build")
# the wavefront kernel's variants are spread over six translation units compiled in parallel (-DLGCA_WAVE_TU=n)
WAVE_TUS = 6


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _units():
    """(source, extra defines, object name) of every translation unit."""
    out = []
    for src in sources():
        stem = os.path.splitext(os.path.basename(src))[0]
        if stem == "lgca_step_wave":
            out += [(src, ["-DLGCA_WAVE_TU=%d" % i], "%s_tu%d.o" % (stem, i)) for i in range(WAVE_TUS)]
        else:
            out.append((src, [], stem + ".o"))
    return out


def build_library(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    extra = os.environ.get("LGCA_B200_NVCC_EXTRA", "").split()  # A/B experiments only
    out = os.environ.get("LGCA_B200_OUT", LIB)
    obj_dir = OBJ_DIR if out == LIB else out + ".build"
    os.makedirs(obj_dir, exist_ok=True)
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    nvcc = _nvcc()
    pv = ["-Xptxas", "-v"] if verbose else []

    def compile_unit(u):
        src, defs, obj = u
        cmd = [nvcc] + NVCC_FLAGS + extra + defs + pv + ["-c", "-o", os.path.join(obj_dir, obj), src]
        return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)

    units = _units()
    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_unit, units))
    failed = False
    for res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout)
        failed |= res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building liblgca_b200.so")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + [os.path.join(obj_dir, u[2]) for u in units]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed linking liblgca_b200.so")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
