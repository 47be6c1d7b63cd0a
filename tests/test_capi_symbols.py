"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/lgca_b200.h declares; without a GPU compute entry points fail loudly (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from lgca_b200.build import build_library
    build_library()
    import lgca_b200
    return lgca_b200.load_library()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lgca_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lgca_b200_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_lists_every_declared_symbol():
    from lgca_b200 import capi
    assert sorted(capi.SYMBOLS) == declared_symbols()


def test_version(lib):
    assert lib.lgca_b200_version() == 1


def test_no_cpu_fallback(lib):
    import lgca_b200
    if lib.lgca_b200_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(lgca_b200.LgcaError, match="no CUDA device"):
        lgca_b200.Engine("FHP_II", 64, 64)


def test_product_never_imports_oracle():
    """The product tree must not reference the oracle / reference build (SPEC: parity-void otherwise)."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "lgca_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"lgca_oracle|liblgca_ref|oracle/|cpu_checkers", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
