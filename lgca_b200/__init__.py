"""lgca_b200 -- B200-native lattice-gas cellular automaton engine (hand-written sm_100a CUDA behind a C-ABI).

The product is ``liblgca_b200.so`` (sources in ``lgca_b200/csrc``, C-ABI in ``include/lgca_b200.h``) and the
C++ ``B200_Lattice<Model>`` host class in ``lgca_b200/host``.  This Python package is only the thin ctypes
binding used by the tests and ``bench.py``; it never computes anything itself and has no CPU fallback.
"""
from .capi import Engine, Group, LgcaError, MODELS, library_path, load_library  # noqa: F401

__all__ = ["Engine", "Group", "LgcaError", "MODELS", "library_path", "load_library"]
