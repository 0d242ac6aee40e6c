"""sph-bvf_b200 -- B200-native SPH-BVF timestep (USER-SSA-TSDPD hot path) behind a C ABI.

The product is libsphbvf.so (hand-written sm_100a CUDA + extern "C" layer, include/sphbvf.h) and
the LAMMPS-style "/cuda" host classes under lammps/.  This Python package is only a thin ctypes
binding used by the tests and bench.py; it contains no arithmetic and no CPU fallback -- if the
shared library is missing, or there is no CUDA device, it raises.

The directory name carries a hyphen (it mirrors the reference repository's name), so import it
with `load_package()` from tests/conftest.py / bench.py, or via importlib by path.
"""
from .capi import Engine, SphbvfError, config_from_meta, lib, FIELDS  # noqa: F401
