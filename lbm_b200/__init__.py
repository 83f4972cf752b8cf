"""lbm_b200 -- B200-native lattice-Boltzmann time step behind the reference's solver interface.

The compute path is the CUDA library `liblbm_b200.so` (built from lbm_b200/csrc for sm_100a) reached through
the C ABI declared in include/lbm_b200.h.  There is no CPU fallback: if the library is missing or no CUDA
device is present, creating a solver raises.
"""
from .capi import (BGK, FAST, FP32, FP64, MRT, STRICT, TRT, HostBuffer, LbmB200Error, Solver, abi_symbols, build, library_path,
                   load_library, mrt_moment_kinds, mrt_rates)

__all__ = ["Solver", "HostBuffer", "LbmB200Error", "build", "load_library", "library_path", "abi_symbols",
           "BGK", "TRT", "MRT", "FP64", "FP32", "STRICT", "FAST", "mrt_moment_kinds", "mrt_rates"]
