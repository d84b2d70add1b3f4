"""pnp-b200: B200-native (sm_100a, fp64) EAFE assembly + BSR-AMG/FGMRES Newton step behind the
C ABI of include/pnpb200/*.h. This package holds the CUDA sources (csrc/), the C++ host-side
mirror of the reference's Linear_PNP/PDE interface (host/) and this ctypes harness binding."""
from . import capi  # noqa: F401
from .capi import PNP, default_params, lib  # noqa: F401
