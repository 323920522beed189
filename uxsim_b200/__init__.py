"""uxsim_b200 -- a B200-native backend for UXsim's per-timestep simulation loop (World.exec_simulation).

Hand-written sm_100a CUDA kernels behind a C-ABI (include/uxsim_b200.h), driven from Python through ctypes
and selected the way the reference selects its C++ mode: ``World(..., cuda=True)``.
There is no CPU fallback: without ``uxsim_b200/libuxsim_b200.so`` (built by ``__graft_entry__.build()``)
creating a world raises ImportError.
"""
from . import scenario  # noqa: F401
from ._capi import CApi, EngineWorld, load_product  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name in ("World", "CudaWorld", "CudaNode", "CudaLink", "CudaVehicle"):
        from . import world

        return getattr(world, name)
    raise AttributeError(name)
