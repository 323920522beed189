"""TEST INFRASTRUCTURE -- pytest plugin that runs the REFERENCE'S OWN, UNMODIFIED test files against this backend.

The reference switches its suite to the C++ engine with `pytest --cpp` (/root/reference/tests/conftest.py:13-47: World.__new__
is patched so that `cpp` defaults to True).  This plugin is the same switch for the CUDA backend: every `World(...)` a reference
test creates is served by `uxsim_b200.World(..., cuda=True)`, and the test's own assertions (known-answer numbers with the
reference's `equal_tolerance`) decide.  Used by tests/test_reference_suite.py:

    PYTHONPATH=tests:oracle python -m pytest -p ref_suite_plugin <reference test file>

  UXSIM_B200_SUITE_LIB=emu      the host emulation of the kernels (CPU container; tests/emu)
  UXSIM_B200_SUITE_LIB=product  libuxsim_b200.so on cuda:0 (GPU box; the reference package and test files staged under baseline/_ref)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "flaky: rerun marker of the reference suite (pytest-rerunfailures is not installed; ignored)")
    from ref_python_shim import import_reference
    import types

    for name in ("nbformat", "nbconvert", "nbconvert.preprocessors"):  # imported at module level by the reference's test_cpp_mode.py for its
        try:                                                            # notebook tests (deselected here); absent from this image
            __import__(name)
        except ImportError:
            m = types.ModuleType(name)
            m.ExecutePreprocessor = m.CellExecutionError = type("Unavailable", (Exception,), {})
            sys.modules[name] = m

    ux = import_reference()  # the reference package (matplotlib stubbed), so that `from uxsim import *` in its tests works
    import uxsim_b200.world as wmod

    if os.environ.get("UXSIM_B200_SUITE_LIB", "product") == "emu":
        from uxsim_b200._capi import CApi

        emu = CApi(os.path.join(ROOT, "tests", "emu", "_build", "libuxsim_emu.so"), "ux_")
        wmod.C.load_product = lambda: emu
    import uxsim.uxsim as mod

    def _cuda_new(cls, *args, cpp=False, **kwargs):  # World.__new__ returning another class's instance skips World.__init__
        kwargs.pop("cuda", None)
        return wmod.CudaWorld(*args, **kwargs)

    mod.World.__new__ = _cuda_new
    ux.World = mod.World
    # the day-to-day solvers: the reference hands `SolverDUE(func_World, cpp=True)` to its C++ adapter (DTAsolvers.py:21-26, 378-383);
    # here every SolverDUE / SolverDSO_D2D is this package's counterpart of that adapter
    import uxsim.DTAsolvers.DTAsolvers as dta_mod
    import uxsim_b200.dta as cuda_dta

    def _solver_new(cuda_cls):
        def _new(cls, func_World, cpp=False):
            return cuda_cls(func_World)
        return _new

    dta_mod.SolverDUE.__new__ = _solver_new(cuda_dta.CudaSolverDUE)
    dta_mod.SolverDSO_D2D.__new__ = _solver_new(cuda_dta.CudaSolverDSO_D2D)
