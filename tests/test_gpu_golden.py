"""The CUDA engine against golden vectors produced by the REFERENCE itself (no oracle involved)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import PHILOX_CASES, PYTHON_CASES  # noqa: E402
from golden_check import check_against_golden  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,mk", PYTHON_CASES, ids=[n for n, _ in PYTHON_CASES])
def test_cuda_matches_reference_python_engine(name, mk, cuda_api):
    E = mk().to_engine(cuda_api, vehicle_logging_timestep_interval=-1)
    E.main_loop()
    check_against_golden(E, "python", name)


@pytest.mark.parametrize("name,mk", PHILOX_CASES, ids=[n for n, _ in PHILOX_CASES])
def test_cuda_matches_reference_cpp_engine_with_philox(name, mk, cuda_api):
    E = mk().to_engine(cuda_api, vehicle_logging_timestep_interval=-1)
    E.main_loop()
    check_against_golden(E, "philox", name)
