"""SURVEY.md 8f-4 features (tests/feature_cases.py) on the CPU: the host emulation of the kernels against the oracle (every
case, bit-exact incl. per-platoon logs), and the oracle against the reference's own C++ engine on the cases without random
draws.  The same cases run on the CUDA engine in tests/test_gpu_features.py."""
import numpy as np
import pytest

from uxsim_b200 import _capi as C

from feature_cases import CASES
from parity import EXACT_FIELDS, compare_engines, compare_logs


@pytest.mark.parametrize("name,runner,det", CASES, ids=[c[0] for c in CASES])
def test_feature_emulated_kernels_match_oracle(name, runner, det, emu_api, oracle_api):
    Ee, Eo = runner(emu_api, vehicle_logging_timestep_interval=1), runner(oracle_api, vehicle_logging_timestep_interval=1)
    assert compare_engines(Ee, Eo) == [] and compare_logs(Ee, Eo) == []
    Ee.close(); Eo.close()


@pytest.mark.parametrize("name,runner,det", [c for c in CASES if c[2]], ids=[c[0] for c in CASES if c[2]])
def test_feature_oracle_matches_reference(name, runner, det, oracle_api, ref_api):
    Eo, Er = runner(oracle_api, vehicle_logging_timestep_interval=1), runner(ref_api, vehicle_logging_timestep_interval=1)
    # as tests/test_oracle_golden.py: everything but the route tables (the edge noise comes from different random streams)
    fields = [f for f in EXACT_FIELDS if f[0] not in ("route_pref", "route_next", "link_num_vehicles")]
    # veh_distance: the oracle sums per link traversal, the reference per step (documented deviation R4 of the oracle) -- same
    # value up to the last bits once a speed change makes the steps non-representable
    bad = [nm for nm, what in fields if nm != "veh_distance" and not np.array_equal(Eo.get_array(what), Er.get_array(what))]
    assert np.allclose(Eo.get_array(C.A_VEH_DISTANCE), Er.get_array(C.A_VEH_DISTANCE), rtol=1e-12, atol=1e-9)
    assert bad == [] and compare_logs(Eo, Er) == []
    Eo.close(); Er.close()
