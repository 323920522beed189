"""SURVEY.md 8f-4 features on the GPU: every case of tests/feature_cases.py (links_prefer / links_avoid, specified routes, toll
timeseries, route-choice penalty, adjust_node_capacity, multi-entry signal groups, mid-run setters between chunks incl. signal
plans and groups that change length) on the CUDA engine against the oracle -- bit-exact, per-platoon logs included."""
import pytest

from feature_cases import CASES
from parity import compare_engines, compare_logs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,runner,det", CASES, ids=[c[0] for c in CASES])
def test_feature_gpu_matches_oracle(name, runner, det, cuda_api, oracle_api):
    Eg, Eo = runner(cuda_api, vehicle_logging_timestep_interval=1), runner(oracle_api, vehicle_logging_timestep_interval=1)
    assert compare_engines(Eg, Eo) == [] and compare_logs(Eg, Eo) == []
    Eg.close(); Eo.close()


@pytest.mark.parametrize("name", ["links_avoid", "toll_timeseries", "midrun_setters"])
def test_feature_gpu_replicas(name, cuda_api, oracle_api):
    """The same features in an ensemble: replica r == the single-seed oracle run with seed + r."""
    runner = dict((c[0], c[1]) for c in CASES)[name]
    Eg = runner(cuda_api, vehicle_logging_timestep_interval=-1, num_replicas=3)
    seed0 = Eg.params.random_seed
    for r in range(3):
        Eo = runner(oracle_api, vehicle_logging_timestep_interval=-1, random_seed=seed0 + r)
        assert compare_engines(Eg, Eo, replica_a=r) == [], f"replica {r}"
        Eo.close()
    Eg.close()
