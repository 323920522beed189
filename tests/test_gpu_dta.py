"""DTA route swap on the GPU against the oracle: bit-exact routes, costs, gaps and counts."""
import numpy as np
import pytest

from uxsim_b200 import _capi as C
from test_dta_oracle import SCENARIOS, check_batched_swap_equals_per_replica, check_swap_against_oracle, check_swap_with_aborts, run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["diverge", "grid5", "siouxfalls", "dta_grid9"])
@pytest.mark.parametrize("mode", [0, 1])
def test_gpu_route_swap_matches_oracle(oracle_api, cuda_api, name, mode):
    moved = check_swap_against_oracle(oracle_api, cuda_api, name, mode)
    if name in ("siouxfalls", "dta_grid9"):
        assert moved > 0


def test_gpu_enumerate_matches_oracle(oracle_api, cuda_api):
    sc = SCENARIOS["siouxfalls"]()
    Eo, Eg = sc.to_engine(oracle_api), sc.to_engine(cuda_api)
    assert Eo.dta_enumerate_route_sets(8, 31337) == Eg.dta_enumerate_route_sets(8, 31337)
    for x, y in zip(Eo.dta_route_sets_csr(), Eg.dta_route_sets_csr()):
        assert np.array_equal(x, y)
    Eo.close(); Eg.close()


def test_gpu_route_swap_fixed_number_of_candidates(ref_api, cuda_api):
    """swap_num >= 0 (std::shuffle of the candidates): against the reference's own dta_solver.cpp when oracle/_ref is built."""
    Er, Eg = (run(a, "dta_grid9", hard_deterministic_mode=True) for a in (ref_api, cuda_api))
    a, b = Er.dta_route_swap(1, 0.0, 5, swap_num=40), Eg.dta_route_swap(1, 0.0, 5, swap_num=40)
    for k in ("rs_offsets", "rs_link_ids", "ra_offsets", "ra_link_ids", "cost_actual"):
        assert np.array_equal(a[k], b[k]), k
    assert (a["n_swap"], a["potential_n_swap"], a["total_t_gap"]) == (b["n_swap"], b["potential_n_swap"], b["total_t_gap"])
    Er.close(); Eg.close()


def test_gpu_route_swap_needs_recorded_traversals(cuda_api):
    sc = SCENARIOS["diverge"]()
    E = sc.to_engine(cuda_api)
    E.main_loop()
    with pytest.raises(RuntimeError, match="RECORD_TRAVERSALS"):
        E.dta_route_swap(0, 0.5, 1)
    E.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_gpu_batched_swap_equals_per_replica(cuda_api, mode):
    check_batched_swap_equals_per_replica(cuda_api, mode)


@pytest.mark.parametrize("mode", [0, 1])
def test_gpu_route_swap_with_aborts(oracle_api, cuda_api, mode):
    """Aborted platoons keep no route and draw nothing (dta_solver.cpp:168 / 308)."""
    check_swap_with_aborts(oracle_api, cuda_api, mode)
