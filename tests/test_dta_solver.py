"""Day-to-day DUE / DSO solvers over the engine (uxsim_b200/dta.py; counterparts of CppSolverDUE / CppSolverDSO_D2D).

The solver logic is exercised on CPU with the host-emulation build standing in for the GPU library (test-only
monkeypatch), and on the GPU with the real library."""
import random

import numpy as np
import pytest

from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S
from uxsim_b200.dta import CudaSolverDSO_D2D, CudaSolverDUE, SolverDUE


def _solve(cls, max_iter=6, **kw):
    random.seed(1)
    sol = cls(lambda: S.dta_grid9().to_world(print_mode=0, vehicle_logging_timestep_interval=-1))
    W = sol.solve(max_iter=max_iter, n_routes_per_od=6, swap_prob=0.2, print_progress=False, **kw)
    return sol, W


def _check_due(sol, W):
    assert len(sol.ttts) == len(sol.t_gaps) == len(sol.route_log) == 6
    assert sol.t_gaps[-1] < 0.7 * sol.t_gaps[0]          # the route cost gap closes
    assert all(n > 0 for n in sol.n_swaps) and all(p >= n for p, n in zip(sol.potential_swaps, sol.n_swaps))
    assert W is sol.W_sol and W.analyzer.trip_completed == W.analyzer.trip_all
    r = sol.route_log[-1]
    name, links = next(iter(r.items()))
    assert len(r) == len(W.VEHICLES) and all(l in W.LINKS_NAME_DICT for l in links)
    assert (W.get_node("n(0, 3)").name, W.get_node("n(8, 3)").name) in W.dict_od_to_routes or len(W.dict_od_to_routes) > 0


def _check_dso(sol, W):
    assert min(sol.ttts) < 0.9 * sol.ttts[0]              # total travel time falls towards the system optimum
    assert W is sol.W_sol and W.analyzer.total_travel_time == min(sol.ttts) or W.analyzer.trip_completed == W.analyzer.trip_all


@pytest.fixture
def emu_as_product(emu_api, monkeypatch):
    monkeypatch.setattr(C, "_product", emu_api)
    return emu_api


def test_due_solver_emulated(emu_as_product):
    _check_due(*_solve(CudaSolverDUE))


def test_dso_solver_emulated(emu_as_product):
    _check_dso(*_solve(CudaSolverDSO_D2D))


def test_solver_with_external_route_sets_emulated(emu_as_product):
    W0 = S.dta_grid9().to_world(print_mode=0, vehicle_logging_timestep_interval=-1)
    W0.finalize_scenario()
    W0._E.dta_enumerate_route_sets(3, 5)
    o, d, od_off, r_off, links = W0._E.dta_route_sets_csr()
    route_sets = {(W0.NODES[o[p]].name, W0.NODES[d[p]].name): [[W0.LINKS[i].name for i in links[r_off[r]:r_off[r + 1]]] for r in range(od_off[p], od_off[p + 1])]
                  for p in range(len(o)) if od_off[p + 1] > od_off[p]}
    sol, W = _solve(CudaSolverDUE, max_iter=3, route_sets=route_sets)
    assert len(sol.ttts) == 3 and sol.potential_swaps[0] > 0


def test_solver_rejects_foreign_worlds():
    with pytest.raises(ValueError, match="cuda=True"):
        SolverDUE(lambda: object()).solve(max_iter=1)


def _check_ensemble(api):
    from uxsim_b200.dta import DayToDayEnsemble

    sc = S.dta_grid9()
    ens = DayToDayEnsemble(sc, 3, mode="DUE", base_seed=10, api=api)
    ens.solve(max_iter=4, n_routes_per_od=6, swap_prob=0.2)
    assert ens.ttts.shape == (3, 4) and (ens.trips > 0).all()
    assert (ens.t_gaps[:, -1] < ens.t_gaps[:, 0]).all()          # every chain closes its gap
    assert len({tuple(ens.ttts[r]) for r in range(3)}) == 3       # chains are independent (different seeds, different routes)
    # chain r of the ensemble == a single-replica solve with seed base_seed + r (same swap seeds)
    solo = DayToDayEnsemble(sc, 1, mode="DUE", base_seed=10 + 1, api=api)
    solo.solve(max_iter=2, n_routes_per_od=6, swap_prob=0.2)
    assert solo.ttts[0, 0] == ens.ttts[1, 0] and solo.n_swaps[0, 0] != 0


def _check_rewind_equals_rebuild(api):
    """dta_next_day (rewind in place, routes handed over on the device) == a new world per day with the routes enforced
    through the host, which is what the reference does (and what the oracle-parity tests cover)."""
    from uxsim_b200.dta import DayToDayEnsemble

    sc, R, days = S.dta_grid9(), 3, 3
    ens = DayToDayEnsemble(sc, R, mode="DUE", base_seed=3, api=api)
    ens.solve(max_iter=days, n_routes_per_od=5, swap_prob=0.3, enum_seed=9, swap_seed=2)
    tab = sc.tables()
    E0 = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1)
    E0.dta_enumerate_route_sets(5, 9)
    prev = [None] * R
    for day in range(days):
        E = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1, num_replicas=R, random_seed=3)
        E.set_option(C.OPT_RECORD_TRAVERSALS, 1)
        E.dta_share_route_sets(E0)
        for r in range(R):
            if prev[r] is not None:
                E.dta_enforce_routes(r, *prev[r])
        E.main_loop()
        res = E.dta_route_swap_all(0, 0.3, (2 * 1000003 + day * 65537) & 0x3FFFFFFF)
        for r in range(R):
            prev[r] = (res[r]["rs_link_ids"], res[r]["rs_offsets"])
            tt = E.get_array(C.A_VEH_TRAVEL_TIME, r)
            assert float(tt[tt >= 0].sum()) * sc.deltan == ens.ttts[r, day]
            assert res[r]["n_swap"] == ens.n_swaps[r, day] and res[r]["potential_n_swap"] == ens.potential_swaps[r, day]
            assert res[r]["total_t_gap"] / len(tt) == ens.t_gaps[r, day]
        E.close()
    for r in range(R):
        assert np.array_equal(prev[r][0], ens.routes[r][0]) and np.array_equal(prev[r][1], ens.routes[r][1])
    E0.close()


def test_rewind_equals_rebuild_emulated(emu_api):
    _check_rewind_equals_rebuild(emu_api)


@pytest.mark.gpu
def test_rewind_equals_rebuild_gpu(cuda_api):
    _check_rewind_equals_rebuild(cuda_api)


def test_day_to_day_ensemble_emulated(emu_api):
    _check_ensemble(emu_api)


@pytest.mark.gpu
def test_day_to_day_ensemble_gpu(cuda_api):
    _check_ensemble(cuda_api)


@pytest.mark.gpu
def test_due_solver_gpu(cuda_api):
    _check_due(*_solve(CudaSolverDUE))


@pytest.mark.gpu
def test_dso_solver_gpu(cuda_api):
    _check_dso(*_solve(CudaSolverDSO_D2D))
