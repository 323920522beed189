"""CPU-side check of the kernel LOGIC: uxsim_b200/csrc/engine.cu compiled with g++ against tests/emu/cuda_emu.h
(kernels run as loops over blocks/threads) must be bit-identical to the oracle.  This is a debugging aid for the
GPU-less container; the GPU parity tests proper are tests/test_gpu_parity.py.  The emulation library is test
infrastructure and is never loaded by the package."""
import numpy as np
import pytest

from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S

from parity import compare_engines, compare_logs, run_pair

CASES = [
    ("1link", S.straight_1link, None), ("spillback", S.straight_3link_spillback, None), ("signal", S.straight_2link_signal, None),
    ("lane_drop", S.lane_drop_2to1, None), ("3lane", S.multilane_3lane, None), ("merge", S.merge_y, None),
    ("dead_end", S.dead_end, None), ("short_chain", S.short_links_chain, None),
    ("merge_stochastic", lambda: S.merge_y(hard_deterministic_mode=False, random_seed=3), None),
    ("grid5", lambda: S.grid(n=5, seed=2, tmax=3000, demand_t=1000), None),
    ("grid5_chunked", lambda: S.grid(n=5, seed=9, tmax=3000, demand_t=1000), [300, 305, 1234, 3000]),
    ("grid6_signals", lambda: S.grid(n=6, seed=5, tmax=3000, demand_t=1200, signals=True), None),
    ("siouxfalls", lambda: S.sioux_falls(seed=1), None),
    ("dta_grid9", lambda: S.dta_grid9(seed=1), None),
    ("gradual", lambda: S.grid(n=5, seed=4, tmax=3000, demand_t=1000, route_choice_update_gradual=True), None),
]


@pytest.mark.parametrize("name,mk,until", CASES, ids=[c[0] for c in CASES])
def test_emulated_kernels_match_oracle(name, mk, until, emu_api, oracle_api):
    Ee, Eo = run_pair(mk(), emu_api, oracle_api, until=until)
    assert compare_engines(Ee, Eo) == []
    assert Ee.get_int(C.I_PLATOON_UPDATES) == Eo.get_int(C.I_PLATOON_UPDATES)


def test_emulated_chicago_levels(emu_api, oracle_api):
    """Chicago sketch (deltan=30) needs 4 transfer dependency levels; first 1500 s."""
    Ee, Eo = run_pair(S.chicago_sketch(seed=0), emu_api, oracle_api, until=[1500])
    assert Ee.get_int(C.I_TRANSFER_LEVELS) == 4
    assert compare_engines(Ee, Eo) == []


LOG_CASES = [("spillback", S.straight_3link_spillback), ("3lane", S.multilane_3lane), ("lane_drop", S.lane_drop_2to1), ("dead_end", S.dead_end),
             ("merge_stochastic", lambda: S.merge_y(hard_deterministic_mode=False, random_seed=3)),
             ("grid6_signals", lambda: S.grid(n=6, seed=5, tmax=3000, demand_t=1200, signals=True))]


@pytest.mark.parametrize("name,mk", LOG_CASES, ids=[c[0] for c in LOG_CASES])
def test_emulated_vehicle_logs_match_oracle(name, mk, emu_api, oracle_api):
    """Full per-platoon logs (log_t/x/v/s/lane/link/state, log_t_link, enter log), mid-run and at the end."""
    sc = mk()
    Ee, Eo = run_pair(sc, emu_api, oracle_api, until=[0.4 * sc.tmax()], vehicle_logging_timestep_interval=1)
    assert compare_logs(Ee, Eo) == []
    Ee.main_loop(), Eo.main_loop()
    assert compare_logs(Ee, Eo) == [] and compare_engines(Ee, Eo) == []


@pytest.mark.parametrize("name,mk", LOG_CASES[1:] + [("siouxfalls", lambda: S.sioux_falls(seed=1))], ids=[c[0] for c in LOG_CASES[1:]] + ["siouxfalls"])
def test_emulated_replica_lane_update(name, mk, emu_api, oracle_api, monkeypatch):
    """k_update_rl (thread per link and replica; what ensembles run) forced on a single scenario: same results, same logs."""
    monkeypatch.setenv("UXSIM_B200_UPDATE", "rl")
    sc = mk()
    Ee, Eo = run_pair(sc, emu_api, oracle_api, vehicle_logging_timestep_interval=1)
    assert compare_engines(Ee, Eo) == [] and compare_logs(Ee, Eo) == []


@pytest.mark.parametrize("name,mk", [("grid6_signals", lambda: S.grid(n=6, seed=5, tmax=3000, demand_t=1200, signals=True)), ("dta_grid9", lambda: S.dta_grid9(seed=1)),
                                     ("lane_drop", S.lane_drop_2to1)], ids=["grid6_signals", "dta_grid9", "lane_drop"])
def test_emulated_link_pre_kernel(name, mk, emu_api, oracle_api, monkeypatch):
    """k_link_pre (scalar link / node bookkeeping as its own launch; what large worlds run) forced on small scenarios."""
    monkeypatch.setenv("UXSIM_B200_PRE_SPLIT", "1")
    Ee, Eo = run_pair(mk(), emu_api, oracle_api, vehicle_logging_timestep_interval=1)
    assert compare_engines(Ee, Eo) == [] and compare_logs(Ee, Eo) == []


@pytest.mark.parametrize("name,mk", LOG_CASES[1:4] + [("grid6_signals", LOG_CASES[-1][1])], ids=[c[0] for c in LOG_CASES[1:4]] + ["grid6_signals"])
def test_emulated_update_lane_groups(name, mk, emu_api, oracle_api, monkeypatch):
    """k_update with two work items per warp (16-lane groups; what large, lightly loaded worlds run) forced on small scenarios."""
    monkeypatch.setenv("UXSIM_B200_UPDATE_GROUP", "16")
    Ee, Eo = run_pair(mk(), emu_api, oracle_api, vehicle_logging_timestep_interval=1)
    assert compare_engines(Ee, Eo) == [] and compare_logs(Ee, Eo) == []


RL_CASES = [("short_chain", S.short_links_chain), ("3lane", S.multilane_3lane), ("dead_end", S.dead_end), ("lane_drop", S.lane_drop_2to1),
            ("merge_stochastic", lambda: S.merge_y(hard_deterministic_mode=False, random_seed=3)),
            ("grid6_signals", lambda: S.grid(n=6, seed=5, tmax=3000, demand_t=1200, signals=True)), ("siouxfalls", lambda: S.sioux_falls(seed=1)),
            ("dta_grid9", lambda: S.dta_grid9(seed=1)), ("gradual", lambda: S.grid(n=5, seed=4, tmax=3000, demand_t=1000, route_choice_update_gradual=True))]


@pytest.mark.parametrize("name,mk", RL_CASES, ids=[c[0] for c in RL_CASES])
def test_emulated_replica_lane_node_kernel(name, mk, emu_api, oracle_api, monkeypatch):
    """k_link_pre<true> + k_node_rl (thread per node and replica; what ensembles of >= 32 replicas run) forced on single
    scenarios, together with the replica-lane link update: same results, same logs."""
    monkeypatch.setenv("UXSIM_B200_NODE", "rl")
    monkeypatch.setenv("UXSIM_B200_UPDATE", "rl")
    Ee, Eo = run_pair(mk(), emu_api, oracle_api, vehicle_logging_timestep_interval=1)
    assert compare_engines(Ee, Eo) == [] and compare_logs(Ee, Eo) == []


def test_emulated_replica_lane_node_kernel_levels_and_ragged_groups(emu_api, oracle_api, monkeypatch):
    """Chicago sketch (4 dependency levels) with 35 replicas = one full lane group + a ragged one; first 900 s; replicas 0, 31, 34."""
    sc = S.chicago_sketch(seed=0)
    Ee = sc.to_engine(emu_api, vehicle_logging_timestep_interval=-1, num_replicas=35)
    Ee.main_loop(until_t=900)
    assert Ee.get_int(C.I_TRANSFER_LEVELS) == 4
    for r in (0, 31, 34):
        Eo = sc.to_engine(oracle_api, vehicle_logging_timestep_interval=-1, random_seed=r)
        Eo.main_loop(until_t=900)
        assert compare_engines(Ee, Eo, replica_a=r) == [], f"replica {r}"


def test_all_replica_getter_matches_per_replica(emu_api):
    """get_array(replica=-1): [R, n] in one transfer == the per-replica getters."""
    from uxsim_b200 import _capi as C
    from uxsim_b200 import scenario as S

    E = S.grid(4, seed=1).to_engine(emu_api, vehicle_logging_timestep_interval=-1, num_replicas=3, random_seed=7)
    E.main_loop()
    for what in (C.A_VEH_ARRIVAL_TS, C.A_VEH_TRAVEL_TIME, C.A_LINK_STAT_COUNT, C.A_VEH_LINK, C.A_LINK_CAPACITY_IN_REMAIN):
        allr = E.get_array_all(what)
        assert allr.shape[0] == 3
        for r in range(3):
            assert np.array_equal(allr[r], E.get_array(what, r)), what
    E.close()
