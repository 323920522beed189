"""The reference-facing API: uxsim_b200.World(..., cuda=True) mirrors CppWorld (uxsim_cpp_wrapper.py)."""
import inspect

import numpy as np
import pytest

import uxsim_b200
from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S
from uxsim_b200.world import CudaLink, CudaNode, CudaWorld


def test_wrapper_signatures_match_class_constructors():
    """tests/test_wrapper_functions.py:10-35 of the reference: addNode/addLink take what the classes take."""
    for fn, cls in ((CudaWorld.addNode, CudaNode), (CudaWorld.addLink, CudaLink)):
        a = list(inspect.signature(fn).parameters)[1:]
        b = list(inspect.signature(cls.__init__).parameters)[2:]
        assert a == b


def test_reference_kwargs_are_accepted():
    """Every World(...) kwarg of the reference (uxsim.py:1677-1690) exists on CudaWorld with the same default."""
    ref = dict(name="", deltan=5, reaction_time=1, duo_update_time=600, duo_update_weight=0.5, duo_noise=0.01,
               route_choice_principle="homogeneous_DUO", route_choice_update_gradual=False, instantaneous_TT_timestep_interval=5,
               no_cyclic_routing=False, adjust_node_capacity=False, eular_dt=120, eular_dx=100, random_seed=None, print_mode=1,
               save_mode=1, show_mode=0, show_progress=1, show_progress_deltat=600, tmax=None, vehicle_logging_timestep_interval=1,
               reduce_memory_delete_vehicle_route_pref=False, hard_deterministic_mode=False, user_attribute=None, user_function=None, threads=1)
    sig = inspect.signature(CudaWorld.__init__).parameters
    for k, v in ref.items():
        assert k in sig and sig[k].default == v, k


def test_only_cuda_mode_is_offered():
    with pytest.raises(NotImplementedError):
        uxsim_b200.World(name="x")


@pytest.mark.gpu
def test_readme_merge_example(cuda_api, oracle_api):
    """README Y-merge through the World API == the same scenario through the raw C-ABI of the oracle."""
    W = uxsim_b200.World(cuda=True, name="", deltan=5, tmax=1200, print_mode=0, save_mode=0, show_mode=0, random_seed=0)
    W.addNode("orig1", 0, 0), W.addNode("orig2", 0, 2), W.addNode("merge", 1, 1), W.addNode("dest", 2, 1)
    W.addLink("link1", "orig1", "merge", length=1000, free_flow_speed=20, number_of_lanes=1)
    W.addLink("link2", "orig2", "merge", length=1000, free_flow_speed=20, number_of_lanes=1)
    W.addLink("link3", "merge", "dest", length=1000, free_flow_speed=20, number_of_lanes=1)
    W.adddemand("orig1", "dest", 0, 1000, 0.45)
    W.adddemand("orig2", "dest", 400, 1000, 0.6)
    assert W.exec_simulation() == 1
    assert len(W.VEHICLES) == 162 and W.TSIZE == 240 and not W.check_simulation_ongoing()
    sc = S.merge_y(hard_deterministic_mode=False, random_seed=0, tmax=1200)
    Eo = sc.to_engine(oracle_api, vehicle_logging_timestep_interval=-1)
    Eo.main_loop()
    L, T = 3, 240
    assert np.array_equal(W.get_link("link3").cum_arrival, Eo.get_array(C.A_CUM_ARRIVAL).reshape(L, T)[2])
    assert np.array_equal(W.get_link("link1").cum_departure, Eo.get_array(C.A_CUM_DEPARTURE).reshape(L, T)[0])
    tt = np.array([v.travel_time for v in W.VEHICLES.values()], dtype=float)
    assert np.array_equal(tt, Eo.get_array(C.A_VEH_TRAVEL_TIME))
    a = W.analyzer
    assert a.trip_all == 810 and 0 < a.trip_completed <= 810 and a.average_travel_time > 100


@pytest.mark.gpu
def test_chunked_exec_and_interventions(cuda_api, oracle_api):
    """exec_simulation(until_t=...) returns 0 until the horizon; signals can be changed between chunks
    (uxsim_cpp_wrapper.py:89-124; test_cpp_mode.py:8485-8643) -- and the result equals the oracle's under the same interventions."""
    W = uxsim_b200.World(cuda=True, deltan=5, tmax=2000, print_mode=0, save_mode=0, random_seed=0)
    W.addNode("o", 0, 0), W.addNode("m", 1, 0, signal=[60, 60]), W.addNode("d", 2, 0)
    l1 = W.addLink("l1", "o", "m", length=1000, signal_group=0)
    W.addLink("l2", "m", "d", length=1000)
    W.adddemand("o", "d", 0, 1000, 0.4)
    assert W.exec_simulation(until_t=500) == 0
    assert W.T == 101 and W.check_simulation_ongoing()
    n_mid = l1.num_vehicles
    assert n_mid >= 0
    W.get_node("m").signal = [30, 90]
    assert W.exec_simulation() == 1
    sc = S.Scenario("sig", S._world(deltan=5, tmax=2000, random_seed=0))
    sc.addNode("o", 0, 0), sc.addNode("m", 1, 0, signal=[60, 60]), sc.addNode("d", 2, 0)
    sc.addLink("l1", "o", "m", length=1000, signal_group=0), sc.addLink("l2", "m", "d", length=1000)
    sc.adddemand("o", "d", 0, 1000, 0.4)
    Eo = sc.to_engine(oracle_api, vehicle_logging_timestep_interval=-1)
    Eo.main_loop(until_t=500)
    Eo.set_node_signal(1, [30.0, 90.0])
    Eo.main_loop()
    L, T = 2, Eo.get_int(C.I_TOTAL_TIMESTEPS)
    for k, l in enumerate(W.LINKS):
        assert np.array_equal(l.cum_departure, Eo.get_array(C.A_CUM_DEPARTURE).reshape(L, T)[k])
        assert np.array_equal(l.cum_arrival, Eo.get_array(C.A_CUM_ARRIVAL).reshape(L, T)[k])
        assert np.array_equal(l.traveltime_actual, Eo.get_array(C.A_TRAVELTIME_ACTUAL).reshape(L, T)[k])
    assert np.array_equal(W.get_node("m").signal_log, Eo.get_array(C.A_SIGNAL_LOG).reshape(3, T)[1])
    assert l1.cum_departure[-1] == 80 * 5
    assert len(W.VEHICLES_RUNNING) == (W._veh("state") == C.VS_RUN).sum()


@pytest.mark.gpu
def test_unsupported_like_cpp_mode(cuda_api):
    W = uxsim_b200.World(cuda=True, deltan=5, tmax=600, print_mode=0, save_mode=0)
    W.addNode("a", 0, 0), W.addNode("b", 1, 0)
    W.addLink("l", "a", "b", length=500)
    with pytest.raises(NotImplementedError):
        W.addVehicle("a", "b", 0, mode="taxi")
    for f in (W.copy, lambda: W.save("x"), lambda: W.save_scenario("x")):  # (load_scenario IS supported here: tests/test_load_scenario.py)
        with pytest.raises(NotImplementedError):
            f()
    with pytest.raises(ValueError):
        W.addNode("a", 0, 0)


@pytest.mark.gpu
def test_vehicle_log_properties(cuda_api):
    """CudaVehicle.log_* mirror CppVehicle (uxsim_cpp_wrapper.py:575-664); the chunked-exec known answer of
    test_verification_straight_road.py:940-989: after 1000 s a platoon logged x=1600 at t=995 and sits at x=1800... here: free flow."""
    W = uxsim_b200.World(cuda=True, deltan=5, tmax=1200, print_mode=0, save_mode=0, random_seed=0)
    W.addNode("o", 0, 0), W.addNode("d", 1, 0)
    link = W.addLink("l", "o", "d", length=2000, free_flow_speed=20)
    veh = W.addVehicle("o", "d", 100)
    W.exec_simulation()
    assert veh.state == "end" and veh.travel_time == 100.0
    assert len(veh.log_t) == len(veh.log_x) == len(veh.log_state) == len(veh.log_link)
    run = [i for i, s in enumerate(veh.log_state) if s == "run"]
    assert veh.log_x[run[0]] == 0 and all(veh.log_v[i] == 20.0 for i in run)
    assert veh.log_link[run[0]] is link and veh.log_state[-1] == "end" and veh.log_state[0] == "home"
    assert [x[1] for x in veh.log_t_link] == ["home", link, "end"]
    route, ts = veh.traveled_route()
    assert route.links == [link] and len(ts) == 2
    assert list(link.vehicles_enter_log.values()) == [veh]


@pytest.mark.gpu
def test_analyzer_edie_and_od_through_the_world_api(cuda_api):
    """BasicAnalyzer (used when the reference package is not installed): Edie k/q/v grids from the device and OD statistics
    agree with each other and with the link curves."""
    from uxsim_b200.analysis import BasicAnalyzer

    W = S.straight_3link_spillback().to_world(print_mode=0)
    W.exec_simulation()
    A = BasicAnalyzer(W)
    A.basic_analysis()
    A.compute_edie_state()
    for l in W.LINKS:
        assert l.k_mat.shape == l.q_mat.shape == l.v_mat.shape and l.k_mat.shape[1] == int(l.length / l.edie_dx)
        assert np.isfinite(l.v_mat).all() and (l.v_mat <= l.u * (1 + 1e-9)).all()
        # vehicle-metres in the grid ~ vehicles that traversed the link x its length (up to the cells cut off at the end)
        assert l.dn_mat.sum() <= l.cum_arrival[-1] * l.length * (1 + 1e-9)
        assert l.dn_mat.sum() > 0.5 * l.cum_departure[-1] * l.length
    od = A.od_analysis()
    assert od["total_trips"].sum() == A.trip_all and od["completed_trips"].sum() == A.trip_completed
    assert np.isclose((od["completed_trips"] * np.where(od["completed_trips"] > 0, od["average_travel_time"], 0)).sum(), A.total_travel_time)


def test_lazy_vehicle_registry_over_emulated_kernels(emu_api, monkeypatch):
    """Host logic of the wrapper on the CPU (kernels emulated, test infrastructure only): ``adddemand`` records batches, no
    proxy object exists until somebody asks for a platoon, and the name -> vehicle mappings behave like the reference's
    OrderedDicts (creation order, renamed vehicles, LIVING / RUNNING subsets)."""
    import uxsim_b200.world as wmod
    from uxsim_b200 import World

    monkeypatch.setattr(wmod.C, "load_product", lambda: emu_api)
    W = World(name="", deltan=5, tmax=1200, print_mode=0, random_seed=0, cuda=True)
    W.addNode("orig1", 0, 0); W.addNode("orig2", 0, 2); W.addNode("merge", 1, 1); W.addNode("dest", 2, 1)
    W.addLink("link1", "orig1", "merge", length=1000, free_flow_speed=20, merge_priority=0.5)
    W.addLink("link2", "orig2", "merge", length=1000, free_flow_speed=20, merge_priority=2)
    W.addLink("link3", "merge", "dest", length=1000, free_flow_speed=20)
    W.adddemand("orig1", "dest", 0, 1000, 0.4); W.adddemand("orig2", "dest", 500, 1000, 0.6)
    special = W.addVehicle("orig1", "dest", 30, name="special")
    n = len(W.VEHICLES)
    assert n == 80 + 60 + 1 and len(W._veh_by_index._made) == 1  # only the vehicle handed back by addVehicle exists
    assert list(W.VEHICLES)[:3] == ["0", "1", "2"] and list(W.VEHICLES)[-1] == "special"
    assert "special" in W.VEHICLES and str(n - 1) not in W.VEHICLES and "17" in W.VEHICLES and "0017" not in W.VEHICLES
    assert W.VEHICLES["special"] is special and special.id == n - 1
    with pytest.raises(ValueError):
        W.addVehicle("orig1", "dest", 40, name="17")
    n = len(W.VEHICLES)  # the rejected vehicle keeps its default name, as in the reference (it is created before the check)
    v17 = W.VEHICLES["17"]
    assert v17.orig.name == "orig1" and v17.dest.name == "dest" and W.VEHICLES["100"].orig.name == "orig2"
    v17.user_attribute = {"tag": 1}
    assert W.VEHICLES["17"] is v17 and W._veh_by_index[17] is v17 and len(W._veh_by_index._made) <= 4
    W.exec_simulation(until_t=300)
    running = W.VEHICLES_RUNNING
    assert 0 < len(running) < len(W.VEHICLES_LIVING) <= n
    assert all(v.state == "run" for v in running.values()) and all(name in W.VEHICLES for name in running)
    W.exec_simulation()
    assert len(W.VEHICLES_RUNNING) == int((W._veh("state") == C.VS_RUN).sum()) < len(running)
    tt = np.array([v.travel_time for v in W.VEHICLES.values()], dtype=float)
    assert len(tt) == n and (tt[tt >= 0] >= 100).all()
    assert [v.name for v in W._veh_by_index][:2] == ["0", "1"] and W._veh_by_index.names()[-2] == "special"


def test_link_and_vehicle_measures_match_reference_wrapper_formulas(emu_api, monkeypatch):
    """The derived measures the reference's wrapper offers (uxsim_cpp_wrapper.py:386-445, 564-566, 695-718) over the emulated kernels."""
    import uxsim_b200.world as wmod

    monkeypatch.setattr(wmod.C, "load_product", lambda: emu_api)
    W = wmod.CudaWorld(name="", deltan=5, tmax=1200, print_mode=0, save_mode=0, random_seed=0)
    W.addNode("o", 0, 0); W.addNode("m", 1, 0); W.addNode("d", 2, 0)
    l1 = W.addLink("l1", "o", "m", length=1000, free_flow_speed=20, jam_density=0.2)
    l2 = W.addLink("l2", "m", "d", length=1000, free_flow_speed=10, jam_density=0.2)
    W.adddemand("o", "d", 0, 500, 0.5)
    v0 = W.VEHICLES["0"]
    v0.add_dests(["m", "d"])
    assert [n.name for n in v0.dest_list] == ["m", "d"]
    W.exec_simulation(until_t=400)
    assert l2.num_vehicles_queue >= 0 and l1.num_vehicles_queue % W.DELTAN == 0
    assert v0.link_arrival_time >= 0.0
    W.exec_simulation()
    assert l1.inflow_between(0, 500) == (l1.arrival_count(500) - l1.arrival_count(0)) / 500
    assert l1.average_density(300) == l1.num_vehicles_t(300) / l1.length
    assert l1.average_flow(300) == l1.average_density(300) * l1.average_speed(300)
    assert l2.average_travel_time_between(0, 10) >= l2.length / l2.u - 1e-9
    assert set(W.ROUTECHOICE.dist_record) == set(range(0, W.TSIZE, W.DELTAT_ROUTE))
    first, last = W.ROUTECHOICE.dist_record[0], W.ROUTECHOICE.dist_record[max(W.ROUTECHOICE.dist_record)]
    assert first.shape == last.shape == (3, 3) and 150.0 <= first[0, 2] <= 150.0 * 1.01  # free flow (+ route choice noise) at the first update
    assert np.array_equal(last, W.ROUTECHOICE.dist)


def _override_signal_pair(wmod):
    """test_verification_node.py::test_override_signal with an array-aware final comparison (the reference's `cum1 == cum2` is
    ambiguous on numpy arrays, in its own C++ mode too): signal plan + groups given at construction == given through override_signal."""
    out = []
    for override in (False, True):
        W = wmod.CudaWorld(name="", deltan=1, tmax=1200, print_mode=0, save_mode=0, random_seed=0)
        for name, x, y in (("orig1", 0, 0), ("orig2", 0, 2), ("orig3", 0, 3), ("dest", 2, 1)):
            W.addNode(name, x, y)
        merge = W.addNode("merge", 1, 1) if override else W.addNode("merge", 1, 1, signal=[30, 60])
        grp = {"link1": [0], "link2": [0, 1], "link2b": [1]}
        for name, o, mp in (("link1", "orig1", 1), ("link2", "orig2", 1), ("link2b", "orig3", 0.5)):
            kw = {} if override else {"signal_group": grp[name]}
            W.addLink(name, o, "merge", length=1000, free_flow_speed=20, jam_density=0.2, merge_priority=mp, **kw)
        W.addLink("link3", "merge", "dest", length=1000, free_flow_speed=20, jam_density=0.2, capacity_in=1)
        W.adddemand("orig1", "dest", 0, 1000, 0.2); W.adddemand("orig2", "dest", 500, 1000, 0.3); W.adddemand("orig3", "dest", 500, 1000, 0.3)
        if override:
            merge.override_signal(signal=[30, 60], groups=[["link1", "link2"], ["link2", "link2b"]])
        W.exec_simulation()
        W.analyzer.basic_analysis()
        out.append((W.analyzer.total_travel_time, [np.array(l.cum_arrival) for l in W.LINKS], list(W.get_node("merge").signal_log)))
    return out


def test_override_signal_equals_construction_time_plan_emulated(emu_api, monkeypatch):
    import uxsim_b200.world as wmod

    monkeypatch.setattr(wmod.C, "load_product", lambda: emu_api)
    a, b = _override_signal_pair(wmod)
    assert a[0] == b[0] and a[2] == b[2] and all(np.array_equal(x, y) for x, y in zip(a[1], b[1]))
    assert len(set(a[2])) == 2  # the signal did switch phases


@pytest.mark.gpu
def test_override_signal_equals_construction_time_plan(cuda_api):
    import uxsim_b200.world as wmod

    a, b = _override_signal_pair(wmod)
    assert a[0] == b[0] and a[2] == b[2] and all(np.array_equal(x, y) for x, y in zip(a[1], b[1]))


def _history_and_abort_checks(wmod):
    def build():
        W = wmod.CudaWorld(name="", deltan=5, tmax=2400, print_mode=0, save_mode=0, random_seed=3, duo_update_time=300)
        for name, x, y in (("o", 0, 0), ("a", 1, 1), ("b", 1, -1), ("d", 2, 0)):
            W.addNode(name, x, y)
        W.addLink("oa", "o", "a", length=1000, free_flow_speed=20); W.addLink("ad", "a", "d", length=1000, free_flow_speed=20, capacity_out=0.3)
        W.addLink("ob", "o", "b", length=1500, free_flow_speed=20); W.addLink("bd", "b", "d", length=1500, free_flow_speed=20)
        W.adddemand("o", "d", 0, 1500, 0.6)
        return W
    # (1) RouteChoice.dist_record: the table of the FIRST route update, read after the whole run, equals the current table of a world
    #     that has only run its first step; the latest entry equals RouteChoice.dist; tables are all-pairs (a, b are no destinations)
    A, B = build(), build()
    A.exec_simulation()
    B.exec_simulation(until_t=B.DELTAT)
    rec = A.ROUTECHOICE.dist_record
    assert np.array_equal(rec[0], B.ROUTECHOICE.dist)
    assert np.array_equal(rec[max(rec)], A.ROUTECHOICE.dist) and not np.array_equal(rec[0], rec[max(rec)])  # congestion changed the costs
    assert np.isfinite(A.ROUTECHOICE.dist[A.get_node("o").id, A.get_node("a").id])
    # (2) Vehicle.state = "abort" before the run: that platoon never departs, everything else runs
    Cw = build()
    Cw.VEHICLES["7"].state = "abort"
    Cw.exec_simulation()
    v = Cw.VEHICLES["7"]
    assert v.state == "abort" and v.arrival_time == -1
    total = sum(l.cum_arrival[-1] for l in (Cw.get_link("oa"), Cw.get_link("ob")))
    assert total == (len(Cw.VEHICLES) - 1) * Cw.DELTAN
    assert sum(1 for x in Cw.VEHICLES.values() if x.state == "end") == len(Cw.VEHICLES) - 1


def test_route_table_history_and_pre_run_abort_emulated(emu_api, monkeypatch):
    import uxsim_b200.world as wmod

    monkeypatch.setattr(wmod.C, "load_product", lambda: emu_api)
    _history_and_abort_checks(wmod)


@pytest.mark.gpu
def test_route_table_history_and_pre_run_abort(cuda_api):
    import uxsim_b200.world as wmod

    _history_and_abort_checks(wmod)
