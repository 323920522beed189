"""Engine features beyond the plain loop, as cases that run on ANY C-ABI backend: links_prefer / links_avoid,
specified routes, toll timeseries, route-choice penalty, adjust_node_capacity, multi-entry signal groups, and the mid-run
setters applied between exec_simulation chunks (test_cpp_mode.py:8485-8643 of the reference).  The CPU suite runs them on the
host emulation of the kernels against the oracle (and the oracle against the reference's own C++ where no random draw is
involved); the GPU suite runs them on the CUDA engine against the oracle."""
import numpy as np

from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S


def _two_route(**kw):
    """orig -> a -> dest over a short (l_top*) and a long (l_bot*) branch plus a detour link, stochastic route choice."""
    s = S.Scenario("two_route", S._world(deltan=5, tmax=3000, **kw))
    for name, x, y in (("orig", 0, 0), ("a", 1, 0), ("top", 2, 1), ("bot", 2, -1), ("dest", 3, 0)):
        s.addNode(name, x, y)
    s.addLink("l0", "orig", "a", length=500, free_flow_speed=20, number_of_lanes=2)
    s.addLink("top1", "a", "top", length=800, free_flow_speed=20)
    s.addLink("top2", "top", "dest", length=800, free_flow_speed=20)
    s.addLink("bot1", "a", "bot", length=1000, free_flow_speed=20)
    s.addLink("bot2", "bot", "dest", length=1000, free_flow_speed=20)
    s.addLink("cross", "top", "bot", length=300, free_flow_speed=10)
    s.adddemand("orig", "dest", 0, 1500, 0.6)
    return s


def links_prefer(api, **kw):
    sc = _two_route(random_seed=11)
    E = sc.to_engine(api, **kw)
    li = sc.link_index()
    P = E.get_int(C.I_NUM_VEHICLES)
    for v in range(0, P, 3):  # every third platoon prefers the long branch
        E.set_vehicle_links(v, 1, [li["bot1"], li["bot2"]])
    E.main_loop()
    return E


def links_avoid(api, **kw):
    sc = _two_route(random_seed=12)
    E = sc.to_engine(api, **kw)
    li = sc.link_index()
    P = E.get_int(C.I_NUM_VEHICLES)
    for v in range(0, P, 2):  # every second platoon must not use the short branch nor the detour
        E.set_vehicle_links(v, 2, [li["top1"], li["cross"]])
    E.main_loop()
    return E


def prefer_and_avoid_with_demand_lists(api, **kw):
    """links_preferred given with the demand itself (add_demand's last argument; uxsim_cpp_wrapper.py:1007-1008)."""
    sc = _two_route(random_seed=13)
    sc.demands.clear()
    E = sc.to_engine(api, **kw)
    ni, li = sc.node_index(), sc.link_index()
    # to_engine froze the (empty) scenario; demand can still be added before the first main_loop
    E.add_demand(ni["orig"], ni["dest"], 0, 800, 0.5, [li["bot1"], li["bot2"]])
    E.add_demand(ni["orig"], ni["dest"], 300, 1200, 0.3)
    E.main_loop()
    return E


def specified_route(api, **kw):
    sc = _two_route(random_seed=14)
    E = sc.to_engine(api, **kw)
    li = sc.link_index()
    P = E.get_int(C.I_NUM_VEHICLES)
    for v in range(0, P, 4):
        E.set_vehicle_links(v, 0, [li["l0"], li["top1"], li["cross"], li["bot2"]])
    E.main_loop()
    return E


def toll_timeseries(api, **kw):
    """Congestion pricing (traffi.cpp:1285-1289): a toll on the short branch between 600 s and 1800 s moves the route search
    to the long branch for that period."""
    sc = _two_route(random_seed=15, duo_update_time=300)
    E = sc.to_engine(api, **kw)
    li = sc.link_index()
    T = E.get_int(C.I_TOTAL_TIMESTEPS)
    toll = np.zeros(T)
    toll[int(600 / sc.delta_t):int(1800 / sc.delta_t)] = 200.0
    E.set_toll_timeseries(li["top1"], toll)
    E.main_loop()
    return E


def route_choice_penalty(api, **kw):
    sc = _two_route(random_seed=16, duo_update_time=300)
    E = sc.to_engine(api, **kw)
    E.set_link_param(sc.link_index()["top2"], C.LINK_ROUTE_CHOICE_PENALTY, 150.0)
    E.main_loop()
    return E


def adjust_node_capacity(api, **kw):
    """adjust_node_capacity=True (traffi.cpp:212-224): nodes without a flow capacity get (max in + max out) / 2."""
    sc = S.grid(n=5, seed=6, tmax=3000, demand_t=1000, adjust_node_capacity=True)
    E = sc.to_engine(api, **kw)
    E.main_loop()
    return E


def adjust_node_capacity_deterministic(api, **kw):
    sc = S.lane_drop_2to1(adjust_node_capacity=True)
    E = sc.to_engine(api, **kw)
    E.main_loop()
    return E


def _signal_cross(**kw):
    """Four-leg signalised intersection, three phases; the east-west legs are green in phases 0 AND 1 (signal_group=[0, 1])."""
    s = S.Scenario("signal_cross", S._world(deltan=5, tmax=2400, **kw))
    s.addNode("c", 0, 0, signal=[40, 20, 30], signal_offset=10)
    for name, x, y in (("w", -1, 0), ("e", 1, 0), ("s", 0, -1), ("n", 0, 1)):
        s.addNode(name, x, y)
    for a, grp in (("w", [0, 1]), ("e", [0, 1]), ("s", [2]), ("n", [1, 2])):
        s.addLink(f"{a}_in", a, "c", length=600, free_flow_speed=15, signal_group=grp)
        s.addLink(f"{a}_out", "c", a, length=600, free_flow_speed=15)
    for o, d, f in (("w", "e", 0.3), ("e", "w", 0.25), ("s", "n", 0.2), ("n", "s", 0.2), ("w", "n", 0.1), ("s", "e", 0.1)):
        s.adddemand(o, d, 0, 1500, f)
    return s


def signal_group_multi(api, **kw):
    E = _signal_cross(random_seed=3).to_engine(api, **kw)
    E.main_loop()
    return E


def signal_group_multi_deterministic(api, **kw):
    E = _signal_cross(hard_deterministic_mode=True).to_engine(api, **kw)
    E.main_loop()
    return E


def midrun_setters(api, **kw):
    """Chunked run with interventions between the chunks: signal timing, free-flow speed, jam density, outflow capacity,
    merge priority, signal group and demand added after the start (test_cpp_mode.py:8527-8643)."""
    sc = _signal_cross(random_seed=5)
    sc.world["tmax"] = 3600
    E = sc.to_engine(api, **kw)
    ni, li = sc.node_index(), sc.link_index()
    E.main_loop(until_t=400)
    E.set_node_signal(ni["c"], [25.0, 25.0, 25.0], -1, -1.0)       # new timing, phase and elapsed time kept
    E.main_loop(until_t=800)
    E.set_link_param(li["w_in"], C.LINK_VMAX, 8.0)                  # change_free_flow_speed
    E.set_link_param(li["e_out"], C.LINK_KAPPA, 0.15)               # change_jam_density
    E.main_loop(until_t=1200)
    E.set_link_param(li["s_in"], C.LINK_CAPACITY_OUT, 0.2)          # capacity_out
    E.set_link_param(li["n_in"], C.LINK_MERGE_PRIORITY, 3.0)
    E.set_node_signal(ni["c"], None, 2, 5.0)                        # jump to phase 2, 5 s elapsed
    E.add_demand(ni["n"], ni["w"], 1300, 1900, 0.25)                # demand added mid-run (a destination with demand before)
    E.main_loop(until_t=2000)
    E.set_link_signal_group(li["s_in"], [0, 2])                     # a group of a different size
    E.set_node_signal(ni["c"], [20.0, 20.0, 20.0, 10.0], -1, -1.0)  # a plan with a different number of phases
    E.main_loop(until_t=2600)
    E.set_node_signal(ni["c"], [30.0, 30.0], -1, -1.0)              # ... and a shorter one (a running phase 2 or 3 restarts at 0)
    E.set_link_param(li["s_in"], C.LINK_CAPACITY_OUT, -1.0)         # restriction lifted
    E.set_link_param(li["w_in"], C.LINK_VMAX, 15.0)
    E.main_loop()
    return E


def midrun_setters_deterministic(api, **kw):
    sc = S.straight_2link_signal()
    E = sc.to_engine(api, **kw)
    E.main_loop(until_t=500)
    E.set_node_signal(1, [30.0, 90.0], -1, -1.0)
    E.set_link_param(0, C.LINK_VMAX, 12.0)
    E.main_loop(until_t=1000)
    E.set_link_param(1, C.LINK_CAPACITY_OUT, 0.3)
    E.set_link_param(0, C.LINK_KAPPA, 0.25)
    E.main_loop()
    return E


#: (name, runner, deterministic) -- deterministic cases involve no random draw, so the reference's own C++ engine must agree too
CASES = [
    ("links_prefer", links_prefer, False), ("links_avoid", links_avoid, False), ("prefer_with_demand", prefer_and_avoid_with_demand_lists, False),
    ("specified_route", specified_route, False), ("toll_timeseries", toll_timeseries, False), ("route_choice_penalty", route_choice_penalty, False),
    ("adjust_node_capacity", adjust_node_capacity, False), ("adjust_node_capacity_det", adjust_node_capacity_deterministic, True),
    ("signal_group_multi", signal_group_multi, False), ("signal_group_multi_det", signal_group_multi_deterministic, True),
    ("midrun_setters", midrun_setters, False), ("midrun_setters_det", midrun_setters_deterministic, True),
]
