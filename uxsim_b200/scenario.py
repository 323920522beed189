"""Backend-neutral scenario description and the scenario builders for the BASELINE.json configs.

A :class:`Scenario` is plain data: World kwargs, node dicts, link dicts and demand records, using the
argument names of the reference API (``World.addNode`` / ``addLink`` / ``adddemand``,
uxsim/uxsim.py:1848-2053).  It can be replayed into

* any library exporting the C-ABI (:meth:`Scenario.to_engine`) -- the CUDA backend in the product, the
  oracle / reference harness in tests;
* the :class:`uxsim_b200.world.CudaWorld` wrapper (:meth:`Scenario.to_world`), which mirrors the reference's
  ``World`` API;
* the reference's own pure-Python ``World`` (done in tests/golden/make_golden.py, never here).

The builders below restate the scenarios named in BASELINE.json / SURVEY.md section 8(d).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


@dataclass
class Scenario:
    name: str = ""
    world: Dict[str, Any] = field(default_factory=dict)      # kwargs of World(...)
    nodes: List[Dict[str, Any]] = field(default_factory=list)  # kwargs of addNode
    links: List[Dict[str, Any]] = field(default_factory=list)  # kwargs of addLink
    demands: List[Tuple] = field(default_factory=list)          # ("demand", o, d, t0, t1, flow, volume) | ("vehicle", o, d, t)
    bulk_vehicles: Optional[Tuple] = None                       # (orig ids, dest ids, departure steps) appended after `demands`

    # ---- construction helpers (reference argument names) ------------------------------------------
    def addNode(self, name, x, y, signal=(0,), signal_offset=0, flow_capacity=None, number_of_lanes=None):
        self.nodes.append(dict(name=str(name), x=float(x), y=float(y), signal=list(signal), signal_offset=signal_offset,
                               flow_capacity=flow_capacity, number_of_lanes=number_of_lanes))
        return str(name)

    def addLink(self, name, start_node, end_node, length, free_flow_speed=20, jam_density=0.2, number_of_lanes=1,
                merge_priority=1, signal_group=(0,), capacity_out=None, capacity_in=None):
        if isinstance(signal_group, int):
            signal_group = [signal_group]
        self.links.append(dict(name=str(name), start_node=str(start_node), end_node=str(end_node), length=float(length),
                               free_flow_speed=float(free_flow_speed), jam_density=float(jam_density),
                               number_of_lanes=int(number_of_lanes), merge_priority=float(merge_priority),
                               signal_group=list(signal_group), capacity_out=capacity_out, capacity_in=capacity_in))
        return str(name)

    def adddemand(self, orig, dest, t_start, t_end, flow=-1, volume=-1):
        self.demands.append(("demand", str(orig), str(dest), float(t_start), float(t_end), float(flow), float(volume)))

    def addVehicle(self, orig, dest, departure_time):
        self.demands.append(("vehicle", str(orig), str(dest), float(departure_time)))

    # ---- derived quantities ------------------------------------------------------------------------------
    @property
    def deltan(self) -> float:
        return self.world.get("deltan", 5)

    @property
    def delta_t(self) -> float:
        return self.world.get("reaction_time", 1) * self.deltan

    def node_index(self) -> Dict[str, int]:
        return {n["name"]: i for i, n in enumerate(self.nodes)}

    def link_index(self) -> Dict[str, int]:
        return {l["name"]: i for i, l in enumerate(self.links)}

    def auto_tmax(self) -> float:
        """uxsim.py:2325-2333 -- (max departure // 1800 + 2) * 1800.  Needs the departure steps, so it replays
        the demand accumulation (cheap)."""
        dt, dn = self.delta_t, self.deltan
        tmax = 0.0
        for d in self.demands:
            if d[0] == "vehicle":
                dep = int(d[3] / dt) * dt
                tmax = max(tmax, dep)
            else:
                _, _, _, t0, t1, flow, volume = d
                if volume > 0:
                    flow = volume / (t1 - t0)
                f = 0.0
                last = None
                for t in range(int(t0 / dt), int(t1 / dt)):
                    f += flow * dt
                    while f >= dn:
                        last = t
                        f -= dn
                if last is not None:
                    tmax = max(tmax, last * dt)
        return (tmax // 1800 + 2) * 1800

    def tmax(self) -> float:
        t = self.world.get("tmax")
        return float(t) if t is not None else float(self.auto_tmax())

    # ---- replay into a C-ABI backend ------------------------------------------------------------------------
    def engine_kwargs(self, **over) -> Dict[str, Any]:
        w = dict(self.world)
        w.update(over)
        logging = w.get("vehicle_logging_timestep_interval", 1)
        return dict(
            t_max=self.tmax() if w.get("tmax") is None else float(w["tmax"]),
            delta_n=float(w.get("deltan", 5)), tau=float(w.get("reaction_time", 1)),
            duo_update_time=float(w.get("duo_update_time", 600)), duo_update_weight=float(w.get("duo_update_weight", 0.5)),
            route_choice_uncertainty=float(w.get("duo_noise", 0.01)), random_seed=int(w.get("random_seed") or 0),
            vehicle_log_mode=bool(logging is not None and logging > 0),
            hard_deterministic_mode=bool(w.get("hard_deterministic_mode", False)),
            route_choice_update_gradual=bool(w.get("route_choice_update_gradual", False)),
            no_cyclic_routing=bool(w.get("no_cyclic_routing", False)),
            adjust_node_capacity=bool(w.get("adjust_node_capacity", False)),
            instantaneous_TT_timestep_interval=int(w.get("instantaneous_TT_timestep_interval", 5)),
            num_replicas=int(w.get("num_replicas", 1)), device=int(w.get("device", -1)),
            num_threads=int(w.get("threads", 1)),
        )

    def to_engine(self, api, **over):
        """Replay the scenario through the C-ABI of ``api`` (a :class:`uxsim_b200._capi.CApi`).  Mirrors what
        CppWorld does call by call (uxsim_cpp_wrapper.py:73-76, 230-232, 1007-1008, 1087-1099)."""
        from ._capi import EngineWorld

        E = EngineWorld(api, **self.engine_kwargs(**over))
        nidx, lidx = self.node_index(), {}
        for n in self.nodes:
            E.add_node(n["x"], n["y"], n["signal"], n["signal_offset"],
                       -1.0 if n["flow_capacity"] is None else n["flow_capacity"],
                       -1 if n["number_of_lanes"] is None else n["number_of_lanes"])
        for l in self.links:
            lidx[l["name"]] = E.add_link(
                nidx[l["start_node"]], nidx[l["end_node"]], l["free_flow_speed"], l["jam_density"], l["length"],
                l["number_of_lanes"], l["merge_priority"], -1.0 if l["capacity_out"] is None else l["capacity_out"],
                -1.0 if l["capacity_in"] is None else l["capacity_in"], l["signal_group"])
        dt, dn = self.delta_t, self.deltan
        vo, vd, vt = [], [], []

        def flush():
            if vo:
                E.add_vehicles(vo, vd, vt)
                vo.clear(), vd.clear(), vt.clear()

        for d in self.demands:
            if d[0] == "vehicle":
                vo.append(nidx[d[1]]), vd.append(nidx[d[2]]), vt.append(int(d[3] / dt))
            else:
                flush()
                _, o, dd, t0, t1, flow, volume = d
                if volume > 0:
                    flow = volume / (t1 - t0)
                E.add_demand(nidx[o], nidx[dd], t0, t1, flow)
        flush()
        if self.bulk_vehicles is not None:
            E.add_vehicles(*self.bulk_vehicles)
        E.initialize_adj_matrix()
        return E

    # ---- flat tables + vectorised ingest (3 FFI calls instead of one per entity) ---------------------------
    def tables(self) -> Dict[str, np.ndarray]:
        """The scenario as flat host arrays in the layout of ux_add_nodes / ux_add_links / ux_add_demands."""
        nidx = self.node_index()
        sig_off = np.zeros(len(self.nodes) + 1, np.int32)
        sig = []
        for i, n in enumerate(self.nodes):
            sig.extend(n["signal"])
            sig_off[i + 1] = len(sig)
        sg_off = np.zeros(len(self.links) + 1, np.int32)
        sg = []
        for i, l in enumerate(self.links):
            sg.extend(l["signal_group"])
            sg_off[i + 1] = len(sg)
        opt = lambda v, d: d if v is None else v
        dem = [d for d in self.demands if d[0] == "demand"]
        veh = [d for d in self.demands if d[0] == "vehicle"]
        flow = np.array([(d[6] / (d[4] - d[3])) if d[6] > 0 else d[5] for d in dem], np.float64)
        dt = self.delta_t
        return dict(
            node_x=np.array([n["x"] for n in self.nodes], np.float64), node_y=np.array([n["y"] for n in self.nodes], np.float64),
            sig_off=sig_off, sig=np.array(sig, np.float64), sig_offset=np.array([n["signal_offset"] for n in self.nodes], np.float64),
            node_fc=np.array([opt(n["flow_capacity"], -1.0) for n in self.nodes], np.float64),
            node_lanes=np.array([opt(n["number_of_lanes"], -1) for n in self.nodes], np.int32),
            l_start=np.array([nidx[l["start_node"]] for l in self.links], np.int32),
            l_end=np.array([nidx[l["end_node"]] for l in self.links], np.int32),
            l_vmax=np.array([l["free_flow_speed"] for l in self.links], np.float64),
            l_kappa=np.array([l["jam_density"] for l in self.links], np.float64),
            l_length=np.array([l["length"] for l in self.links], np.float64),
            l_lanes=np.array([l["number_of_lanes"] for l in self.links], np.int32),
            l_mp=np.array([l["merge_priority"] for l in self.links], np.float64),
            l_cap_out=np.array([opt(l["capacity_out"], -1.0) for l in self.links], np.float64),
            l_cap_in=np.array([opt(l["capacity_in"], -1.0) for l in self.links], np.float64),
            sg_off=sg_off, sg=np.array(sg, np.int32),
            d_orig=np.array([nidx[d[1]] for d in dem], np.int32), d_dest=np.array([nidx[d[2]] for d in dem], np.int32),
            d_t0=np.array([d[3] for d in dem], np.float64), d_t1=np.array([d[4] for d in dem], np.float64), d_flow=flow,
            v_orig=np.concatenate([np.array([nidx[d[1]] for d in veh], np.int32)] + ([self.bulk_vehicles[0]] if self.bulk_vehicles else [])),
            v_dest=np.concatenate([np.array([nidx[d[2]] for d in veh], np.int32)] + ([self.bulk_vehicles[1]] if self.bulk_vehicles else [])),
            v_ts=np.concatenate([np.array([int(d[3] / dt) for d in veh], np.int32)] + ([self.bulk_vehicles[2]] if self.bulk_vehicles else [])),
        )

    def engine_from_tables(self, api, tab: Dict[str, np.ndarray], **over):
        """Create + fill + freeze a native world from flat host arrays (the e2e path of bench.py)."""
        from ._capi import EngineWorld

        E = EngineWorld(api, **self.engine_kwargs(**over))
        E.add_nodes(tab["node_x"], tab["node_y"], tab["sig_off"], tab["sig"], tab["sig_offset"], tab["node_fc"], tab["node_lanes"])
        E.add_links(tab["l_start"], tab["l_end"], tab["l_vmax"], tab["l_kappa"], tab["l_length"], tab["l_lanes"], tab["l_mp"],
                    tab["l_cap_out"], tab["l_cap_in"], tab["sg_off"], tab["sg"])
        if len(tab["d_orig"]):
            E.add_demands(tab["d_orig"], tab["d_dest"], tab["d_t0"], tab["d_t1"], tab["d_flow"])
        if len(tab["v_orig"]):
            E.add_vehicles(tab["v_orig"], tab["v_dest"], tab["v_ts"])
        E.initialize_adj_matrix()
        return E

    def to_world(self, **over):
        """Replay through the reference-style API of the CUDA backend (uxsim_b200.World(cuda=True))."""
        from .world import World

        kw = dict(self.world)
        kw.update(over)
        kw.setdefault("print_mode", 0), kw.setdefault("save_mode", 0), kw.setdefault("show_mode", 0)
        W = World(cuda=True, **kw)
        replay_into(self, W)
        return W


def replay_into(sc: Scenario, W):
    """Feed a Scenario into any object with the reference World API (addNode/addLink/adddemand/addVehicle)."""
    for n in sc.nodes:
        W.addNode(n["name"], n["x"], n["y"], signal=n["signal"], signal_offset=n["signal_offset"],
                  flow_capacity=n["flow_capacity"], number_of_lanes=n["number_of_lanes"])
    for l in sc.links:
        W.addLink(l["name"], l["start_node"], l["end_node"], length=l["length"], free_flow_speed=l["free_flow_speed"],
                  jam_density=l["jam_density"], number_of_lanes=l["number_of_lanes"], merge_priority=l["merge_priority"],
                  signal_group=l["signal_group"], capacity_out=l["capacity_out"], capacity_in=l["capacity_in"])
    for d in sc.demands:
        if d[0] == "vehicle":
            W.addVehicle(d[1], d[2], d[3])
        else:
            _, o, dd, t0, t1, flow, volume = d
            W.adddemand(o, dd, t0, t1, flow=flow, volume=volume)
    return W


# =====================================================================================================
# small deterministic known-answer scenarios (tests/test_verification_straight_road.py and _node.py of
# the reference define the same networks; cited per builder)
# =====================================================================================================

def _world(**kw):
    base = dict(print_mode=0, save_mode=0, show_mode=0, random_seed=0)
    base.update(kw)
    return base


def straight_1link(deltan=5, tmax=2000, flow=0.5, t_end=1000, **kw) -> Scenario:
    """test_verification_straight_road.py:19-50 (test_1link)."""
    s = Scenario("1link", _world(deltan=deltan, tmax=tmax, **kw))
    s.addNode("orig", 0, 0), s.addNode("dest", 1, 1)
    s.addLink("link", "orig", "dest", length=1000, free_flow_speed=20, jam_density=0.2)
    s.adddemand("orig", "dest", 0, t_end, flow)
    return s


def straight_2link_bottleneck_capacity_out(deltan=5, tmax=2000, **kw) -> Scenario:
    """test_verification_straight_road.py (test_2link_bottleneck_due_to_capacity_out)."""
    s = Scenario("2link_capout", _world(deltan=deltan, tmax=tmax, **kw))
    s.addNode("orig", 0, 0), s.addNode("mid", 1, 1), s.addNode("dest", 2, 2)
    s.addLink("link1", "orig", "mid", length=1000, free_flow_speed=20, jam_density=0.2, capacity_out=0.4)
    s.addLink("link2", "mid", "dest", length=1000, free_flow_speed=20, jam_density=0.2)
    s.adddemand("orig", "dest", 0, 500, 0.8)
    s.adddemand("orig", "dest", 500, 1500, 0.4)
    return s


def straight_2link_bottleneck_speed(deltan=5, tmax=2000, **kw) -> Scenario:
    """Free-flow speed drop 20 -> 10 m/s (test_2link_bottleneck_due_to_free_flow_speed)."""
    s = Scenario("2link_speed", _world(deltan=deltan, tmax=tmax, **kw))
    s.addNode("orig", 0, 0), s.addNode("mid", 1, 1), s.addNode("dest", 2, 2)
    s.addLink("link1", "orig", "mid", length=1000, free_flow_speed=20, jam_density=0.2)
    s.addLink("link2", "mid", "dest", length=1000, free_flow_speed=10, jam_density=0.2)
    s.adddemand("orig", "dest", 0, 500, 0.8)
    s.adddemand("orig", "dest", 500, 1500, 0.4)
    return s


def straight_3link_spillback(deltan=5, tmax=2000, **kw) -> Scenario:
    """test_verification_straight_road.py:690-748 (test_3link_queuespillback)."""
    s = Scenario("3link_spillback", _world(deltan=deltan, tmax=tmax, **kw))
    s.addNode("orig", 0, 0), s.addNode("mid1", 1, 1), s.addNode("mid2", 2, 2), s.addNode("dest", 3, 3)
    s.addLink("link1", "orig", "mid1", length=1000, free_flow_speed=20, jam_density=0.2)
    s.addLink("link2", "mid1", "mid2", length=1000, free_flow_speed=20, jam_density=0.2, capacity_out=0.4)
    s.addLink("link3", "mid2", "dest", length=1000, free_flow_speed=20, jam_density=0.2)
    s.adddemand("orig", "dest", 0, 400, 0.8)
    s.adddemand("orig", "dest", 400, 800, 0.4)
    s.adddemand("orig", "dest", 800, 1200, 0.2)
    return s


def straight_2link_signal(deltan=5, tmax=2000, **kw) -> Scenario:
    """test_verification_straight_road.py:750-839 (test_2link_signal)."""
    s = Scenario("2link_signal", _world(deltan=deltan, tmax=tmax, **kw))
    s.addNode("orig", 0, 0), s.addNode("mid", 1, 1, signal=[60, 60]), s.addNode("dest", 2, 2)
    s.addLink("link1", "orig", "mid", length=1000, free_flow_speed=20, jam_density=0.2, signal_group=0)
    s.addLink("link2", "mid", "dest", length=1000, free_flow_speed=20, jam_density=0.2)
    s.adddemand("orig", "dest", 0, 1000, 0.4)
    return s


def straight_node_capacity(deltan=1, tmax=2000, **kw) -> Scenario:
    """Node flow_capacity bottleneck (test_verification_straight_road.py, node-capacity cases)."""
    s = Scenario("node_capacity", _world(deltan=deltan, tmax=tmax, **kw))
    s.addNode("orig", 0, 0), s.addNode("mid", 1, 1, flow_capacity=0.4), s.addNode("dest", 2, 2)
    s.addLink("link1", "orig", "mid", length=1000, free_flow_speed=20, jam_density=0.2)
    s.addLink("link2", "mid", "dest", length=1000, free_flow_speed=20, jam_density=0.2)
    s.adddemand("orig", "dest", 0, 500, 0.8)
    s.adddemand("orig", "dest", 500, 1500, 0.3)
    return s


def lane_drop_2to1(deltan=5, tmax=2000, **kw) -> Scenario:
    """2-lane link feeding a 1-lane link (test_verification_multilane.py)."""
    s = Scenario("lane_drop", _world(deltan=deltan, tmax=tmax, **kw))
    s.addNode("orig", 0, 0), s.addNode("mid", 1, 1), s.addNode("dest", 2, 2)
    s.addLink("link1", "orig", "mid", length=1000, free_flow_speed=20, jam_density=0.4, number_of_lanes=2)
    s.addLink("link2", "mid", "dest", length=1000, free_flow_speed=20, jam_density=0.2, number_of_lanes=1)
    s.adddemand("orig", "dest", 0, 500, 1.2)
    s.adddemand("orig", "dest", 500, 1500, 0.4)
    return s


def multilane_3lane(deltan=5, tmax=2000, **kw) -> Scenario:
    """3-lane straight road (multi-lane queue / lane assignment / trip-end flush)."""
    s = Scenario("3lane", _world(deltan=deltan, tmax=tmax, **kw))
    s.addNode("orig", 0, 0), s.addNode("mid", 1, 1), s.addNode("dest", 2, 2)
    s.addLink("link1", "orig", "mid", length=1000, free_flow_speed=20, jam_density=0.6, number_of_lanes=3)
    s.addLink("link2", "mid", "dest", length=600, free_flow_speed=15, jam_density=0.6, number_of_lanes=3, capacity_out=1.0)
    s.adddemand("orig", "dest", 0, 800, 1.8)
    return s


def merge_y(deltan=5, tmax=2000, priority1=1.0, priority2=1.0, hard_deterministic_mode=True, **kw) -> Scenario:
    """README Y-merge / test_verification_node.py merge cases.  Deterministic when hard_deterministic_mode."""
    s = Scenario("merge", _world(deltan=deltan, tmax=tmax, hard_deterministic_mode=hard_deterministic_mode, **kw))
    s.addNode("orig1", 0, 0), s.addNode("orig2", 0, 2), s.addNode("merge", 1, 1), s.addNode("dest", 2, 1)
    s.addLink("link1", "orig1", "merge", length=1000, free_flow_speed=20, merge_priority=priority1)
    s.addLink("link2", "orig2", "merge", length=1000, free_flow_speed=20, merge_priority=priority2)
    s.addLink("link3", "merge", "dest", length=1000, free_flow_speed=20)
    s.adddemand("orig1", "dest", 0, 1000, 0.45)
    s.adddemand("orig2", "dest", 400, 1000, 0.6)
    return s


def diverge_2route(deltan=5, tmax=3000, hard_deterministic_mode=True, **kw) -> Scenario:
    """Two routes of clearly different free-flow time between one OD pair (tie-free route choice;
    test_verification_route_choice.py style)."""
    s = Scenario("diverge", _world(deltan=deltan, tmax=tmax, hard_deterministic_mode=hard_deterministic_mode,
                                   duo_update_time=100, **kw))
    s.addNode("orig", 0, 0), s.addNode("mid1", 1, 1), s.addNode("mid2", 1, -1), s.addNode("dest", 2, 0)
    s.addLink("link1a", "orig", "mid1", length=1000, free_flow_speed=20)
    s.addLink("link1b", "mid1", "dest", length=1000, free_flow_speed=20, capacity_out=0.3)
    s.addLink("link2a", "orig", "mid2", length=1500, free_flow_speed=20)
    s.addLink("link2b", "mid2", "dest", length=1500, free_flow_speed=20)
    s.adddemand("orig", "dest", 0, 1500, 0.6)
    return s


def dead_end(deltan=5, tmax=1500, **kw) -> Scenario:
    """A vehicle routed into a dead end aborts its trip (test_verification_exceptional.py)."""
    s = Scenario("dead_end", _world(deltan=deltan, tmax=tmax, **kw))
    s.addNode("orig", 0, 0), s.addNode("mid", 1, 0), s.addNode("dest", 2, 0), s.addNode("deadend", 1, 1)
    s.addLink("l1", "orig", "mid", length=500, free_flow_speed=20)
    s.addLink("l2", "mid", "dest", length=500, free_flow_speed=20)
    s.addLink("l3", "mid", "deadend", length=500, free_flow_speed=20)
    s.adddemand("orig", "dest", 0, 500, 0.4)
    s.adddemand("orig", "deadend", 0, 500, 0.2)
    return s


def short_links_chain(deltan=5, tmax=1500, reverse_ids=True, **kw) -> Scenario:
    """Links shorter than one platoon spacing / one free-flow step, with node ids DEcreasing downstream, so that
    the serial id-ordered Node.transfer visibility decides the outcome."""
    s = Scenario("short_chain", _world(deltan=deltan, tmax=tmax, **kw))
    names = ["n0", "n1", "n2", "n3", "n4"]
    order = list(reversed(names)) if reverse_ids else names
    for k, n in enumerate(order):
        s.addNode(n, names.index(n), 0)
    s.addLink("l0", "n0", "n1", length=400, free_flow_speed=20)
    s.addLink("l1", "n1", "n2", length=20, free_flow_speed=20)   # shorter than delta*deltan = 25 m
    s.addLink("l2", "n2", "n3", length=120, free_flow_speed=20)  # shorter than 2*u*dt + delta*deltan = 225 m
    s.addLink("l3", "n3", "n4", length=400, free_flow_speed=10)
    s.adddemand("n0", "n4", 0, 800, 0.5)
    return s


# =====================================================================================================
# BASELINE.json configs
# =====================================================================================================

def grid(n=11, deltan=5, tmax=7200, demand_flow=0.035, demand_t=3600, length=1000, seed=0, signals=False, **kw) -> Scenario:
    """C2: demos_and_examples/example_04en_automatic_network_generation.py:4-47 -- n x n grid, 4-neighbour links,
    boundary-to-boundary demand.  ``signals=True`` adds the C4-style two-phase signal at every node."""
    s = Scenario(f"grid{n}", _world(deltan=deltan, tmax=tmax, random_seed=seed, no_cyclic_routing=True, **kw))
    for i in range(n):
        for j in range(n):
            s.addNode(f"n{(i, j)}", i, j, signal=[60, 60] if signals else (0,))
    for i in range(n):
        for j in range(n):
            if i != n - 1:
                s.addLink(f"l{(i, j, i + 1, j)}", f"n{(i, j)}", f"n{(i + 1, j)}", length=length, signal_group=0)
            if i != 0:
                s.addLink(f"l{(i, j, i - 1, j)}", f"n{(i, j)}", f"n{(i - 1, j)}", length=length, signal_group=0)
            if j != n - 1:
                s.addLink(f"l{(i, j, i, j + 1)}", f"n{(i, j)}", f"n{(i, j + 1)}", length=length, signal_group=1)
            if j != 0:
                s.addLink(f"l{(i, j, i, j - 1)}", f"n{(i, j)}", f"n{(i, j - 1)}", length=length, signal_group=1)
    for n1 in [(0, j) for j in range(n)]:
        for n2 in [(n - 1, j) for j in range(n)]:
            s.adddemand(f"n{n2}", f"n{n1}", 0, demand_t, demand_flow)
            s.adddemand(f"n{n1}", f"n{n2}", 0, demand_t, demand_flow)
    for n1 in [(i, 0) for i in range(n)]:
        for n2 in [(i, n - 1) for i in range(n)]:
            s.adddemand(f"n{n2}", f"n{n1}", 0, demand_t, demand_flow)
            s.adddemand(f"n{n1}", f"n{n2}", 0, demand_t, demand_flow)
    return s


def grid_random_od(n=100, n_platoons=1_000_000, deltan=5, tmax=7200, demand_t=3600, length=1000, seed=0, **kw) -> Scenario:
    """C4: n x n grid, signal [60, 60] at every node (east-west group 0, north-south group 1), random
    OD pairs with uniform departure times.  Node id = i*n + j."""
    s = Scenario(f"grid{n}_randomOD", _world(deltan=deltan, tmax=tmax, random_seed=seed, **kw))
    for i in range(n):
        for j in range(n):
            s.addNode(f"{i * n + j}", i, j, signal=[60, 60])
    for i in range(n):
        for j in range(n):
            a = i * n + j
            if i != n - 1:
                s.addLink(f"{a}_{a + n}", f"{a}", f"{a + n}", length=length, signal_group=0)
            if i != 0:
                s.addLink(f"{a}_{a - n}", f"{a}", f"{a - n}", length=length, signal_group=0)
            if j != n - 1:
                s.addLink(f"{a}_{a + 1}", f"{a}", f"{a + 1}", length=length, signal_group=1)
            if j != 0:
                s.addLink(f"{a}_{a - 1}", f"{a}", f"{a - 1}", length=length, signal_group=1)
    rng = np.random.default_rng(seed)
    o = rng.integers(0, n * n, n_platoons)
    d = rng.integers(0, n * n, n_platoons)
    same = o == d
    while same.any():
        d[same] = rng.integers(0, n * n, int(same.sum()))
        same = o == d
    t = rng.uniform(0, demand_t, n_platoons)
    s.bulk_vehicles = (o.astype(np.int32), d.astype(np.int32), (t / s.delta_t).astype(np.int32))  # addVehicle: step = int(t/dt), uxsim.py:1045-1048
    return s


def _load_npz(name: str):
    p = os.path.join(_DATA, name)
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} missing; regenerate with tests/golden/make_scenario_fixtures.py")
    return np.load(p, allow_pickle=False)


def from_tables(name: str, world: Dict[str, Any], z) -> Scenario:
    s = Scenario(name, world)
    z = {k: z[k] for k in z.files}  # NpzFile decompresses on every access: load each table once
    for nm, x, y in zip(z["node_name"], z["node_x"], z["node_y"]):
        s.addNode(str(nm), float(x), float(y))
    has = lambda k: k in z
    L = len(z["link_name"])
    for i in range(L):
        s.addLink(str(z["link_name"][i]), str(z["link_start"][i]), str(z["link_end"][i]), length=float(z["link_length"][i]),
                  free_flow_speed=float(z["link_u"][i]), jam_density=float(z["link_kappa"][i]),
                  number_of_lanes=int(z["link_lanes"][i]) if has("link_lanes") else 1,
                  merge_priority=float(z["link_merge_priority"][i]) if has("link_merge_priority") else 1.0,
                  capacity_out=(None if not has("link_capacity_out") or z["link_capacity_out"][i] < 0 else float(z["link_capacity_out"][i])),
                  capacity_in=(None if not has("link_capacity_in") or z["link_capacity_in"][i] < 0 else float(z["link_capacity_in"][i])))
    for i in range(len(z["dem_orig"])):
        s.adddemand(str(z["dem_orig"][i]), str(z["dem_dest"][i]), float(z["dem_t0"][i]), float(z["dem_t1"][i]),
                    flow=float(z["dem_flow"][i]), volume=float(z["dem_volume"][i]))
    return s


def sioux_falls(deltan=5, tmax=7200, seed=0, **kw) -> Scenario:
    """C1: tests/test_verification_sioux_falls.py:26-34 (dat/siouxfalls_{nodes,links,demand}.csv)."""
    return from_tables("siouxfalls", _world(deltan=deltan, tmax=tmax, random_seed=seed, **kw), _load_npz("siouxfalls.npz"))


def chicago_sketch(deltan=30, tmax=10000, seed=42, **kw) -> Scenario:
    """C3: demos_and_examples/example_23en_huge_scale_simulation_at_Chicago.py:3-11
    (dat/chicago_sketch.uxsim_scenario replayed through addNode/addLink/adddemand).  ``deltan=5`` is C3'.

    Data: Chicago-Sketch network, Transportation Networks for Research Core Team,
    https://github.com/bstabler/TransportationNetworks (academic research use; cite the source)."""
    return from_tables("chicago_sketch", _world(deltan=deltan, tmax=tmax, random_seed=seed, **kw), _load_npz("chicago_sketch.npz"))


def dta_grid9(deltan=10, tmax=4800, seed=0, **kw) -> Scenario:
    """C5(a): the 9x9 DUE/DSO grid of demos_and_examples/example_28en_benchmark_cpp_mode.py:113-168 (node
    flow_capacity 1.6, centre cross at 20 m/s, else 10 m/s; demand 0.08 veh/s for 2400 s between the
    middle-third boundary nodes)."""
    imax = jmax = 9
    c = 4
    s = Scenario("dta_grid9", _world(deltan=deltan, tmax=tmax, random_seed=seed, duo_update_time=300, **kw))
    for i in range(imax):
        for j in range(jmax):
            s.addNode(f"n{(i, j)}", i, j, flow_capacity=1.6)
    for i in range(imax):
        for j in range(jmax):
            if i != imax - 1:
                s.addLink(f"l{(i, j, i + 1, j)}", f"n{(i, j)}", f"n{(i + 1, j)}", length=1000, free_flow_speed=20 if j == c else 10)
            if i != 0:
                s.addLink(f"l{(i, j, i - 1, j)}", f"n{(i, j)}", f"n{(i - 1, j)}", length=1000, free_flow_speed=20 if j == c else 10)
            if j != jmax - 1:
                s.addLink(f"l{(i, j, i, j + 1)}", f"n{(i, j)}", f"n{(i, j + 1)}", length=1000, free_flow_speed=20 if i == c else 10)
            if j != 0:
                s.addLink(f"l{(i, j, i, j - 1)}", f"n{(i, j)}", f"n{(i, j - 1)}", length=1000, free_flow_speed=20 if i == c else 10)
    q, dur, outer = 0.08, 2400, 3
    for n1 in [(0, j) for j in range(outer, jmax - outer)]:
        for n2 in [(imax - 1, j) for j in range(outer, jmax - outer)]:
            s.adddemand(f"n{n1}", f"n{n2}", 0, dur, q)
        for n2 in [(i, jmax - 1) for i in range(outer, imax - outer)]:
            s.adddemand(f"n{n1}", f"n{n2}", 0, dur, q)
    for n1 in [(i, 0) for i in range(outer, imax - outer)]:
        for n2 in [(i, jmax - 1) for i in range(outer, imax - outer)]:
            s.adddemand(f"n{n1}", f"n{n2}", 0, dur, q)
        for n2 in [(imax - 1, j) for j in range(outer, jmax - outer)]:
            s.adddemand(f"n{n1}", f"n{n2}", 0, dur, q)
    return s
