"""``World(..., cuda=True)``: the reference's scenario API in front of the CUDA engine.

Mirrors ``uxsim/uxsim_cpp_wrapper.py`` (``CppWorld`` 800-1533, ``CppNode`` 40-181, ``CppLink`` 184-482,
``CppVehicle`` 486-725): same constructor kwargs, same ``addNode / addLink / addVehicle / adddemand* /
finalize_scenario / exec_simulation / check_simulation_ongoing / get_node / get_link / defRoute`` methods, same
attributes for ``analyzer.py`` (``W.VEHICLES / LINKS / NODES``, per-link ``cum_arrival / cum_departure /
traveltime_actual / traveltime_instant``, per-vehicle ``travel_time / arrival_time / distance_traveled / state``),
same unsupported set (taxi mode, ``save / copy / load_scenario`` raise ``NotImplementedError`` exactly like cpp
mode, uxsim_cpp_wrapper.py:929-931, 1468-1472, 1529-1533).  ``user_function`` callbacks are stored, never called
(as in cpp mode).

Selection works like ``cpp=True`` (uxsim.py:1671-1675): ``uxsim_b200.World(..., cuda=True)`` returns a
:class:`CudaWorld`.  Simulation logic lives in ``csrc/engine.cu`` only; this file forwards parameters, keeps the
name<->id maps, and hands results back.  Results are fetched lazily and cached as numpy arrays; with
``as_tensor=True`` helpers the same buffers are exposed as zero-copy torch CUDA tensors
(:func:`uxsim_b200.ensemble.device_tensor`).

If the reference package ``uxsim`` is importable, its unmodified ``Analyzer`` is attached as ``W.analyzer``;
otherwise :class:`uxsim_b200.analysis.BasicAnalyzer` provides the ``basic_analysis`` / ``print_simple_stats``
numbers from the same attributes.
"""
from __future__ import annotations

import csv
import os
import math
import string
import time
import warnings
import bisect
from collections import OrderedDict, deque
from collections import defaultdict as ddict
from collections.abc import Mapping, Sequence

import numpy as np

from . import _capi as C

_STATE_NAMES = ["home", "wait", "run", "end", "abort"]


class CudaNode:
    """uxsim_cpp_wrapper.py:40-181 (CppNode)."""

    def __init__(self, W, name, x, y, signal=[0], signal_offset=0, signal_offset_old=None, flow_capacity=None,
                 number_of_lanes=None, auto_rename=False, attribute=None, user_attribute=None, user_function=None):
        self.W = W
        self.attribute, self.user_attribute, self.user_function = attribute, user_attribute, user_function
        if signal_offset_old is not None:
            signal_offset = sum(signal) - signal_offset_old
        self.inlinks, self.outlinks = dict(), dict()
        self.id = len(W.NODES)
        self.name = name
        if self.name in W.NODES_NAME_DICT:
            if auto_rename:
                self.name = self.name + "_renamed" + "".join(W.rng.choice(list(string.ascii_letters + string.digits), size=8))
            else:
                raise ValueError(f"Node name {self.name} already used by another node.")
        W.NODES.append(self)
        W.NODES_NAME_DICT[self.name] = self
        self.x, self.y = float(x), float(y)
        self._signal = [float(s) for s in signal]
        self.signal_offset = float(signal_offset)
        self.flow_capacity = flow_capacity
        self.number_of_lanes = number_of_lanes
        if flow_capacity is not None and number_of_lanes is None:
            self.number_of_lanes = math.ceil(flow_capacity / 0.8)
        self.flag_lanes_automatically_determined = flow_capacity is not None and number_of_lanes is None
        eid = W._E.add_node(self.x, self.y, self._signal, self.signal_offset,
                            -1.0 if flow_capacity is None else float(flow_capacity),
                            -1 if number_of_lanes is None else int(number_of_lanes))
        assert eid == self.id

    def __repr__(self):
        return f"<Node {self.name}>"

    @property
    def signal(self):
        return list(self._signal)

    @signal.setter
    def signal(self, value):
        self._signal = [float(v) for v in value]
        self.W._E.set_node_signal(self.id, self._signal)

    @property
    def cycle_length(self):
        return sum(self._signal)

    @property
    def signal_phase(self):
        return int(self.W._E.get_array(C.A_NODE_SIGNAL_PHASE)[self.id])

    @signal_phase.setter
    def signal_phase(self, value):
        self.W._E.set_node_signal(self.id, None, phase=int(value))

    @property
    def signal_t(self):
        return float(self.W._E.get_array(C.A_NODE_SIGNAL_T)[self.id])

    @signal_t.setter
    def signal_t(self, value):
        self.W._E.set_node_signal(self.id, None, signal_t=float(value))

    @property
    def signal_log(self):
        return list(self.W._series("signal_log")[self.id][: self.W.T])

    def override_signal(self, signal, groups=None, signal_offset=0, reset=False, signal_offset_old=None):
        """uxsim_cpp_wrapper.py:126-170 (uxsim.py:143-190): a new signal plan for this node and, with `groups`, the phases in
        which each in-link has green.  `groups[i]` lists the links (names or objects) that are green in phase i."""
        self.signal = signal
        self.signal_offset = signal_offset
        if reset:
            self.W._E.set_node_signal(self.id, None, phase=0, signal_t=0.0)
        if signal_offset_old is not None:
            self.signal_offset = self.cycle_length - signal_offset_old
        self.signal_offset = float(self.signal_offset)  # (like the reference's C++ mode, the offset only documents the plan once the node exists)
        if groups is None:
            return None
        if len(groups) != len(signal):
            raise ValueError("The length of `group` must match the length of `signal`.")
        phase_info = ddict(list)
        for i, group in enumerate(groups):
            for link in group:
                l = self.W.get_link(link)
                if l not in self.inlinks.values():
                    raise ValueError(f"Link {l.name} is not an incoming link of node {self.name}.")
                phase_info[l].append(i)
        for l, phases in phase_info.items():
            l.signal_group = phases
            self.W._E.set_link_signal_group(l.id, phases)

    def adjust_node_capacity(self):
        """uxsim_cpp_wrapper.py:172-181: the engine derives the capacity of nodes without one from the connected links when the
        network is frozen (ux_initialize_adj_matrix, traffi.cpp:212-224); this reads the result back."""
        fc = float(self.W._E.get_array(C.A_NODE_FLOW_CAPACITY)[self.id])
        nl = int(self.W._E.get_array(C.A_NODE_NUM_LANES)[self.id])
        if (nl if nl > 0 else None) != self.number_of_lanes:
            self.flag_lanes_automatically_determined = True
        self.flow_capacity = fc if fc >= 0.0 else None
        self.number_of_lanes = nl if nl > 0 else None


class _NativeLinkView:
    """What code written against CppLink reaches through `link._cpp_link` (the nanobind Link, bindings.cpp:819-835): owned numpy copies of
    the link's series and its current counters."""

    def __init__(self, link):
        self._l = link

    def get_cum_arrival_np(self):
        return np.array(self._l.cum_arrival)

    def get_cum_departure_np(self):
        return np.array(self._l.cum_departure)

    def get_traveltime_actual_np(self):
        return np.array(self._l.traveltime_actual)

    def get_traveltime_instant_np(self):
        return np.array(self._l.traveltime_instant)

    @property
    def vehicle_count(self):
        return self._l.num_vehicles // self._l.W.DELTAN

    @property
    def id(self):
        return self._l.id


class CudaLink:
    """uxsim_cpp_wrapper.py:184-482 (CppLink)."""

    @property
    def _cpp_link(self):
        return _NativeLinkView(self)

    def __init__(self, W, name, start_node, end_node, length, free_flow_speed=20, jam_density=0.2, jam_density_per_lane=None,
                 number_of_lanes=1, merge_priority=1, signal_group=[0], capacity_out=None, capacity_in=None,
                 congestion_pricing=None, eular_dx=None, attribute=None, user_attribute=None, user_function=None, auto_rename=False):
        self.W = W
        self.start_node, self.end_node = W.get_node(start_node), W.get_node(end_node)
        self.number_of_lanes = int(number_of_lanes)
        if self.number_of_lanes != number_of_lanes:
            raise ValueError(f"number_of_lanes must be an integer. Got {number_of_lanes}.")
        self.signal_group = [signal_group] if isinstance(signal_group, int) else list(signal_group)
        self.id = len(W.LINKS)
        self.name = name if name is not None else self.start_node.name + "-" + self.end_node.name
        if self.name in W.LINKS_NAME_DICT:
            if auto_rename:
                self.name = self.name + "_renamed" + "".join(W.rng.choice(list(string.ascii_letters + string.digits), size=8))
            else:
                raise ValueError(f"Link name {self.name} already used by another link.")
        W.LINKS.append(self)
        W.LINKS_NAME_DICT[self.name] = self
        self.start_node.outlinks[self.name] = self
        self.end_node.inlinks[self.name] = self
        self.attribute, self.user_attribute, self.user_function = attribute, user_attribute, user_function
        kappa = jam_density if jam_density_per_lane is None else jam_density_per_lane * number_of_lanes
        # derived FD parameters, uxsim.py:519-532
        self.length = length
        self.u = self.free_flow_speed = float(free_flow_speed)
        self.kappa = self.jam_density = float(kappa)
        self.jam_density_per_lane = jam_density_per_lane
        self.tau = W.REACTION_TIME / self.number_of_lanes
        self.w = 1 / self.tau / self.kappa
        self.capacity = self.u * self.w * self.kappa / (self.u + self.w)
        self.delta = 1 / self.kappa
        self.delta_per_lane = self.delta * self.number_of_lanes
        self.q_star, self.k_star = self.capacity, self.capacity / self.u
        self.merge_priority = float(merge_priority)
        self._capacity_out = self.capacity * 2 if capacity_out is None else float(capacity_out)
        self._capacity_in = self.capacity * 2 if capacity_in is None else float(capacity_in)
        self.congestion_pricing = congestion_pricing
        self._route_choice_penalty = 0
        self.vehicles, self.vehicles_enter_log = deque(), {}
        self.tss, self.xss, self.cs, self.ls, self.names = [], [], [], [], []
        self.eular_dx = eular_dx
        self.edie_dx = max(self.length / 10, self.u * W.DELTAT) if eular_dx is None else eular_dx
        eid = W._E.add_link(self.start_node.id, self.end_node.id, self.u, self.kappa, float(length), self.number_of_lanes,
                            self.merge_priority, -1.0 if capacity_out is None else float(capacity_out),
                            -1.0 if capacity_in is None else float(capacity_in), self.signal_group)
        assert eid == self.id

    def __repr__(self):
        return f"<Link {self.name}>"

    def init_after_tmax_fix(self):
        self.edie_dt = self.W.EULAR_DT
        shape = [int(self.W.TMAX / self.edie_dt) + 1, int(self.length / self.edie_dx)]
        self.k_mat, self.q_mat, self.v_mat = np.zeros(shape), np.zeros(shape), np.zeros(shape)
        self.tn_mat, self.dn_mat = np.zeros(shape), np.zeros(shape)
        self.an = self.edie_dt * self.edie_dx

    # time series handed back from the device (uxsim_cpp_wrapper.py:292-333)
    @property
    def cum_arrival(self):
        return self.W._series("cum_arrival")[self.id]

    @property
    def cum_departure(self):
        return self.W._series("cum_departure")[self.id]

    @property
    def traveltime_instant(self):
        return self.W._series("traveltime_instant")[self.id]

    @property
    def traveltime_actual(self):
        return self.W._series("traveltime_actual")[self.id]

    @traveltime_actual.setter
    def traveltime_actual(self, value):
        pass

    @property
    def capacity_out(self):
        return self._capacity_out

    @capacity_out.setter
    def capacity_out(self, v):
        self._capacity_out = float(v)
        self.W._E.set_link_param(self.id, C.LINK_CAPACITY_OUT, float(v))

    @property
    def capacity_in(self):
        return self._capacity_in

    @capacity_in.setter
    def capacity_in(self, v):
        self._capacity_in = float(v)
        self.W._E.set_link_param(self.id, C.LINK_CAPACITY_IN, float(v))

    @property
    def capacity_in_remain(self):
        return float(self.W._E.get_array(C.A_LINK_CAPACITY_IN_REMAIN)[self.id])

    @capacity_in_remain.setter
    def capacity_in_remain(self, v):
        self.W._E.set_link_param(self.id, C.LINK_CAPACITY_IN_REMAIN, float(v))

    @property
    def capacity_out_remain(self):
        return float(self.W._E.get_array(C.A_LINK_CAPACITY_OUT_REMAIN)[self.id])

    @capacity_out_remain.setter
    def capacity_out_remain(self, v):
        self.W._E.set_link_param(self.id, C.LINK_CAPACITY_OUT_REMAIN, float(v))

    @property
    def route_choice_penalty(self):
        return self._route_choice_penalty

    @route_choice_penalty.setter
    def route_choice_penalty(self, v):
        self._route_choice_penalty = v
        self.W._E.set_link_param(self.id, C.LINK_ROUTE_CHOICE_PENALTY, float(v))

    def get_toll(self, t):
        return self.congestion_pricing(t) if self.congestion_pricing else self.route_choice_penalty

    # current state
    @property
    def num_vehicles(self):
        return int(self.W._E.get_array(C.A_LINK_NUM_VEHICLES)[self.id]) * self.W.DELTAN

    @property
    def density(self):
        return self.num_vehicles / self.length

    @property
    def speed(self):
        on = (self.W._veh("link") == self.id) & (self.W._veh("state") == C.VS_RUN)
        return float(self.W._veh("v")[on].mean()) if on.any() else self.u

    @property
    def flow(self):
        return self.density * self.speed

    def _idx(self, t):
        n = self.W.TSIZE
        return min(max(int(t / self.W.DELTAT), 0), n - 1)

    def arrival_count(self, t):
        return self.cum_arrival[self._idx(t)]

    def departure_count(self, t):
        return self.cum_departure[self._idx(t)]

    def num_vehicles_t(self, t):
        return self.arrival_count(t) - self.departure_count(t)

    def instant_travel_time(self, t):
        return self.traveltime_instant[self._idx(t)]

    def actual_travel_time(self, t):
        return self.traveltime_actual[self._idx(t)]

    def average_speed(self, t):
        tt = self.actual_travel_time(t)
        return self.length / tt if tt > 0 else self.u

    # uxsim_cpp_wrapper.py:386-445: derived link measures
    @property
    def num_vehicles_queue(self):
        """Vehicles on the link slower than its free-flow speed (Link::count_vehicles_in_queue, traffi.cpp:589-597)."""
        on_link = self.W._veh("link") == self.id
        return int(np.count_nonzero(on_link & (self.W._veh("v") < self.u))) * self.W.DELTAN

    def inflow_between(self, t0, t1):
        return (self.arrival_count(t1) - self.arrival_count(t0)) / (t1 - t0) if t1 != t0 else 0

    def average_travel_time_between(self, t0, t1):
        t_list = [t for t in self.vehicles_enter_log.keys() if t0 <= t < t1]
        if len(t_list) == 0:
            return self.length / self.u
        return np.average([self.actual_travel_time(t) for t in t_list])

    def average_density(self, t):
        return self.num_vehicles_t(t) / self.length

    def average_flow(self, t):
        return self.average_density(t) * self.average_speed(t)

    def change_free_flow_speed(self, new_value):
        if new_value >= 0:
            self.W._E.set_link_param(self.id, C.LINK_VMAX, float(new_value))
            self.u = self.free_flow_speed = float(new_value)
            self.w = 1 / self.tau / self.kappa
            self.capacity = self.u * self.w * self.kappa / (self.u + self.w)
            self.W._invalidate()

    def change_jam_density(self, new_value):
        if new_value >= 0:
            self.W._E.set_link_param(self.id, C.LINK_KAPPA, float(new_value))
            self.kappa = self.jam_density = float(new_value)
            self.w = 1 / self.tau / self.kappa
            self.capacity = self.u * self.w * self.kappa / (self.u + self.w)
            self.delta = 1 / self.kappa
            self.delta_per_lane = self.delta * self.number_of_lanes
            self.W._invalidate()


class CudaVehicle:
    """Lazy proxy of one platoon (uxsim_cpp_wrapper.py:486-725).  All state is read from the world's cached arrays."""

    __slots__ = ("W", "id", "name", "orig", "dest", "color", "attribute", "user_attribute", "user_function", "mode", "dest_list",
                 "node_event", "route_pref", "links_prefer", "links_avoid", "specified_route", "route_choice_principle", "link_old",
                 "_log_cache")  # (_log_cache: callers written against CppVehicle reset it to force a re-read; logs are re-read per state here anyway)

    def __repr__(self):
        return f"<Vehicle {self.name}>"

    @property
    def state(self):
        return _STATE_NAMES[int(self.W._veh("state")[self.id])]

    @state.setter
    def state(self, value):
        """uxsim_cpp_wrapper.py:517-519.  What callers do with it is take a platoon out before the run (`veh.state = "abort"`, e.g. to
        measure the externality of one trip); the engine then never releases it from its origin (UX_OPT_ABORT_VEHICLE)."""
        name = value if isinstance(value, str) else _STATE_NAMES[int(value)]
        if name != "abort":
            raise NotImplementedError("only state = 'abort' (before the simulation starts) can be set in CUDA mode")
        self.W._E.set_option(C.OPT_ABORT_VEHICLE, self.id)
        self.W._invalidate()

    @property
    def departure_time(self):
        return int(self.W._veh("depart_ts")[self.id])

    @property
    def departure_time_in_second(self):
        return self.departure_time * self.W.DELTAT

    @property
    def arrival_time(self):
        return int(self.W._veh("arrival_ts")[self.id])

    @property
    def travel_time(self):
        tt = float(self.W._veh("travel_time")[self.id])
        return -1 if tt < 0 else tt

    @property
    def distance_traveled(self):
        return float(self.W._veh("distance")[self.id])

    @property
    def flag_trip_aborted(self):
        return int(self.W._veh("state")[self.id] == C.VS_ABORT)

    @property
    def link(self):
        l = int(self.W._veh("link")[self.id])
        return self.W.LINKS[l] if l >= 0 else None

    @property
    def x(self):
        return float(self.W._veh("x")[self.id])

    @property
    def v(self):
        return float(self.W._veh("v")[self.id])

    @property
    def lane(self):
        return int(self.W._veh("lane")[self.id])

    @property
    def link_arrival_time(self):  # uxsim_cpp_wrapper.py:564-566
        return float(self.W._veh("link_arrival_time")[self.id])

    def add_dest(self, dest, order=-1):  # uxsim_cpp_wrapper.py:709-718 (taxi mode keeps the list; the loop does not read it)
        dest = self.W.get_node(dest)
        if order == -1:
            self.dest_list.append(dest)
        else:
            self.dest_list.insert(order, dest)

    def add_dests(self, dests):
        for d in dests:
            self.add_dest(d)

    def enforce_route(self, route, set_avoid=False):
        links = [self.W.get_link(l) for l in route]
        self.specified_route = links
        self.links_prefer = list(links)
        self.W._E.set_vehicle_links(self.id, 0, [l.id for l in links])

    def set_links_prefer(self, links):  # uxsim_cpp_wrapper.py:695-707
        self.links_prefer = [self.W.get_link(l) for l in links]
        self.W._E.set_vehicle_links(self.id, 1, [l.id for l in self.links_prefer])

    def set_links_avoid(self, links):
        self.links_avoid = [self.W.get_link(l) for l in links]
        self.W._E.set_vehicle_links(self.id, 2, [l.id for l in self.links_avoid])

    def get_xy_coords(self, t=-1):
        l = self.link
        if l is None:
            return self.orig.x, self.orig.y
        r = self.x / l.length if l.length > 0 else 0
        return (l.start_node.x + r * (l.end_node.x - l.start_node.x), l.start_node.y + r * (l.end_node.y - l.start_node.y))

    # ---- per-step logs (uxsim_cpp_wrapper.py:575-664): slices of the world's flat arrays, home steps prepended ----
    def _log(self, key):
        return self.W._vehicle_log(self.id)[key]

    @property
    def log_t(self):
        return self._log("log_t")

    @property
    def log_x(self):
        return self._log("log_x")

    @property
    def log_v(self):
        return self._log("log_v")

    @property
    def log_s(self):
        return self._log("log_s")

    @property
    def log_lane(self):
        return self._log("log_lane")

    # The three list-valued logs are converted once per state of the run and kept next to the arrays they come from
    # (uxsim_cpp_wrapper.py:580-664 caches them the same way): analyzer.py indexes `veh.log_state[i]` inside loops over i.
    @property
    def log_state(self):
        d = self.W._vehicle_log(self.id)
        if "log_state_names" not in d:
            d["log_state_names"] = [_STATE_NAMES[s] if 0 <= s < 5 else "end" for s in d["log_state"]]
        return d["log_state_names"]

    @property
    def log_link(self):
        d = self.W._vehicle_log(self.id)
        if "log_link_objs" not in d:
            links = self.W.LINKS
            d["log_link_objs"] = [links[i] if 0 <= i < len(links) else -1 for i in d["log_link"]]
        return d["log_link_objs"]

    @property
    def log_t_link(self):
        d = self.W._vehicle_log(self.id)
        if "log_t_link" not in d:
            links = self.W.LINKS
            out = []
            for t, lid in zip(d["log_t_link_t"], d["log_t_link_id"]):
                lid = int(lid)
                out.append([t, "home" if lid == -1 else "end" if lid == -2 else links[lid] if 0 <= lid < len(links) else -1])
            d["log_t_link"] = out
        return d["log_t_link"]

    def traveled_route(self, include_arrival_time=True, include_departure_time=False):
        """uxsim_cpp_wrapper.py:666-681."""
        links, ts = [], []
        ltl = self.log_t_link
        for t, l in ltl:
            if l == "home" and include_departure_time:
                ts.append(t)
            if isinstance(l, CudaLink):
                ts.append(t)
                links.append(l)
        if ltl and include_arrival_time:
            t, l = ltl[-1]
            ts.append(t if l == "end" else -1)
        return CudaRoute(self.W, links, trust_input=True), ts


class CudaRoute:
    """uxsim_cpp_wrapper.py:729-765 (CppRoute)."""

    def __init__(self, W, links, name="", trust_input=False):
        self.W, self.name = W, name
        self.links = [W.get_link(l) for l in links]
        if not trust_input:
            for a, b in zip(self.links[:-1], self.links[1:]):
                if a.end_node != b.start_node:
                    raise ValueError(f"Route links not connected: {a.name} -> {b.name}")
        self.links_name = [l.name for l in self.links]

    def __repr__(self):
        return f"<Route {self.name}: {self.links_name}>"

    def __iter__(self):
        return iter(self.links)

    def __len__(self):
        return len(self.links)

    def __getitem__(self, i):
        return self.links[i]

    def __eq__(self, other):
        return self.links_name == other.links_name if hasattr(other, "links_name") else NotImplemented

    def actual_travel_time(self, t, return_details=False):
        tt, ct, det = 0, t, []
        for l in self.links:
            x = l.actual_travel_time(ct)
            tt += x
            ct += x
            det.append(x)
        return (tt, det) if return_details else tt


class CudaRouteChoice:
    """uxsim_cpp_wrapper.py:768-797 (CppRouteChoice)."""

    def __init__(self, W):
        self.W = W

    @property
    def dist(self):
        n = len(self.W.NODES)
        return self.W._E.get_array(C.A_ROUTE_DIST).reshape(n, n)

    @property
    def next(self):
        n = len(self.W.NODES)
        return self.W._E.get_array(C.A_ROUTE_NEXT).reshape(n, n)

    @property
    def route_pref(self):
        return self.W._E.get_array(C.A_ROUTE_PREF).reshape(len(self.W.NODES), len(self.W.LINKS))

    @property
    def dist_record(self):
        """uxsim_cpp_wrapper.py:782-797 (World::route_dist_record): timestep of a route update -> the all-pairs table it produced.  The
        engine keeps the edge costs of every update (8 B per node pair) instead of an N x N table per update and searches on demand
        (UX_OPT_ROUTE_TABLES_AT), so the mapping is lazy: a table is computed when it is first looked up."""
        return _DistRecord(self.W)


class _DistRecord(Mapping):
    def __init__(self, W):
        self.W, self._tables = W, {}

    def _steps(self):
        return range(0, max(self.W.T, 1) if self.W.T < self.W.TSIZE else self.W.TSIZE, self.W.DELTAT_ROUTE)

    def __getitem__(self, ts):
        if ts not in self._steps():
            raise KeyError(ts)
        key = (ts, self.W.T)
        if key not in self._tables:
            E, n = self.W._E, len(self.W.NODES)
            E.set_option(C.OPT_ROUTE_TABLES_AT, int(ts))
            try:
                self._tables[key] = E.get_array(C.A_ROUTE_DIST).reshape(n, n)
            finally:
                E.set_option(C.OPT_ROUTE_TABLES_AT, -1)
        return self._tables[key]

    def __iter__(self):
        return iter(self._steps())

    def __len__(self):
        return len(self._steps())


class _VehicleRegistry(Sequence):
    """All platoons of a world in creation order, WITHOUT a Python object per platoon (building 10^5-10^6 proxies
    is the wall-clock floor of the wrapper, uxsim_cpp_wrapper.py:1204-1251).  ``adddemand`` records one batch (first id, count,
    origin, destination, attribute, colours); a :class:`CudaVehicle` proxy is made the first time somebody asks for that
    platoon and kept from then on, so attributes a caller sets on it persist.  Indexing = platoon id."""

    def __init__(self, W):
        self.W, self.n = W, 0
        self._first, self._batch = [], []       # batch k covers ids [_first[k], _first[k] + count)
        self._made = {}                         # id -> CudaVehicle
        self._name, self._id_of_name = {}, {}   # only platoons whose name is not str(id)

    def add_batch(self, count, orig, dest, attribute, colors):
        if count > 0:
            self._first.append(self.n)
            self._batch.append((count, orig, dest, attribute, colors))
            self.n += count

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(self.n))]
        i = int(i)
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        v = self._made.get(i)
        if v is None:
            k = bisect.bisect_right(self._first, i) - 1
            _, orig, dest, attribute, colors = self._batch[k]
            c = colors[i - self._first[k]]
            v = CudaVehicle.__new__(CudaVehicle)
            v.W, v.id, v.name = self.W, i, self.name_of(i)
            v.orig, v.dest = orig, dest
            v.color = (float(c[0]), float(c[1]), float(c[2]))
            v.attribute, v.user_attribute, v.user_function = attribute, None, None
            v.mode, v.dest_list, v.node_event, v.route_pref = "single_trip", [], {}, {}
            v.links_prefer, v.links_avoid, v.specified_route, v.route_choice_principle, v.link_old = [], [], None, self.W.route_choice_principle, None
            self._made[i] = v
        return v

    def __iter__(self):
        return (self[i] for i in range(self.n))

    def name_of(self, i):
        return self._name.get(i) or str(i)

    def names(self):
        return [self._name.get(i) or str(i) for i in range(self.n)]

    def id_of(self, name):
        """Platoon id of a name, or -1."""
        i = self._id_of_name.get(name)
        if i is not None:
            return i
        if isinstance(name, str) and name.isdigit() and str(int(name)) == name and int(name) < self.n and int(name) not in self._name:
            return int(name)
        return -1

    def rename(self, i, name):
        self._name[i] = name
        self._id_of_name[name] = i
        if i in self._made:
            self._made[i].name = name


class _VehicleDict(Mapping):
    """``W.VEHICLES`` / ``VEHICLES_LIVING`` / ``VEHICLES_RUNNING``: name -> vehicle in creation order over a
    :class:`_VehicleRegistry` (all platoons, or the given ids), read-only like the reference's use of them after set-up."""

    def __init__(self, reg, ids=None):
        self._reg, self._ids = reg, ids
        self._idset = None

    def __len__(self):
        return len(self._reg) if self._ids is None else len(self._ids)

    def __iter__(self):
        r = self._reg
        return (r.name_of(int(i)) for i in (range(len(r)) if self._ids is None else self._ids))

    def __contains__(self, name):
        i = self._reg.id_of(name)
        if i < 0:
            return False
        if self._ids is None:
            return True
        if self._idset is None:
            self._idset = set(int(k) for k in self._ids)
        return i in self._idset

    def __getitem__(self, name):
        if name not in self:
            raise KeyError(name)
        return self._reg[self._reg.id_of(name)]

    def values(self):
        r = self._reg
        return [r[int(i)] for i in (range(len(r)) if self._ids is None else self._ids)]

    def items(self):
        return [(v.name, v) for v in self.values()]


class CudaWorld:
    """uxsim_cpp_wrapper.py:800-1533 (CppWorld), in front of libuxsim_b200.so."""

    def __init__(self, name="", deltan=5, reaction_time=1, duo_update_time=600, duo_update_weight=0.5, duo_noise=0.01,
                 route_choice_principle="homogeneous_DUO", route_choice_update_gradual=False, instantaneous_TT_timestep_interval=5,
                 eular_dt=120, eular_dx=100, random_seed=None, print_mode=1, save_mode=1, show_mode=0, show_progress=1,
                 show_progress_deltat=600, tmax=None, vehicle_logging_timestep_interval=1, reduce_memory_delete_vehicle_route_pref=False,
                 hard_deterministic_mode=False, no_cyclic_routing=False, adjust_node_capacity=False, meta_data=None,
                 user_attribute=None, user_function=None, threads=1, num_replicas=1, device=-1, all_destinations=False):
        if not isinstance(threads, int) or isinstance(threads, bool) or (threads < 1 and threads != -1):  # uxsim_cpp_wrapper.py:852-853
            raise ValueError(f"threads must be a positive integer or -1 (all cores), got {threads!r}.")
        self.W = self
        self.rng = np.random.default_rng(seed=random_seed)
        self.random_seed = random_seed
        self.TMAX, self.DELTAN, self.REACTION_TIME = tmax, deltan, reaction_time
        self.DUO_UPDATE_TIME, self.DUO_UPDATE_WEIGHT, self.DUO_NOISE, self.EULAR_DT = duo_update_time, duo_update_weight, duo_noise, eular_dt
        self.DELTAT = self.REACTION_TIME * self.DELTAN
        self.DELTAT_ROUTE = max(1, int(self.DUO_UPDATE_TIME / self.DELTAT))
        self._veh_by_index = _VehicleRegistry(self)
        self.VEHICLES, self.VEHICLES_LIVING, self.VEHICLES_RUNNING = _VehicleDict(self._veh_by_index), _VehicleDict(self._veh_by_index), _VehicleDict(self._veh_by_index, [])
        self.NODES, self.LINKS, self.NODES_NAME_DICT, self.LINKS_NAME_DICT = [], [], {}, {}
        self.vehicle_logging_timestep_interval = vehicle_logging_timestep_interval
        self._log_mode = vehicle_logging_timestep_interval > 0
        self.route_choice_principle = route_choice_principle
        self.route_choice_update_gradual = route_choice_update_gradual
        self.instantaneous_TT_timestep_interval = min(int(instantaneous_TT_timestep_interval), self.DELTAT_ROUTE)
        self.show_progress = show_progress
        self.show_progress_deltat_timestep = int(show_progress_deltat / self.DELTAT)
        self.name = name
        self.hard_deterministic_mode, self.no_cyclic_routing, self.adjust_node_capacity = hard_deterministic_mode, no_cyclic_routing, adjust_node_capacity
        self.meta_data = meta_data if meta_data is not None else {}
        self.network_info, self.demand_info = ddict(list), ddict(list)
        self.finalized = 0
        self.world_start_time = time.time()
        self.print_mode, self.save_mode, self.show_mode = print_mode, save_mode, show_mode
        self.print = print if print_mode else (lambda *a, **k: None)
        self.user_attribute, self.user_function = user_attribute, user_function
        self.num_threads = threads
        seed = random_seed if random_seed is not None else int(np.random.default_rng().integers(0, 2**31 - 1))
        # the horizon is fixed at finalize time (set_t_max), like CppWorld._ensure_cpp_world (878-898)
        self._E = C.EngineWorld(
            C.load_product(), t_max=float(tmax if tmax is not None else 7200), delta_n=float(deltan), tau=float(reaction_time),
            duo_update_time=float(duo_update_time), duo_update_weight=float(duo_update_weight), route_choice_uncertainty=float(duo_noise),
            random_seed=int(seed), vehicle_log_mode=vehicle_logging_timestep_interval > 0, hard_deterministic_mode=hard_deterministic_mode,
            route_choice_update_gradual=route_choice_update_gradual, no_cyclic_routing=no_cyclic_routing,
            adjust_node_capacity=adjust_node_capacity, instantaneous_TT_timestep_interval=int(instantaneous_TT_timestep_interval),
            num_replicas=num_replicas, device=device)
        if all_destinations:  # keep route preferences for every node (needed to add demand to NEW destinations mid-run)
            self._E.set_option(C.OPT_ALL_DESTINATIONS, 1)
        self._cache = {}
        self._simulation_done = False
        self.T, self.TIME = 0, 0

    # ---- result cache ------------------------------------------------------------------------------------
    def _invalidate(self):
        self._cache = {}

    _SERIES = {"cum_arrival": C.A_CUM_ARRIVAL, "cum_departure": C.A_CUM_DEPARTURE, "traveltime_instant": C.A_TRAVELTIME_INSTANT,
               "traveltime_actual": C.A_TRAVELTIME_ACTUAL, "signal_log": C.A_SIGNAL_LOG}
    _VEH = {"state": C.A_VEH_STATE, "arrival_ts": C.A_VEH_ARRIVAL_TS, "depart_ts": C.A_VEH_DEPART_TS, "travel_time": C.A_VEH_TRAVEL_TIME,
            "distance": C.A_VEH_DISTANCE, "link": C.A_VEH_LINK, "x": C.A_VEH_X, "v": C.A_VEH_V, "lane": C.A_VEH_LANE,
            "link_arrival_time": C.A_VEH_LINK_ARRIVAL_TIME}

    def _series(self, key):
        if key not in self._cache:
            rows = len(self.NODES) if key == "signal_log" else len(self.LINKS)
            self._cache[key] = self._E.get_array(self._SERIES[key]).reshape(rows, self.TSIZE)
        return self._cache[key]

    def _vehicle_log(self, vid):
        """Per-platoon view of the flat SoA logs (uxsim_cpp_wrapper.py:1309-1367), fetched once per state of the run."""
        if not self._log_mode:  # what the engine was created with: a caller may lower the attribute afterwards (the reference's Python solvers do), the records are still kept
            raise RuntimeError("vehicle logging is off (vehicle_logging_timestep_interval=-1)")
        if "logs" not in self._cache:
            E = self._E
            flat = {k: E.get_array(getattr(C, "A_" + k.upper())) for k in
                    ("log_offsets", "log_t", "log_x", "log_v", "log_state", "log_s", "log_lane", "log_link", "log_n_missing", "ltl_offsets", "ltl_t", "ltl_id")}
            if isinstance(self.DELTAT, (int, np.integer)):
                flat["log_t"] = flat["log_t"].astype(np.int64)
            self._cache["logs"] = flat
            self._cache["logs_per_vehicle"] = {}
        per = self._cache["logs_per_vehicle"]
        if vid not in per and vid + 1 >= len(self._cache["logs"]["log_offsets"]):
            # a platoon added after the last exec_simulation chunk: the engine takes it in at the next chunk, its logs are empty so far
            f = self._cache["logs"]
            per[vid] = dict(log_t=f["log_t"][:0], log_x=f["log_x"][:0], log_v=f["log_v"][:0], log_s=f["log_s"][:0], log_state=f["log_state"][:0],
                            log_lane=f["log_lane"][:0], log_link=f["log_link"][:0], log_t_link_t=f["ltl_t"][:0], log_t_link_id=f["ltl_id"][:0])
        if vid not in per:
            f = self._cache["logs"]
            s, e = int(f["log_offsets"][vid]), int(f["log_offsets"][vid + 1])
            ls, le = int(f["ltl_offsets"][vid]), int(f["ltl_offsets"][vid + 1])
            nm = int(f["log_n_missing"][vid])
            d = dict(log_t=f["log_t"][s:e], log_x=f["log_x"][s:e], log_v=f["log_v"][s:e], log_s=f["log_s"][s:e], log_state=f["log_state"][s:e],
                     log_lane=f["log_lane"][s:e], log_link=f["log_link"][s:e], log_t_link_t=f["ltl_t"][ls:le], log_t_link_id=f["ltl_id"][ls:le])
            if nm > 0:  # home steps before the first recorded entry
                d["log_t"] = np.concatenate([np.arange(nm, dtype=d["log_t"].dtype) * self.DELTAT, d["log_t"]])
                for k in ("log_x", "log_v", "log_s"):
                    d[k] = np.concatenate([np.full(nm, -1.0), d[k]])
                d["log_state"] = np.concatenate([np.zeros(nm, np.int32), d["log_state"]])
                for k in ("log_lane", "log_link"):
                    d[k] = np.concatenate([np.full(nm, -1, np.int32), d[k]])
            per[vid] = d
        return per[vid]

    def _build_vehicles_enter_log(self):
        """uxsim_cpp_wrapper.py:1377-1398."""
        for l in self.LINKS:
            l.vehicles_enter_log = {}
        E = self._E
        for lid, t, vi in zip(E.get_array(C.A_ENTER_LINK), E.get_array(C.A_ENTER_TIME), E.get_array(C.A_ENTER_VEH)):
            self.LINKS[int(lid)].vehicles_enter_log[float(t)] = self._veh_by_index[int(vi)]

    def _veh(self, key):
        k = "veh_" + key
        if k not in self._cache:
            self._cache[k] = self._E.get_array(self._VEH[key])
        return self._cache[k]

    # ---- scenario definition -----------------------------------------------------------------------------
    def addNode(self, name, x, y, signal=[0], signal_offset=0, signal_offset_old=None, flow_capacity=None, number_of_lanes=None,
                auto_rename=False, attribute=None, user_attribute=None, user_function=None):
        return CudaNode(self, name, x, y, signal=signal, signal_offset=signal_offset, signal_offset_old=signal_offset_old,
                        flow_capacity=flow_capacity, number_of_lanes=number_of_lanes, auto_rename=auto_rename, attribute=attribute,
                        user_attribute=user_attribute, user_function=user_function)

    def addLink(self, name, start_node, end_node, length, free_flow_speed=20, jam_density=0.2, jam_density_per_lane=None,
                number_of_lanes=1, merge_priority=1, signal_group=[0], capacity_out=None, capacity_in=None, congestion_pricing=None,
                eular_dx=None, attribute=None, user_attribute=None, user_function=None, auto_rename=False):
        return CudaLink(self, name, start_node, end_node, length, free_flow_speed=free_flow_speed, jam_density=jam_density,
                        jam_density_per_lane=jam_density_per_lane, number_of_lanes=number_of_lanes, merge_priority=merge_priority,
                        signal_group=signal_group, capacity_out=capacity_out, capacity_in=capacity_in, congestion_pricing=congestion_pricing,
                        eular_dx=eular_dx, attribute=attribute, user_attribute=user_attribute, user_function=user_function, auto_rename=auto_rename)

    def _register(self, n_new, orig, dest, attribute=None):
        self._veh_by_index.add_batch(n_new, orig, dest, attribute, self.rng.random(size=(n_new, 3)))

    def addVehicle(self, orig, dest, departure_time, name=None, route_pref=None, route_choice_principle=None, mode="single_trip",
                   links_prefer=[], links_avoid=[], trip_abort=1, departure_time_is_time_step=0, attribute=None, user_attribute=None,
                   user_function=None, auto_rename=False, direct_call=True):
        if mode == "taxi":
            raise NotImplementedError("taxi mode is not supported in CUDA mode")  # as cpp mode, uxsim_cpp_wrapper.py:929-931
        o, d = self.get_node(orig), self.get_node(dest)
        ts = int(departure_time) if departure_time_is_time_step else int(departure_time / self.DELTAT)
        self._E.add_vehicles([o.id], [d.id], [ts])
        self._register(1, o, d, attribute)
        v = self._veh_by_index[-1]
        if name is not None:
            name = str(name)
            if name in self.VEHICLES and name != v.name:
                if not auto_rename:
                    raise ValueError(f"Vehicle name {name} already used by another vehicle. Please specify a unique name.")
                name = name + "_renamed" + "".join(self.rng.choice(list(string.ascii_letters + string.digits), size=8))
            self._veh_by_index.rename(v.id, name)
        if links_prefer:
            v.links_prefer = [self.get_link(l) for l in links_prefer]
            self._E.set_vehicle_links(v.id, 1, [l.id for l in v.links_prefer])
        if links_avoid:
            v.links_avoid = [self.get_link(l) for l in links_avoid]
            self._E.set_vehicle_links(v.id, 2, [l.id for l in v.links_avoid])
        v.user_attribute, v.user_function = user_attribute, user_function
        self._invalidate()
        return v

    def adddemand(self, orig, dest, t_start, t_end, flow=-1, volume=-1, attribute=None, direct_call=True):
        """uxsim.py:2021-2053; the FP64 accumulation happens in ux_add_demand (traffi.cpp:1910-1941)."""
        if volume > 0:
            flow = volume / (t_end - t_start)
        o, d = self.get_node(orig), self.get_node(dest)
        n = self._E.add_demand(o.id, d.id, float(t_start), float(t_end), float(flow))
        self._register(n, o, d, attribute)
        self.demand_info["adddemand"].append(dict(orig=o.name, dest=d.name, t_start=t_start, t_end=t_end, flow=flow, volume=volume))
        self._invalidate()

    def adddemand_point2point(self, x_orig, y_orig, x_dest, y_dest, t_start, t_end, flow=-1, volume=-1, attribute=None, direct_call=True):
        self.adddemand(self.get_nearest_node(x_orig, y_orig), self.get_nearest_node(x_dest, y_dest), t_start, t_end, flow, volume, attribute)

    def adddemand_area2area(self, x_orig, y_orig, radious_orig, x_dest, y_dest, radious_dest, t_start, t_end, flow=-1, volume=-1,
                            attribute=None, direct_call=True):
        warnings.warn("`adddemand_area2area()` is not recommended. Use `adddemand_area2area2()`.", FutureWarning)
        origs = [o for o in self.get_nodes_in_area(x_orig, y_orig, radious_orig) if o.outlinks] or [self.get_nearest_node(x_orig, y_orig)]
        dests = [d for d in self.get_nodes_in_area(x_dest, y_dest, radious_dest) if d.inlinks] or [self.get_nearest_node(x_dest, y_dest)]
        self._nodes2nodes(origs, dests, t_start, t_end, flow, volume, attribute)

    def adddemand_nodes2nodes(self, origs, dests, t_start, t_end, flow=-1, volume=-1, attribute=None, direct_call=True):
        warnings.warn("`adddemand_nodes2nodes()` is not recommended. Use `adddemand_nodes2nodes2()`.", FutureWarning)
        origs = [self.get_node(o) for o in origs if self.get_node(o).outlinks]
        dests = [self.get_node(d) for d in dests if self.get_node(d).inlinks]
        self._nodes2nodes(origs, dests, t_start, t_end, flow, volume, attribute)

    def _nodes2nodes(self, origs, dests, t_start, t_end, flow, volume, attribute):
        if flow != -1:
            flow = flow / (len(origs) * len(dests))
        if volume != -1:
            volume = volume / (len(origs) * len(dests))
        for o in origs:
            for d in dests:
                self.adddemand(o, d, t_start, t_end, flow, volume, attribute)

    def adddemand_area2area2(self, x_orig, y_orig, radious_orig, x_dest, y_dest, radious_dest, t_start, t_end, flow=-1, volume=-1,
                             attribute=None, direct_call=True):
        origs = [o for o in self.get_nodes_in_area(x_orig, y_orig, radious_orig) if o.outlinks] or [self.get_nearest_node(x_orig, y_orig)]
        dests = [d for d in self.get_nodes_in_area(x_dest, y_dest, radious_dest) if d.inlinks] or [self.get_nearest_node(x_dest, y_dest)]
        self.adddemand_nodes2nodes2(origs, dests, t_start, t_end, flow, volume, attribute)

    def adddemand_nodes2nodes2(self, origs, dests, t_start, t_end, flow=-1, volume=-1, attribute=None, direct_call=True):
        """uxsim.py:2262-2310: random OD assignment, then one platoon per draw (bulk-registered)."""
        origs = [self.get_node(o) for o in origs if self.get_node(o).outlinks]
        dests = [self.get_node(d) for d in dests if self.get_node(d).inlinks]
        if flow >= 0 and volume == -1:
            volume = flow * (t_end - t_start)
        size = int(volume / self.DELTAN)
        ts = np.linspace(t_start, t_end, size)
        os_ = self.rng.choice(origs, size=size)
        ds_ = self.rng.choice(dests, size=size)
        for t, o, d in zip(ts, os_, ds_):
            d2 = d if o != d else self.rng.choice([dd for dd in dests if dd != d])
            self.addVehicle(o, d2, t, attribute=attribute)

    # ---- simulation ----------------------------------------------------------------------------------------
    def finalize_scenario(self, tmax=None):
        """uxsim_cpp_wrapper.py:1087-1121."""
        if self.TMAX is None:
            if tmax is None:
                m = self._E.get_int(C.I_MAX_DEPART_TS) * self.DELTAT
                self.TMAX = (m // 1800 + 2) * 1800
            else:
                self.TMAX = tmax
        self._E.set_t_max(float(self.TMAX))
        self.T, self.TIME = 0, 0
        self.TSIZE = int(self.TMAX / self.DELTAT)
        self.Q_AREA = ddict(lambda: np.zeros(int(self.TMAX / self.EULAR_DT)))
        self.K_AREA = ddict(lambda: np.zeros(int(self.TMAX / self.EULAR_DT)))
        for l in self.LINKS:
            l.init_after_tmax_fix()
            if l.congestion_pricing is not None:
                self._E.set_toll_timeseries(l.id, [float(l.congestion_pricing(t * self.DELTAT)) for t in range(self.TSIZE)])
        self.ADJ_MAT = np.zeros([len(self.NODES), len(self.NODES)])
        self.ADJ_MAT_LINKS, self.NODE_PAIR_LINKS = dict(), dict()
        for l in self.LINKS:
            self.ADJ_MAT[l.start_node.id, l.end_node.id] = 1
            self.ADJ_MAT_LINKS[l.start_node.id, l.end_node.id] = l
            self.NODE_PAIR_LINKS[l.start_node.name, l.end_node.name] = l
        self._E.initialize_adj_matrix()
        if self.adjust_node_capacity:  # sync the wrapper-side attributes (uxsim_cpp_wrapper.py:1101-1104)
            for node in self.NODES:
                node.adjust_node_capacity()
        self.ROUTECHOICE = CudaRouteChoice(self)
        self.analyzer = _make_analyzer(self)
        self.finalized = 1
        self.sim_start_time = time.time()
        self.print("simulating...")

    def exec_simulation(self, until_t=None, duration_t=None, duration_t2=None):
        """uxsim_cpp_wrapper.py:1165-1202 -- same end-step arithmetic, one FFI crossing."""
        if not self.finalized:
            self.finalize_scenario()
        start_ts = self.T
        if until_t is not None:
            end_ts = int(until_t / self.DELTAT)
        elif duration_t2 is not None:
            end_ts = start_ts + int(duration_t2 / self.DELTAT) - 1
        elif duration_t is not None:
            end_ts = start_ts + int(duration_t / self.DELTAT)
        else:
            end_ts = self.TSIZE
        end_ts = min(end_ts, self.TSIZE)
        if start_ts == end_ts == self.TSIZE:
            self.simulation_terminated()
            return 1
        if end_ts < start_ts:
            raise Exception("exec_simulation error: Simulation duration is not positive. Check until_t or duration_t or duration_t2")
        self._E.main_loop(-1.0, float(end_ts * self.DELTAT))
        self._invalidate()
        self.T = self._E.get_int(C.I_TIMESTEP)
        self.TIME = self.T * self.DELTAT
        self._sync_registries()
        if self.T >= self.TSIZE:
            self.simulation_terminated()
            return 1
        return 0

    def _sync_registries(self):
        st = self._veh("state")
        self.VEHICLES_LIVING = _VehicleDict(self._veh_by_index, np.nonzero(st < C.VS_END)[0])
        self.VEHICLES_RUNNING = _VehicleDict(self._veh_by_index, np.nonzero(st == C.VS_RUN)[0])

    def simulation_terminated(self):
        self.print(" simulation finished")
        self._simulation_done = True
        if self.vehicle_logging_timestep_interval > 0 and not getattr(self, "_skip_log_on_terminate", False):
            self._build_vehicles_enter_log()
        self.analyzer.basic_analysis()

    def check_simulation_ongoing(self):
        return True if not self.finalized else self.T <= self.TSIZE - 1

    # ---- lookups ---------------------------------------------------------------------------------------------
    def get_node(self, node):
        if node is None or isinstance(node, CudaNode):
            return node
        if isinstance(node, (str, np.str_)):
            return self.NODES_NAME_DICT.get(str(node))
        if hasattr(node, "name"):
            return self.NODES_NAME_DICT.get(node.name)
        raise Exception(f"'{node}' is not Node in this World")

    def get_link(self, link):
        if link is None:
            return None
        if isinstance(link, CudaLink):
            return link if link.W is self else self.LINKS_NAME_DICT.get(link.name)
        if isinstance(link, (str, np.str_)):
            return self.LINKS_NAME_DICT.get(str(link))
        if hasattr(link, "name"):
            return self.LINKS_NAME_DICT.get(link.name)
        raise Exception(f"'{link}' is not Link in this World")

    def get_nearest_node(self, x, y):
        return min(self.NODES, key=lambda n: (n.x - x) ** 2 + (n.y - y) ** 2) if self.NODES else None

    def get_nodes_in_area(self, x, y, r):
        return [n for n in self.NODES if (n.x - x) ** 2 + (n.y - y) ** 2 < r ** 2]

    def defRoute(self, links, name="", trust_input=False):
        return CudaRoute(self, links, name=name, trust_input=trust_input)

    # ---- I/O (uxsim_cpp_wrapper.py:1435-1472) ---------------------------------------------------------------
    def load_scenario_from_csv(self, fname_node, fname_link, fname_demand, tmax=None):
        self.generate_Nodes_from_csv(fname_node)
        self.generate_Links_from_csv(fname_link)
        self.generate_demand_from_csv(fname_demand)
        if tmax is not None:
            self.TMAX = tmax

    def generate_Nodes_from_csv(self, fname):
        with open(fname) as f:
            for r in csv.reader(f):
                if r[1] != "x":
                    self.addNode(r[0], float(r[1]), float(r[2]))

    def generate_Links_from_csv(self, fname):
        with open(fname) as f:
            for r in csv.reader(f):
                if r[3] != "length":
                    self.addLink(r[0], r[1], r[2], length=float(r[3]), free_flow_speed=float(r[4]), jam_density=float(r[5]), merge_priority=float(r[6]))

    def generate_demand_from_csv(self, fname):
        with open(fname) as f:
            for r in csv.reader(f):
                if r[2] != "start_t":
                    if len(r) > 5 and r[5] != "":
                        self.adddemand(r[0], r[1], float(r[2]), float(r[3]), float(r[4]), float(r[5]))
                    else:
                        self.adddemand(r[0], r[1], float(r[2]), float(r[3]), float(r[4]))

    def save_scenario(self, fname, network=True, demand=True):
        raise NotImplementedError("save_scenario is not supported in CUDA mode (as in C++ mode).")

    def load_scenario(self, fname, network=True, demand=True):
        """scenario_reader_writer.py:152-224: read a `.uxsim_scenario` file (a pickled dict of the recorded addNode / addLink /
        adddemand* keyword arguments) into this world.  The reference's C++ mode does not support this (uxsim_cpp_wrapper.py:1529-1533);
        here the plain `adddemand` records -- 8 406 of them in dat/chicago_sketch.uxsim_scenario -- go through the vectorised ingest
        (one FFI call, ux_add_demands), everything else through the same methods the reference's loader calls."""
        import pickle

        with open(fname, "rb") as f:
            try:
                dat = pickle.load(f)
            except Exception:  # files written by dill with objects plain pickle cannot resolve
                import dill

                f.seek(0)
                dat = dill.load(f)
        self.print(f"loading scenario from '{fname}'")
        meta = dat.get("meta_data")
        if meta:
            if isinstance(meta, dict):
                for key in meta:
                    self.print("", key, ":", meta[key])
            else:
                self.print("", meta)
        if network:
            for n in dat["Nodes"]:
                self.addNode(**n)
            for l in dat["Links"]:
                self.addLink(**l)
            self.print(" Number of loaded nodes:", len(dat["Nodes"]))
            self.print(" Number of loaded links:", len(dat["Links"]))
        if demand:
            for demand_type in dat:
                if demand_type == "adddemand":
                    self._adddemand_batch(dat[demand_type])
                elif demand_type in ("adddemand_point2point", "adddemand_area2area", "adddemand_nodes2nodes", "adddemand_area2area2", "adddemand_nodes2nodes2"):
                    for dem in dat[demand_type]:
                        getattr(self, demand_type)(**dem)
                else:
                    continue
                self.print(f" Number of loaded `{demand_type}`s:", len(dat[demand_type]))

    def _adddemand_batch(self, demands):
        """Many `adddemand(**kwargs)` records at once: one ux_add_demands call creates the platoons of all of them in order
        (the same FP64 accumulation per record), the registry gets one batch per run of equal (orig, dest)."""
        if not demands:
            return
        if any(d.get("attribute") is not None for d in demands):  # per-record attributes: the per-call path keeps them apart
            for d in demands:
                self.adddemand(**d)
            return
        nodes = [(self.get_node(d["orig"]), self.get_node(d["dest"])) for d in demands]
        t0 = np.array([d["t_start"] for d in demands], np.float64)
        t1 = np.array([d["t_end"] for d in demands], np.float64)
        flow = np.array([(d.get("volume", -1) / (d["t_end"] - d["t_start"])) if d.get("volume", -1) > 0 else d.get("flow", -1) for d in demands], np.float64)
        p0 = self._E.get_int(C.I_NUM_VEHICLES)
        self._E.add_demands([o.id for o, _ in nodes], [d.id for _, d in nodes], t0, t1, flow)
        p1 = self._E.get_int(C.I_NUM_VEHICLES)
        if p1 > p0:
            vo, vd = self._E.get_array(C.A_VEH_ORIG)[p0:p1], self._E.get_array(C.A_VEH_DEST)[p0:p1]
            cut = np.flatnonzero((np.diff(vo) != 0) | (np.diff(vd) != 0)) + 1
            for a, b in zip(np.r_[0, cut], np.r_[cut, p1 - p0]):
                self._register(int(b - a), self.NODES[int(vo[a])], self.NODES[int(vd[a])], None)
        for (o, d), dem, f in zip(nodes, demands, flow):
            self.demand_info["adddemand"].append(dict(orig=o.name, dest=d.name, t_start=dem["t_start"], t_end=dem["t_end"], flow=float(f), volume=dem.get("volume", -1)))
        self._invalidate()

    def copy(self):
        raise NotImplementedError("copy() is not supported in CUDA mode (as in C++ mode)")

    def save(self, fname):
        raise NotImplementedError("save() is not supported in CUDA mode (as in C++ mode)")

    def show_network(self, width=1, left_handed=1, figsize=(6, 6), network_font_size=10, node_size=6, show_id=True):
        """uxsim_cpp_wrapper.py:1496-1527.  Drawing is outside the simulation loop (out of scope): the sketch below needs
        matplotlib; like the reference's `catch_exceptions_and_warn`, a failure to draw is a warning, never an error."""
        try:
            import matplotlib.pyplot as plt

            os.makedirs(f"out{self.name}", exist_ok=True)
            fig = plt.figure(figsize=figsize)
            ax = fig.add_subplot(111, aspect="equal")
            side = 1.0 if left_handed else -1.0
            for l in self.LINKS:  # a link is drawn bent towards its driving side, so that the two directions of a road do not overlap
                a, b = l.start_node, l.end_node
                ox, oy = side * (a.y - b.y) * 0.05, side * (b.x - a.x) * 0.05
                px = [a.x, (2 * a.x + b.x) / 3 + ox, (a.x + 2 * b.x) / 3 + ox, b.x]
                py = [a.y, (2 * a.y + b.y) / 3 + oy, (a.y + 2 * b.y) / 3 + oy, b.y]
                ax.plot(px, py, "gray", lw=width, zorder=6, solid_capstyle="butt")
                if network_font_size > 0:
                    ax.text(px[1], py[1], f"{l.id}: {l.name}" if show_id else l.name, c="b", zorder=20, fontsize=network_font_size)
            for n in self.NODES:
                ax.plot(n.x, n.y, "o", c="gray", ms=node_size, zorder=10)
                if network_font_size > 0:
                    ax.text(n.x, n.y, f"{n.id}: {n.name}" if show_id else n.name, c="g", ha="center", va="top", zorder=20, fontsize=network_font_size)
            fig.tight_layout()
            if self.save_mode:
                fig.savefig(f"out{self.name}/network.png")
            if self.show_mode:
                plt.show()
            else:
                plt.close("all")
        except Exception as e:  # noqa: BLE001
            warnings.warn(f"show_network: {type(e).__name__}: {e}")

    def on_time(self, time_val):
        return self.T == int(time_val / self.DELTAT)

    def change_print_mode(self, print_mode):
        self.print_mode = print_mode
        self.print = print if print_mode else (lambda *a, **k: None)

    def print_scenario_stats(self):
        print("simulation setting:" if self.finalized else "simulation setting (not finalized):")
        print(" scenario name:", self.name)
        print(" simulation duration:\t", self.TMAX, "s")
        print(" number of vehicles:\t", len(self.VEHICLES) * self.DELTAN, "veh")
        print(" total road length:\t", sum(l.length for l in self.LINKS), "m")
        print(" timestep size:\t", self.DELTAT, "s")
        print(" platoon size:\t\t", self.DELTAN, "veh")
        print(" number of platoons:\t", len(self.VEHICLES))
        print(" number of links:\t", len(self.LINKS))
        print(" number of nodes:\t", len(self.NODES))


def _make_analyzer(W):
    try:
        from uxsim.analyzer import Analyzer  # the reference's own, unmodified, when it is installed

        return Analyzer(W)
    except Exception:
        from .analysis import BasicAnalyzer

        return BasicAnalyzer(W)


class World:
    """Factory with the reference's selection idiom: ``World(..., cuda=True)`` -> :class:`CudaWorld`
    (uxsim.py:1671-1675 does the same for ``cpp=True``).  Without ``cuda=True`` this package has nothing to offer:
    it is a backend, not a second CPU engine."""

    def __new__(cls, *args, cuda=False, cpp=False, **kwargs):
        if cuda:
            return CudaWorld(*args, **kwargs)
        raise NotImplementedError("uxsim_b200 only provides the CUDA engine: use World(..., cuda=True), or the reference package "
                                  "`uxsim` for its Python / C++ engines.")
