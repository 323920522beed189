"""ctypes binding of the C-ABI declared in ``include/uxsim_b200.h``.

This is the only place the package touches native code.  ``load_product()`` loads
``uxsim_b200/libuxsim_b200.so`` (hand-written sm_100a CUDA + C++ host driver, built by
``__graft_entry__.build()``) and raises if it is missing: there is no CPU fallback.

The binding class is generic over the symbol prefix so that the test-suite can drive the two
test-infrastructure libraries that export the same ABI (``oracle/liboracle.so`` with prefix ``uxo_`` and
``oracle/_ref/libuxsim_ref.so`` with prefix ``uxr_``) through identical code; the package itself never
loads those.

Reference counterpart: the nanobind module ``uxsim_cpp`` (uxsim/trafficpp/bindings.cpp) as used by
uxsim/uxsim_cpp_wrapper.py:26-37.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "libuxsim_b200.so")

# ---- constants mirrored from include/uxsim_b200.h -------------------------------------------------
UX_OK = 0
UX_ERR_RUNTIME, UX_ERR_INVALID, UX_ERR_OUT_OF_RANGE, UX_ERR_CUDA, UX_ERR_CAPACITY = -1, -2, -3, -4, -5
VS_HOME, VS_WAIT, VS_RUN, VS_END, VS_ABORT = 0, 1, 2, 3, 4

LINK_VMAX, LINK_KAPPA, LINK_CAPACITY_OUT, LINK_CAPACITY_IN = 1, 2, 3, 4
LINK_CAPACITY_OUT_REMAIN, LINK_CAPACITY_IN_REMAIN, LINK_MERGE_PRIORITY, LINK_ROUTE_CHOICE_PENALTY = 5, 6, 7, 8

I_NUM_NODES, I_NUM_LINKS, I_NUM_VEHICLES, I_TIMESTEP, I_TOTAL_TIMESTEPS, I_NUM_REPLICAS = 1, 2, 3, 4, 5, 6
I_ROUTE_UPDATE_STEPS, I_PLATOON_UPDATES, I_TRIPS_COMPLETED, I_LOG_ENTRIES, I_KERNEL_LAUNCHES = 7, 8, 9, 10, 11
I_NUM_ACTIVE_DESTS, I_TRANSFER_LEVELS, I_LINK_EVENTS = 12, 13, 14
I_H2D_BYTES, I_D2H_BYTES, I_RING_SLOTS, I_MAX_DEPART_TS, I_KERNEL_LAUNCHES_OF = 15, 16, 17, 18, 100
K_NAMES = ("k_link_pre", "k_node", "k_update", "k_edge_cost", "k_sssp", "k_duo", "misc")
OPT_KERNEL_TIMING, OPT_ALL_DESTINATIONS, OPT_RECORD_TRAVERSALS, OPT_ROUTE_TABLES_AT, OPT_ABORT_VEHICLE = 1, 2, 3, 4, 5
D_DELTA_T, D_TIME, D_T_MAX, D_AVE_V_SUM, D_STAT_SAMPLE_COUNT, D_LAST_LOOP_MS, D_KERNEL_MS_OF = 1, 2, 3, 4, 5, 6, 100
D_UPLOAD_MS, D_GRAPH_BUILD_MS, D_LAST_LOOP_WALL_MS = 7, 8, 9
D_DTA_N_SWAP, D_DTA_POTENTIAL_N_SWAP, D_DTA_TOTAL_T_GAP, D_DTA_SWAP_MS, D_DTA_N_SWAP_OF = 10, 11, 12, 13, 1000

A_CUM_ARRIVAL, A_CUM_DEPARTURE, A_TRAVELTIME_INSTANT, A_TRAVELTIME_ACTUAL = 1, 2, 3, 4
A_VEH_STATE, A_VEH_ARRIVAL_TS, A_VEH_DEPART_TS, A_VEH_TRAVEL_TIME, A_VEH_DISTANCE = 10, 11, 12, 13, 14
A_VEH_ORIG, A_VEH_DEST, A_VEH_LINK, A_VEH_X, A_VEH_V, A_VEH_LANE, A_VEH_LINK_ARRIVAL_TIME = 15, 16, 17, 18, 19, 20, 21
A_SIGNAL_LOG, A_NODE_SIGNAL_PHASE, A_NODE_SIGNAL_T, A_NODE_FLOW_CAPACITY, A_NODE_NUM_LANES = 30, 31, 32, 33, 34
A_LINK_NUM_VEHICLES, A_LINK_CAPACITY_IN_REMAIN, A_LINK_CAPACITY_OUT_REMAIN = 40, 41, 42
A_LINK_STAT_COUNT, A_LINK_TRIPS_COMPLETED, A_LINK_AVE_V_SUM = 43, 44, 45
A_ROUTE_PREF, A_ROUTE_DIST, A_ROUTE_NEXT = 50, 51, 52
A_LOG_OFFSETS, A_LOG_T, A_LOG_X, A_LOG_V, A_LOG_STATE, A_LOG_S, A_LOG_LANE, A_LOG_LINK = 60, 61, 62, 63, 64, 65, 66, 67
A_LOG_N_MISSING, A_LTL_OFFSETS, A_LTL_T, A_LTL_ID = 68, 69, 70, 71
A_DTA_OD_ORIG, A_DTA_OD_DEST, A_DTA_OD_OFFSETS, A_DTA_ROUTE_OFFSETS, A_DTA_ROUTE_LINKS = 100, 101, 102, 103, 104
A_EDIE_OFFSETS, A_EDIE_NX, A_EDIE_TN, A_EDIE_DN = 120, 121, 122, 123
A_OD_ORIG, A_OD_DEST, A_OD_STATS, A_LINKC_STATS = 130, 131, 132, 133
A_DTA_RS_OFFSETS, A_DTA_RS_LINKS, A_DTA_RA_OFFSETS, A_DTA_RA_LINKS, A_DTA_COST_ACTUAL, A_DTA_T_GAP, A_DTA_CHOICE, A_DTA_RA_ENTRY_T = 110, 111, 112, 113, 114, 115, 116, 117
A_ENTER_LINK, A_ENTER_TIME, A_ENTER_VEH = 80, 81, 82

DT_F64, DT_I32, DT_I64 = 1, 2, 3
_NP_DTYPE = {DT_F64: np.float64, DT_I32: np.int32, DT_I64: np.int64}

_EXC = {
    UX_ERR_RUNTIME: RuntimeError,
    UX_ERR_INVALID: ValueError,
    UX_ERR_OUT_OF_RANGE: IndexError,
    UX_ERR_CUDA: RuntimeError,
    UX_ERR_CAPACITY: RuntimeError,
}

#: every symbol include/uxsim_b200.h declares (without prefix); tests check each library exports all of them
ABI_SYMBOLS = (
    "create_world", "destroy_world", "add_node", "add_link", "add_demand", "add_nodes", "add_links", "add_demands",
    "add_vehicles", "set_vehicle_links", "set_option",
    "batch_enforce_routes_csr", "set_t_max", "set_toll_timeseries", "initialize_adj_matrix", "main_loop",
    "check_simulation_ongoing", "set_link_param", "set_link_signal_group", "set_node_signal", "get_int",
    "get_double", "array_size", "get_array", "get_device_array", "ensemble_link_stats_device",
    "ensemble_link_stats", "last_error", "dta_enumerate_route_sets", "dta_set_route_sets", "dta_share_route_sets",
    "dta_enforce_routes", "dta_route_swap", "dta_next_day", "edie_state", "analysis_summary",
)


class WorldParams(C.Structure):
    """``ux_world_params`` (include/uxsim_b200.h)."""

    _fields_ = [
        ("t_max", C.c_double),
        ("delta_n", C.c_double),
        ("tau", C.c_double),
        ("duo_update_time", C.c_double),
        ("duo_update_weight", C.c_double),
        ("route_choice_uncertainty", C.c_double),
        ("random_seed", C.c_int64),
        ("print_mode", C.c_int32),
        ("vehicle_log_mode", C.c_int32),
        ("hard_deterministic_mode", C.c_int32),
        ("route_choice_update_gradual", C.c_int32),
        ("no_cyclic_routing", C.c_int32),
        ("adjust_node_capacity", C.c_int32),
        ("instantaneous_TT_timestep_interval", C.c_int32),
        ("num_replicas", C.c_int32),
        ("device", C.c_int32),
        ("num_threads", C.c_int32),
        ("reserved0", C.c_int32),
    ]


def _as_i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class CApi:
    """Typed access to one shared library exporting the uxsim_b200 C-ABI under ``prefix``."""

    def __init__(self, lib_path: str, prefix: str = "ux_"):
        if not os.path.exists(lib_path):
            raise ImportError(
                f"native library not found: {lib_path}. Build it with `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (needs nvcc). There is no CPU fallback."
            )
        self.path = lib_path
        self.prefix = prefix
        self.lib = C.CDLL(lib_path)
        p = C.c_void_p
        i32p, f64p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int64)

        def fn(name, restype, argtypes):
            f = getattr(self.lib, prefix + name)
            f.restype = restype
            f.argtypes = argtypes
            return f

        self.create_world = fn("create_world", p, [C.POINTER(WorldParams)])
        self.destroy_world = fn("destroy_world", None, [p])
        self.add_node = fn("add_node", C.c_int, [p, C.c_double, C.c_double, f64p, C.c_int, C.c_double, C.c_double, C.c_int])
        self.add_link = fn("add_link", C.c_int, [p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                                  C.c_double, C.c_double, C.c_double, i32p, C.c_int])
        self.add_demand = fn("add_demand", C.c_int64, [p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, i32p, C.c_int])
        self.add_nodes = fn("add_nodes", C.c_int64, [p, C.c_int64, f64p, f64p, i32p, f64p, f64p, f64p, i32p])
        self.add_links = fn("add_links", C.c_int64, [p, C.c_int64, i32p, i32p, f64p, f64p, f64p, i32p, f64p, f64p, f64p, i32p, i32p])
        self.add_demands = fn("add_demands", C.c_int64, [p, C.c_int64, i32p, i32p, f64p, f64p, f64p])
        self.set_option = fn("set_option", C.c_int, [p, C.c_int, C.c_double])
        self.add_vehicles = fn("add_vehicles", C.c_int64, [p, C.c_int64, i32p, i32p, i32p])
        self.set_vehicle_links = fn("set_vehicle_links", C.c_int, [p, C.c_int64, C.c_int, i32p, C.c_int])
        self.batch_enforce_routes_csr = fn("batch_enforce_routes_csr", C.c_int, [p, i32p, i64p, C.c_int64])
        self.set_t_max = fn("set_t_max", C.c_int, [p, C.c_double])
        self.set_toll_timeseries = fn("set_toll_timeseries", C.c_int, [p, C.c_int, f64p, C.c_int])
        self.initialize_adj_matrix = fn("initialize_adj_matrix", C.c_int, [p])
        self.main_loop = fn("main_loop", C.c_int, [p, C.c_double, C.c_double])
        self.check_simulation_ongoing = fn("check_simulation_ongoing", C.c_int, [p])
        self.set_link_param = fn("set_link_param", C.c_int, [p, C.c_int, C.c_int, C.c_double])
        self.set_link_signal_group = fn("set_link_signal_group", C.c_int, [p, C.c_int, i32p, C.c_int])
        self.set_node_signal = fn("set_node_signal", C.c_int, [p, C.c_int, f64p, C.c_int, C.c_int, C.c_double])
        self.get_int = fn("get_int", C.c_int64, [p, C.c_int])
        self.get_double = fn("get_double", C.c_double, [p, C.c_int])
        self.array_size = fn("array_size", C.c_int64, [p, C.c_int, C.POINTER(C.c_int)])
        self.get_array = fn("get_array", C.c_int64, [p, C.c_int, C.c_int, C.c_void_p, C.c_int64])
        self.get_device_array = fn("get_device_array", C.c_int64, [p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)])
        self.ensemble_link_stats_device = fn("ensemble_link_stats_device", C.c_int64, [p, C.POINTER(C.c_void_p)])
        self.ensemble_link_stats = fn("ensemble_link_stats", C.c_int64, [p, f64p, C.c_int64])
        self.last_error = fn("last_error", C.c_char_p, [p])
        # DTA solver batch operations
        self.dta_enumerate_route_sets = fn("dta_enumerate_route_sets", C.c_int64, [p, C.c_int, C.c_uint])
        self.dta_set_route_sets = fn("dta_set_route_sets", C.c_int, [p, C.c_int64, i32p, i32p, i64p, i64p, i32p])
        self.dta_share_route_sets = fn("dta_share_route_sets", C.c_int, [p, p])
        self.dta_enforce_routes = fn("dta_enforce_routes", C.c_int, [p, C.c_int, i32p, i64p, C.c_int64])
        self.dta_next_day = fn("dta_next_day", C.c_int, [p])
        self.edie_state = fn("edie_state", C.c_int, [p, C.c_int, f64p, f64p])
        self.analysis_summary = fn("analysis_summary", C.c_int, [p, C.c_int])
        self.dta_route_swap = fn("dta_route_swap", C.c_int, [p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint])

    def exported(self, name: str) -> bool:
        return hasattr(self.lib, self.prefix + name)


_product: Optional[CApi] = None


def load_product() -> CApi:
    """Load the CUDA backend.  Raises ImportError if it has not been built -- by design there is no fallback."""
    global _product
    if _product is None:
        _product = CApi(PRODUCT_LIB, "ux_")
    return _product


class EngineWorld:
    """A native world handle plus numpy-typed helpers.  Thin: one method per C-ABI entry point."""

    def __init__(self, api: CApi, *, t_max: float, delta_n: float, tau: float = 1.0, duo_update_time: float = 600.0,
                 duo_update_weight: float = 0.5, route_choice_uncertainty: float = 0.01, random_seed: int = 0,
                 vehicle_log_mode: bool = False, hard_deterministic_mode: bool = False,
                 route_choice_update_gradual: bool = False, no_cyclic_routing: bool = False,
                 adjust_node_capacity: bool = False, instantaneous_TT_timestep_interval: int = 5,
                 num_replicas: int = 1, device: int = -1, num_threads: int = 1):
        self.api = api
        delta_t = tau * delta_n
        route_ts = max(1, int(duo_update_time / delta_t))
        # uxsim.py:1806-1808: the TT interval is clamped to the route-update interval
        itt = min(int(instantaneous_TT_timestep_interval), route_ts)
        self.params = WorldParams(
            t_max=float(t_max), delta_n=float(delta_n), tau=float(tau), duo_update_time=float(duo_update_time),
            duo_update_weight=float(duo_update_weight), route_choice_uncertainty=float(route_choice_uncertainty),
            random_seed=int(random_seed), print_mode=0, vehicle_log_mode=int(bool(vehicle_log_mode)),
            hard_deterministic_mode=int(bool(hard_deterministic_mode)),
            route_choice_update_gradual=int(bool(route_choice_update_gradual)),
            no_cyclic_routing=int(bool(no_cyclic_routing)), adjust_node_capacity=int(bool(adjust_node_capacity)),
            instantaneous_TT_timestep_interval=itt, num_replicas=int(num_replicas), device=int(device),
            num_threads=int(num_threads), reserved0=0)
        self.h = api.create_world(C.byref(self.params))
        if not self.h:
            raise RuntimeError("create_world failed: " + (api.last_error(None) or b"").decode())
        self.delta_t = delta_t

    # -- error handling ---------------------------------------------------------------------------
    def _check(self, rc: int) -> int:
        if rc < 0:
            msg = (self.api.last_error(self.h) or b"").decode(errors="replace")
            raise _EXC.get(int(rc), RuntimeError)(msg)
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.api.destroy_world(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scenario ------------------------------------------------------------------------------------
    def add_node(self, x=0.0, y=0.0, signal=(0.0,), signal_offset=0.0, flow_capacity=-1.0, number_of_lanes=0) -> int:
        s = _as_f64(list(signal))
        return self._check(self.api.add_node(self.h, float(x), float(y), s.ctypes.data_as(C.POINTER(C.c_double)), len(s),
                                             float(signal_offset), float(flow_capacity), int(number_of_lanes)))

    def add_link(self, start, end, vmax, kappa, length, lanes=1, merge_priority=1.0, capacity_out=-1.0, capacity_in=-1.0,
                 signal_group=(0,)) -> int:
        g = _as_i32(list(signal_group))
        return self._check(self.api.add_link(self.h, int(start), int(end), float(vmax), float(kappa), float(length), int(lanes),
                                             float(merge_priority), float(capacity_out), float(capacity_in),
                                             g.ctypes.data_as(C.POINTER(C.c_int32)), len(g)))

    def add_demand(self, orig, dest, start_t, end_t, flow, links_preferred: Sequence[int] = ()) -> int:
        g = _as_i32(list(links_preferred))
        return self._check(self.api.add_demand(self.h, int(orig), int(dest), float(start_t), float(end_t), float(flow),
                                               g.ctypes.data_as(C.POINTER(C.c_int32)), len(g)))

    def add_nodes(self, x, y, signal_off, signal, signal_offset, flow_capacity, number_of_lanes) -> int:
        x, y, sig, so, fc = _as_f64(x), _as_f64(y), _as_f64(signal), _as_f64(signal_offset), _as_f64(flow_capacity)
        off, nl = _as_i32(signal_off), _as_i32(number_of_lanes)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        return self._check(self.api.add_nodes(self.h, len(x), x.ctypes.data_as(dp), y.ctypes.data_as(dp), off.ctypes.data_as(ip),
                                              sig.ctypes.data_as(dp), so.ctypes.data_as(dp), fc.ctypes.data_as(dp), nl.ctypes.data_as(ip)))

    def add_links(self, start, end, vmax, kappa, length, lanes, merge_priority, capacity_out, capacity_in, sg_off, sg) -> int:
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        a = [_as_i32(start), _as_i32(end), _as_f64(vmax), _as_f64(kappa), _as_f64(length), _as_i32(lanes), _as_f64(merge_priority),
             _as_f64(capacity_out), _as_f64(capacity_in), _as_i32(sg_off), _as_i32(sg)]
        ptrs = [v.ctypes.data_as(ip if v.dtype == np.int32 else dp) for v in a]
        return self._check(self.api.add_links(self.h, len(a[0]), *ptrs))

    def add_demands(self, orig, dest, t0, t1, flow) -> int:
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        o, d, a, b, f = _as_i32(orig), _as_i32(dest), _as_f64(t0), _as_f64(t1), _as_f64(flow)
        return self._check(self.api.add_demands(self.h, len(o), o.ctypes.data_as(ip), d.ctypes.data_as(ip), a.ctypes.data_as(dp),
                                                b.ctypes.data_as(dp), f.ctypes.data_as(dp)))

    def set_option(self, option: int, value: float):
        self._check(self.api.set_option(self.h, int(option), float(value)))

    def add_vehicles(self, orig, dest, depart_ts) -> int:
        o, d, t = _as_i32(orig), _as_i32(dest), _as_i32(depart_ts)
        ip = C.POINTER(C.c_int32)
        return self._check(self.api.add_vehicles(self.h, len(o), o.ctypes.data_as(ip), d.ctypes.data_as(ip), t.ctypes.data_as(ip)))

    def set_vehicle_links(self, vehicle: int, kind: int, links: Sequence[int]):
        g = _as_i32(list(links))
        self._check(self.api.set_vehicle_links(self.h, int(vehicle), int(kind), g.ctypes.data_as(C.POINTER(C.c_int32)), len(g)))

    def batch_enforce_routes_csr(self, link_ids, offsets):
        ids, off = _as_i32(link_ids), np.ascontiguousarray(offsets, dtype=np.int64)
        self._check(self.api.batch_enforce_routes_csr(self.h, ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                      off.ctypes.data_as(C.POINTER(C.c_int64)), len(off) - 1))

    def set_t_max(self, t_max: float):
        self._check(self.api.set_t_max(self.h, float(t_max)))

    def set_toll_timeseries(self, link: int, toll):
        t = _as_f64(toll)
        self._check(self.api.set_toll_timeseries(self.h, int(link), t.ctypes.data_as(C.POINTER(C.c_double)), len(t)))

    def initialize_adj_matrix(self):
        self._check(self.api.initialize_adj_matrix(self.h))

    # -- hot path -------------------------------------------------------------------------------------
    def main_loop(self, duration_t: float = -1.0, until_t: float = -1.0):
        self._check(self.api.main_loop(self.h, float(duration_t), float(until_t)))

    def check_simulation_ongoing(self) -> bool:
        return bool(self.api.check_simulation_ongoing(self.h))

    # -- DTA solver batch operations --------------------------------------------------
    def dta_enumerate_route_sets(self, k: int, seed: int) -> int:
        """World.enumerate_od_route_sets_ids_cpp (bindings.cpp:461-465): returns the number of routes."""
        return self._check(self.api.dta_enumerate_route_sets(self.h, int(k), int(seed) & 0xFFFFFFFF))

    def dta_set_route_sets(self, orig, dest, od_offsets, route_offsets, route_links):
        """World.make_od_route_set_store (bindings.cpp:467-489) from CSR arrays."""
        o, d, rl = _as_i32(orig), _as_i32(dest), _as_i32(route_links)
        oo, ro = np.ascontiguousarray(od_offsets, dtype=np.int64), np.ascontiguousarray(route_offsets, dtype=np.int64)
        ip, lp = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        self._check(self.api.dta_set_route_sets(self.h, len(o), o.ctypes.data_as(ip), d.ctypes.data_as(ip), oo.ctypes.data_as(lp),
                                                ro.ctypes.data_as(lp), rl.ctypes.data_as(ip)))

    def dta_share_route_sets(self, other: "EngineWorld"):
        """Use ``other``'s route-set store (one store per solve, as the reference's ODRouteSetStore object)."""
        self._check(self.api.dta_share_route_sets(self.h, other.h))

    def dta_enforce_routes(self, replica: int, link_ids, offsets):
        ids, off = _as_i32(link_ids), np.ascontiguousarray(offsets, dtype=np.int64)
        self._check(self.api.dta_enforce_routes(self.h, int(replica), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                off.ctypes.data_as(C.POINTER(C.c_int64)), len(off) - 1))

    def edie_state(self, edie_dt, edie_dx, replica: int = 0):
        """Analyzer.compute_edie_state (analyzer.py:249-330) for every link: list of (tn_mat, dn_mat), each ``[nt_l, nx_l]``
        (vehicle-seconds and vehicle-metres per cell); k = tn / (dt dx), q = dn / (dt dx), v = q / k."""
        a, b = _as_f64(edie_dt), _as_f64(edie_dx)
        dp = C.POINTER(C.c_double)
        self._check(self.api.edie_state(self.h, int(replica), a.ctypes.data_as(dp), b.ctypes.data_as(dp)))
        off, nx = self.get_array(A_EDIE_OFFSETS, replica), self.get_array(A_EDIE_NX, replica)
        tn, dn = self.get_array(A_EDIE_TN, replica), self.get_array(A_EDIE_DN, replica)
        out = []
        for l in range(len(nx)):
            shape = (-1, int(nx[l])) if nx[l] > 0 else (0, 0)
            out.append((tn[off[l]:off[l + 1]].reshape(shape), dn[off[l]:off[l + 1]].reshape(shape)))
        return out

    def analysis_summary(self, replica: int = 0) -> dict:
        """Analyzer.od_analysis / link_analysis_coarse (analyzer.py:121-203) as reductions inside the engine: ``od_orig, od_dest``
        (node ids of the K pairs that occur), ``od`` ``[K, 6]`` = platoons, completed platoons, mean / std travel time (-1 without
        completed trips), mean / std distance travelled; ``link`` ``[L, 4]`` = volume, vehicles remaining, mean / std of the
        positive entries of traveltime_actual (-1 without volume)."""
        self._check(self.api.analysis_summary(self.h, int(replica)))
        return dict(od_orig=self.get_array(A_OD_ORIG, replica), od_dest=self.get_array(A_OD_DEST, replica),
                    od=self.get_array(A_OD_STATS, replica).reshape(-1, 6), link=self.get_array(A_LINKC_STATS, replica).reshape(-1, 4))

    def dta_next_day(self):
        """Rewind the finished world to t = 0 with the last route swap's routes enforced (CUDA engine only)."""
        self._check(self.api.dta_next_day(self.h))

    def dta_route_sets_csr(self):
        """ODRouteSetStore.to_csr: (orig, dest, od_offsets, route_offsets, route_links), pairs sorted by (orig, dest)."""
        return tuple(self.get_array(a) for a in (A_DTA_OD_ORIG, A_DTA_OD_DEST, A_DTA_OD_OFFSETS, A_DTA_ROUTE_OFFSETS, A_DTA_ROUTE_LINKS))

    def dta_route_swap(self, mode: int, swap_prob: float, rng_seed: int, swap_num: int = -1, has_external_route_sets: bool = False,
                       replica: int = 0) -> dict:
        """World.route_swap_due (mode 0) / route_swap_dso (mode 1) (bindings.cpp:661-706): same keys as the nanobind dict."""
        self._check(self.api.dta_route_swap(self.h, int(replica), int(mode), float(swap_prob), int(swap_num),
                                            1 if has_external_route_sets else 0, int(rng_seed) & 0xFFFFFFFF))
        g = lambda a: self.get_array(a, replica)
        return {"rs_link_ids": g(A_DTA_RS_LINKS), "rs_offsets": g(A_DTA_RS_OFFSETS), "ra_link_ids": g(A_DTA_RA_LINKS),
                "ra_offsets": g(A_DTA_RA_OFFSETS), "cost_actual": g(A_DTA_COST_ACTUAL), "n_swap": self.get_double(D_DTA_N_SWAP),
                "potential_n_swap": self.get_double(D_DTA_POTENTIAL_N_SWAP), "total_t_gap": self.get_double(D_DTA_TOTAL_T_GAP)}

    def dta_route_swap_all(self, mode: int, swap_prob: float, rng_seed: int, swap_num: int = -1, has_external_route_sets: bool = False,
                           routes: bool = True) -> list:
        """Route swap of every replica in one call (replica r draws from ``rng_seed + r``): list of the per-replica dicts.
        ``routes=False`` leaves the link lists on the device (``dta_next_day`` consumes them there)."""
        self._check(self.api.dta_route_swap(self.h, -1, int(mode), float(swap_prob), int(swap_num), 1 if has_external_route_sets else 0,
                                            int(rng_seed) & 0xFFFFFFFF))
        out = []
        for r in range(self.get_int(I_NUM_REPLICAS)):
            g = lambda a: self.get_array(a, r)
            if not routes:  # scalars only: nothing per-platoon crosses the bus
                out.append({"n_swap": self.get_double(D_DTA_N_SWAP_OF + r),
                            "potential_n_swap": self.get_double(D_DTA_N_SWAP_OF + 1000 + r), "total_t_gap": self.get_double(D_DTA_N_SWAP_OF + 2000 + r)})
                continue
            out.append({"rs_link_ids": g(A_DTA_RS_LINKS), "rs_offsets": g(A_DTA_RS_OFFSETS), "ra_link_ids": g(A_DTA_RA_LINKS),
                        "ra_offsets": g(A_DTA_RA_OFFSETS), "cost_actual": g(A_DTA_COST_ACTUAL), "n_swap": self.get_double(D_DTA_N_SWAP_OF + r),
                        "potential_n_swap": self.get_double(D_DTA_N_SWAP_OF + 1000 + r), "total_t_gap": self.get_double(D_DTA_N_SWAP_OF + 2000 + r)})
        return out

    # -- interventions --------------------------------------------------------------------------------
    def set_link_param(self, link: int, field: int, value: float):
        self._check(self.api.set_link_param(self.h, int(link), int(field), float(value)))

    def set_link_signal_group(self, link: int, group: Sequence[int]):
        g = _as_i32(list(group))
        self._check(self.api.set_link_signal_group(self.h, int(link), g.ctypes.data_as(C.POINTER(C.c_int32)), len(g)))

    def set_node_signal(self, node: int, signal=None, phase: int = -1, signal_t: float = -1.0):
        if signal is None:
            self._check(self.api.set_node_signal(self.h, int(node), None, 0, int(phase), float(signal_t)))
        else:
            s = _as_f64(list(signal))
            self._check(self.api.set_node_signal(self.h, int(node), s.ctypes.data_as(C.POINTER(C.c_double)), len(s), int(phase), float(signal_t)))

    # -- results --------------------------------------------------------------------------------------
    def get_int(self, what: int) -> int:
        return int(self._check(self.api.get_int(self.h, int(what))))

    def get_double(self, what: int) -> float:
        return float(self.api.get_double(self.h, int(what)))

    def get_array_all(self, what: int, out: np.ndarray = None) -> np.ndarray:
        """A per-replica state array of EVERY replica, shape ``[R, n]``, in one device->host transfer (CUDA engine).
        ``out``: a caller-owned C-contiguous buffer of that shape and dtype to receive it -- a page-locked one (e.g.
        ``torch.empty(..., pin_memory=True).numpy()``), reused from run to run, turns the copy into a plain DMA; a fresh
        ``np.empty`` costs a page fault per 4 KB on top of the pageable copy (75 MB: 15 ms against 3)."""
        dt = C.c_int(0)
        n = self._check(self.api.array_size(self.h, int(what), C.byref(dt)))
        R = self.get_int(I_NUM_REPLICAS)
        if out is not None:
            if out.shape != (R, int(n)) or out.dtype != _NP_DTYPE[dt.value] or not out.flags.c_contiguous:
                raise ValueError(f"out must be a C-contiguous {(R, int(n))} array of {_NP_DTYPE[dt.value]}")
        else:
            out = np.empty((R, int(n)), dtype=_NP_DTYPE[dt.value])
        if out.size:
            self._check(self.api.get_array(self.h, int(what), -1, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def get_array(self, what: int, replica: int = 0) -> np.ndarray:
        dt = C.c_int(0)
        n = self._check(self.api.array_size(self.h, int(what), C.byref(dt)))
        out = np.empty(int(n), dtype=_NP_DTYPE[dt.value])
        if n > 0:
            m = self._check(self.api.get_array(self.h, int(what), int(replica), out.ctypes.data_as(C.c_void_p), int(n)))
            if m < n:  # variable-length per-replica results: array_size is an upper bound
                out = out[:m]
        return out

    def get_device_array(self, what: int, replica: int = 0):
        """Borrowed device pointer as (ptr, n_elements, numpy dtype).  CUDA backend only."""
        ptr, dt = C.c_void_p(0), C.c_int(0)
        n = self._check(self.api.get_device_array(self.h, int(what), int(replica), C.byref(ptr), C.byref(dt)))
        return ptr.value, int(n), _NP_DTYPE[dt.value]

    def ensemble_link_stats(self) -> np.ndarray:
        L = self.get_int(I_NUM_LINKS)
        out = np.empty(L * 4, dtype=np.float64)
        self._check(self.api.ensemble_link_stats(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), out.size))
        return out.reshape(L, 4)

    # convenient shaped views
    def link_series(self, what: int, replica: int = 0) -> np.ndarray:
        L, T = self.get_int(I_NUM_LINKS), self.get_int(I_TOTAL_TIMESTEPS)
        return self.get_array(what, replica).reshape(L, T)
