"""Shared parity helpers: run one Scenario through two C-ABI backends and compare every result array."""
import numpy as np

from uxsim_b200 import _capi as C

#: arrays that must be bit-identical between the CUDA engine and the oracle
EXACT_FIELDS = [
    ("cum_arrival", C.A_CUM_ARRIVAL), ("cum_departure", C.A_CUM_DEPARTURE), ("traveltime_instant", C.A_TRAVELTIME_INSTANT),
    ("traveltime_actual", C.A_TRAVELTIME_ACTUAL), ("veh_state", C.A_VEH_STATE), ("veh_arrival_ts", C.A_VEH_ARRIVAL_TS),
    ("veh_travel_time", C.A_VEH_TRAVEL_TIME), ("veh_distance", C.A_VEH_DISTANCE), ("signal_log", C.A_SIGNAL_LOG),
    ("veh_link", C.A_VEH_LINK), ("link_num_vehicles", C.A_LINK_NUM_VEHICLES), ("cap_in_remain", C.A_LINK_CAPACITY_IN_REMAIN),
    ("cap_out_remain", C.A_LINK_CAPACITY_OUT_REMAIN), ("route_pref", C.A_ROUTE_PREF), ("route_next", C.A_ROUTE_NEXT),
    ("link_stat_count", C.A_LINK_STAT_COUNT), ("link_trips_completed", C.A_LINK_TRIPS_COMPLETED),
]
RUNNING_FIELDS = [("veh_x", C.A_VEH_X), ("veh_v", C.A_VEH_V), ("veh_lane", C.A_VEH_LANE)]  # defined for RUN platoons only


def compare_engines(Ea, Eb, replica_a=0, replica_b=0, fields=None):
    """Return a list of (field, n_mismatching_elements); empty means bit-identical."""
    bad = []
    for name, what in (fields or EXACT_FIELDS):
        x, y = Ea.get_array(what, replica_a), Eb.get_array(what, replica_b)
        if x.shape != y.shape or not np.array_equal(x, y):
            bad.append((name, int((x != y).sum()) if x.shape == y.shape else -1))
    run = Eb.get_array(C.A_VEH_STATE, replica_b) == C.VS_RUN
    for name, what in RUNNING_FIELDS:
        x, y = Ea.get_array(what, replica_a)[run], Eb.get_array(what, replica_b)[run]
        if not np.array_equal(x, y):
            bad.append((name, int((x != y).sum())))
    dest = np.unique(Eb.get_array(C.A_VEH_DEST, replica_b))
    N = Eb.get_int(C.I_NUM_NODES)
    if Eb.get_int(C.I_TIMESTEP) > 0 and len(dest):
        x = Ea.get_array(C.A_ROUTE_DIST, replica_a).reshape(N, N)[:, dest]
        y = Eb.get_array(C.A_ROUTE_DIST, replica_b).reshape(N, N)[:, dest]
        if not np.array_equal(x, y):
            bad.append(("route_dist", int((x != y).sum())))
    return bad


def run_pair(sc, api_a, api_b, until=None, kw_a=None, kw_b=None, **kw):
    kw.setdefault("vehicle_logging_timestep_interval", -1)
    Ea = sc.to_engine(api_a, **{**kw, **(kw_a or {})})
    Eb = sc.to_engine(api_b, **{**kw, **(kw_b or {})})
    if until is None:
        Ea.main_loop(), Eb.main_loop()
    else:
        for u in until:
            Ea.main_loop(until_t=u), Eb.main_loop(until_t=u)
    return Ea, Eb


def summary_stats(E, replica=0):
    st = E.get_array(C.A_VEH_STATE, replica)
    tt = E.get_array(C.A_VEH_TRAVEL_TIME, replica)
    done = st == C.VS_END
    L, T = E.get_int(C.I_NUM_LINKS), E.get_int(C.I_TOTAL_TIMESTEPS)
    flow = E.get_array(C.A_CUM_ARRIVAL, replica).reshape(L, T)[:, -1]
    return int(done.sum()), float(tt[done].mean()) if done.any() else 0.0, flow


LOG_FIELDS = [(k[2:].lower(), getattr(C, k)) for k in dir(C) if k.startswith(("A_LOG_", "A_LTL_", "A_ENTER_"))]


def compare_logs(Ea, Eb, replica_a=0, replica_b=0):
    """Per-platoon logs in the reference's compact flat layout; returns the names of differing arrays."""
    bad = []
    for name, what in LOG_FIELDS:
        x, y = Ea.get_array(what, replica_a), Eb.get_array(what, replica_b)
        if x.shape != y.shape or not np.array_equal(x, y):
            bad.append(name)
    return bad
