"""Summary statistics of a finished run, computed from the arrays the engine hands back.

Used as ``W.analyzer`` when the reference package (and with it ``uxsim/analyzer.py``) is not installed -- e.g. on the
GPU box.  It restates the numbers of ``Analyzer.basic_analysis`` / ``od_analysis`` / ``print_simple_stats``
(analyzer.py:98-181, 1166-1190) with vectorised numpy instead of per-vehicle Python loops; plotting, pandas
export and Edie states stay with the reference's analyzer.
"""
from __future__ import annotations

import numpy as np

from . import _capi as C


_FF_CACHE = {}


def free_flow_times(W) -> np.ndarray:
    """All-pairs free-flow travel time (analyzer.py:147-163 runs scipy's dense floyd_warshall on length/u).  Same
    distances from a sparse all-sources Dijkstra (O(N E log N) instead of O(N^3): 0.6 s -> 30 ms on the Chicago sketch),
    cached per network so that the worlds of a day-to-day solve compute it once."""
    n = len(W.NODES)
    key = (n, tuple((l.start_node.id, l.end_node.id, float(l.length), float(l.u), l.capacity_in == 0) for l in W.LINKS))
    hit = _FF_CACHE.get(key)
    if hit is not None:
        return hit
    best = {}
    for l in W.LINKS:  # analyzer.py:146-157: the matrix cell is overwritten link by link, so the LAST link of a node pair decides
        k = (l.start_node.id, l.end_node.id)
        best[k] = np.inf if l.capacity_in == 0 else l.length / l.u
    best = {k: v for k, v in best.items() if np.isfinite(v)}
    try:
        from scipy.sparse import csr_matrix
        from scipy.sparse.csgraph import dijkstra

        if best:
            ij = np.array(list(best.keys()))
            m = csr_matrix((np.array(list(best.values())), (ij[:, 0], ij[:, 1])), shape=(n, n))
        else:
            m = csr_matrix((n, n))
        d = dijkstra(m, directed=True)
    except ImportError:
        d = np.full((n, n), np.inf)
        for (i, j), v in best.items():
            d[i, j] = v
        np.fill_diagonal(d, 0.0)
        for k in range(n):
            d = np.minimum(d, d[:, k:k + 1] + d[k:k + 1, :])
    if len(_FF_CACHE) > 8:
        _FF_CACHE.clear()
    _FF_CACHE[key] = d
    return d


class BasicAnalyzer:
    def __init__(self, W):
        self.W = W
        self.average_speed = 0.0
        self.average_speed_count = 0
        self.trip_all = self.trip_completed = 0
        self.total_travel_time = self.average_travel_time = self.total_delay = self.average_delay = -1
        self.total_distance_traveled = 0.0

    def basic_analysis(self):
        W = self.W
        dn = W.DELTAN
        st, tt = W._veh("state"), W._veh("travel_time")
        dist = W._veh("distance")
        o = W._E.get_array(C.A_VEH_ORIG)
        d = W._E.get_array(C.A_VEH_DEST)
        done = tt >= 0
        self.trip_all = len(st) * dn
        self.trip_completed = int(done.sum()) * dn
        self.total_distance_traveled = float(dist.sum()) * dn
        if done.any():
            ff = free_flow_times(W)
            self.total_travel_time = float(tt[done].sum()) * dn
            self.average_travel_time = self.total_travel_time / self.trip_completed
            self.total_delay = float((tt[done] - ff[o[done], d[done]]).sum()) * dn
            self.average_delay = self.total_delay / self.trip_completed
        else:
            self.total_travel_time = self.average_travel_time = self.total_delay = self.average_delay = -1
        upd = W._E.get_array(C.A_LINK_STAT_COUNT).sum()
        self.average_speed_count = int(upd)

    def print_simple_stats(self, force_print=False):
        W = self.W
        p = print if force_print else W.print
        p("results:")
        p(f" number of completed trips:\t {self.trip_completed} / {self.trip_all}")
        if self.trip_completed > 0:
            p(f" total travel time:\t\t {self.total_travel_time:.1f} s")
            p(f" average travel time of trips:\t {self.average_travel_time:.1f} s")
            p(f" average delay of trips:\t {self.average_delay:.1f} s")
            p(f" delay ratio:\t\t\t {self.average_delay / self.average_travel_time:.3f}")
            p(f" total distance traveled:\t {self.total_distance_traveled:.1f} m")

    def compute_edie_state(self):
        """Edie's traffic state per link (analyzer.py:249-330): fills ``tn_mat, dn_mat, k_mat, q_mat, v_mat`` of every link
        from the device (one pass over the RUN records, ``ux_edie_state``) instead of Python loops over every log entry."""
        W = self.W
        if getattr(self, "flag_edie_state_computed", 0):
            return 0
        self.flag_edie_state_computed = 1
        res = W._E.edie_state([l.edie_dt for l in W.LINKS], [l.edie_dx for l in W.LINKS])
        for l, (tn, dn) in zip(W.LINKS, res):
            nt, nx = tn.shape
            l.tn_mat[:nt, :nx], l.dn_mat[:nt, :nx] = tn, dn
            l.k_mat, l.q_mat = tn / l.edie_dt / l.edie_dx, dn / l.edie_dt / l.edie_dx
            with np.errstate(invalid="ignore", divide="ignore"):
                l.v_mat = np.nan_to_num(l.q_mat / l.k_mat, nan=l.u)

    def od_analysis(self):
        """OD-specific stats (analyzer.py:121-181) as a reduction inside the engine (``ux_analysis_summary``: one thread per OD pair
        over the per-platoon results): dict of arrays over the OD pairs that occur, keys ``orig, dest`` (node ids),
        ``total_trips, completed_trips, free_travel_time, average_travel_time, stddiv_travel_time,
        average_distance_traveled_per_veh, stddiv_distance_traveled_per_veh`` -- the columns of ``od_to_pandas``.  As in the
        reference, a pair without a completed trip reports 0 for the averages (its lists are never evaluated there)."""
        W = self.W
        dn = W.DELTAN
        res = W._E.analysis_summary()
        o, d, od = res["od_orig"].astype(np.int64), res["od_dest"].astype(np.int64), res["od"]
        has = od[:, 1] > 0
        ff = free_flow_times(W)
        z = np.zeros(len(o))
        return dict(orig=o, dest=d, total_trips=od[:, 0] * dn, completed_trips=od[:, 1] * dn, free_travel_time=np.where(has, ff[o, d], 0.0),
                    average_travel_time=np.where(has, od[:, 2], z), stddiv_travel_time=np.where(has, od[:, 3], z),
                    average_distance_traveled_per_veh=np.where(has, od[:, 4], z), stddiv_distance_traveled_per_veh=np.where(has, od[:, 5], z))

    def link_summary(self):
        """traffic_volume / vehicles_remain / free, average and std travel time per link (link_analysis_coarse, analyzer.py:183-203),
        reduced inside the engine (one thread per link over the travel-time event table)."""
        W = self.W
        lk = W._E.analysis_summary()["link"]
        return dict(traffic_volume=lk[:, 0], vehicles_remain=lk[:, 1], free_travel_time=np.array([l.length / l.u for l in W.LINKS]),
                    average_travel_time=lk[:, 2], stddiv_travel_time=lk[:, 3])
