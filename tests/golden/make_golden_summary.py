"""Golden OD / link summaries from the REFERENCE's own analyzer (dev container only; needs /root/reference):

    python tests/golden/make_golden_summary.py

Runs the reference's pure-Python engine on deterministic scenarios, then Analyzer.od_analysis and
Analyzer.link_analysis_coarse (analyzer.py:121-203), and stores what od_to_pandas / link_to_pandas would print:
per OD pair (node ids) trips, completed trips, mean / std travel time, mean / std distance travelled; per link
volume, remaining vehicles, mean / std of the positive entries of traveltime_actual."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

from uxsim_b200 import scenario as S  # noqa: E402

SUMMARY_CASES = [
    ("3link_spillback", S.straight_3link_spillback),
    ("2link_signal", S.straight_2link_signal),
    ("merge_unequal", lambda: S.merge_y(priority1=1, priority2=2)),
    ("diverge", S.diverge_2route),
    ("dead_end", S.dead_end),
]


def run(sc):
    from ref_python_shim import import_reference

    ux = import_reference()
    W = ux.World(**{**sc.world, "vehicle_logging_timestep_interval": -1})
    S.replay_into(sc, W)
    W.exec_simulation()
    A = W.analyzer
    A.od_analysis()
    A.link_analysis_coarse()
    pairs = sorted(A.od_trips.keys(), key=lambda od: od[0].id * len(W.NODES) + od[1].id)
    od = np.array([[A.od_trips[p], A.od_trips_comp[p], A.od_tt_ave[p], A.od_tt_std[p], A.od_dist_ave[p], A.od_dist_std[p]] for p in pairs], dtype=np.float64).reshape(-1, 6)
    link = np.array([[A.linkc_volume[l], A.linkc_remain[l], A.linkc_tt_ave[l], A.linkc_tt_std[l]] for l in W.LINKS], dtype=np.float64).reshape(-1, 4)
    return dict(od_orig=np.array([p[0].id for p in pairs], dtype=np.int32), od_dest=np.array([p[1].id for p in pairs], dtype=np.int32), od=od, link=link,
                deltan=np.float64(W.DELTAN))


if __name__ == "__main__":
    for name, mk in SUMMARY_CASES:
        out = run(mk())
        np.savez_compressed(os.path.join(HERE, f"golden_summary_{name}.npz"), **out)
        print("summary", name, out["od"].shape, out["link"].shape, "completed", out["od"][:, 1].sum())
