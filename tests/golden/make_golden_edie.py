"""Golden Edie states from the REFERENCE's own analyzer (dev container only; needs /root/reference):

    python tests/golden/make_golden_edie.py

Runs the reference's pure-Python engine with per-step logs on deterministic scenarios, then
Analyzer.compute_edie_state (analyzer.py:249-330), and stores per link tn_mat / dn_mat with the cell sizes used."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

from uxsim_b200 import scenario as S  # noqa: E402

EDIE_CASES = [
    ("3link_spillback", S.straight_3link_spillback),
    ("2link_signal", S.straight_2link_signal),
    ("lane_drop", S.lane_drop_2to1),
    ("merge_unequal", lambda: S.merge_y(priority1=1, priority2=2)),
]


def run(sc):
    from ref_python_shim import import_reference

    ux = import_reference()
    W = ux.World(**{**sc.world, "vehicle_logging_timestep_interval": 1})
    S.replay_into(sc, W)
    W.exec_simulation()
    W.analyzer.compute_edie_state()
    out = dict(edie_dt=np.array([l.edie_dt for l in W.LINKS], dtype=np.float64), edie_dx=np.array([l.edie_dx for l in W.LINKS], dtype=np.float64))
    for i, l in enumerate(W.LINKS):
        nt, nx = int(W.TMAX / l.edie_dt), int(l.length / l.edie_dx)
        out[f"tn_{i}"] = np.asarray(l.tn_mat, dtype=np.float64)[:nt, :nx]
        out[f"dn_{i}"] = np.asarray(l.dn_mat, dtype=np.float64)[:nt, :nx]
    return out


if __name__ == "__main__":
    for name, mk in EDIE_CASES:
        np.savez_compressed(os.path.join(HERE, f"golden_edie_{name}.npz"), **run(mk()))
        print("edie", name)
