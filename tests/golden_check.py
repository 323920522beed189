"""Compare an engine run against a committed golden .npz (see tests/golden/make_golden.py)."""
import hashlib
import os

import numpy as np

from uxsim_b200 import _capi as C

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def check_against_golden(E, kind, name, replica=0):
    g = np.load(os.path.join(GOLDEN_DIR, f"golden_{kind}_{name}.npz"))
    dn = float(g["deltan"])
    L, T = E.get_int(C.I_NUM_LINKS), E.get_int(C.I_TOTAL_TIMESTEPS)
    arr = E.get_array(C.A_CUM_ARRIVAL, replica).reshape(L, T)
    dep = E.get_array(C.A_CUM_DEPARTURE, replica).reshape(L, T)
    ai, di = np.rint(arr / dn).astype(np.int32), np.rint(dep / dn).astype(np.int32)
    assert np.array_equal(ai * dn, arr) and np.array_equal(di * dn, dep), "cumulative counts are not multiples of deltan"
    assert np.array_equal(ai[:, ::8], g["arr_sub"]), "cum_arrival differs from the reference"
    assert np.array_equal(di[:, ::8], g["dep_sub"]), "cum_departure differs from the reference"
    assert np.array_equal(ai[:, -1], g["arr_last"]) and np.array_equal(di[:, -1], g["dep_last"])
    assert sha(ai) == str(g["arr_sha"]) and sha(di) == str(g["dep_sha"]), "full cumulative curves differ (hash)"
    assert np.array_equal(E.get_array(C.A_VEH_ARRIVAL_TS, replica), g["arrival_ts"]), "arrival timesteps differ"
    if "travel_time" in g.files:  # python engine: (arrival - departure step) * dt
        tt = E.get_array(C.A_VEH_TRAVEL_TIME, replica)
        assert np.array_equal(tt, g["travel_time"])
    if "tt_actual_sha" in g.files:
        assert sha(E.get_array(C.A_TRAVELTIME_ACTUAL, replica)) == str(g["tt_actual_sha"]), "traveltime_actual differs"
        assert sha(E.get_array(C.A_VEH_TRAVEL_TIME, replica)) == str(g["travel_time_sha"])
        assert E.get_int(C.I_TRIPS_COMPLETED) == int(g["trips_completed"])
