"""Generate the committed golden vectors from the REFERENCE ITSELF (dev container only; needs /root/reference):

    python tests/golden/make_golden.py

  golden_python_<case>.npz : reference pure-Python engine (import shim oracle/ref_python_shim.py)
  golden_philox_<case>.npz : reference C++ engine with Philox draws (philox_patch_reference.py)

Stored per case (platoon counts as int32 = curve / deltan, exact): cum_arrival / cum_departure at every 8th step
and at the last step, SHA-256 of the full int32 arrays, per-platoon arrival step, and for the C++ cases the SHA-256
of the FP64 travel-time series as well.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

from cases import PHILOX_CASES, PYTHON_CASES  # noqa: E402
from uxsim_b200 import _capi as C  # noqa: E402
from uxsim_b200 import scenario as S  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pack(dn, arr, dep, arrival_ts, extra=None):
    ai, di = np.rint(arr / dn).astype(np.int32), np.rint(dep / dn).astype(np.int32)
    assert np.array_equal(ai * dn, arr) and np.array_equal(di * dn, dep)
    out = dict(deltan=np.float64(dn), arr_sub=ai[:, ::8], dep_sub=di[:, ::8], arr_last=ai[:, -1], dep_last=di[:, -1],
               arr_sha=np.array(sha(ai)), dep_sha=np.array(sha(di)), arrival_ts=arrival_ts.astype(np.int32))
    out.update(extra or {})
    return out


def run_python(sc):
    from ref_python_shim import import_reference

    ux = import_reference()
    W = ux.World(**{**sc.world, "vehicle_logging_timestep_interval": -1})
    S.replay_into(sc, W)
    W.exec_simulation()
    arr = np.array([l.cum_arrival for l in W.LINKS], dtype=np.float64)
    dep = np.array([l.cum_departure for l in W.LINKS], dtype=np.float64)
    ats = np.array([v.arrival_time if v.state == "end" else -1 for v in W.VEHICLES.values()], dtype=np.int32)
    tt = np.array([v.travel_time if v.state == "end" else -1.0 for v in W.VEHICLES.values()], dtype=np.float64)
    return pack(sc.deltan, arr, dep, ats, dict(travel_time=tt))


def run_philox(sc, api):
    E = sc.to_engine(api, vehicle_logging_timestep_interval=-1)
    E.main_loop()
    L, T = E.get_int(C.I_NUM_LINKS), E.get_int(C.I_TOTAL_TIMESTEPS)
    arr, dep = E.get_array(C.A_CUM_ARRIVAL).reshape(L, T), E.get_array(C.A_CUM_DEPARTURE).reshape(L, T)
    extra = dict(tt_actual_sha=np.array(sha(E.get_array(C.A_TRAVELTIME_ACTUAL))), travel_time_sha=np.array(sha(E.get_array(C.A_VEH_TRAVEL_TIME))),
                 platoon_updates=np.int64(E.get_int(C.I_PLATOON_UPDATES)), trips_completed=np.int64(E.get_int(C.I_TRIPS_COMPLETED)))
    return pack(sc.deltan, arr, dep, E.get_array(C.A_VEH_ARRIVAL_TS), extra)


if __name__ == "__main__":
    for name, mk in PYTHON_CASES:
        np.savez_compressed(os.path.join(HERE, f"golden_python_{name}.npz"), **run_python(mk()))
        print("python", name)
    import philox_patch_reference

    lib = philox_patch_reference.build()
    api = C.CApi(lib, "uxr_")
    for name, mk in PHILOX_CASES:
        np.savez_compressed(os.path.join(HERE, f"golden_philox_{name}.npz"), **run_philox(mk(), api))
        print("philox", name)
