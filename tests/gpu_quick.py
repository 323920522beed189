"""Ad-hoc GPU timing probe (not a test): wall time of main_loop for a few workloads / replica counts."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from uxsim_b200 import _capi as C, scenario as S

api = C.load_product()
def probe(name, sc, R):
    t0 = time.perf_counter()
    E = sc.to_engine(api, vehicle_logging_timestep_interval=-1, num_replicas=R)
    E.get_array(C.A_LINK_NUM_VEHICLES)  # force upload
    t1 = time.perf_counter()
    E.main_loop()
    t2 = time.perf_counter()
    upd = E.get_int(C.I_PLATOON_UPDATES)
    arr = E.get_array(C.A_CUM_ARRIVAL); t3 = time.perf_counter()
    print(f"{name:16s} R={R:3d} build+upload {t1-t0:6.3f}s  main_loop {1e3*(t2-t1):9.2f} ms  get_cum {1e3*(t3-t2):7.2f} ms  updates(rep0) {upd}  "
          f"agg {upd*R/(t2-t1):.3e} upd/s  launches {E.get_int(C.I_KERNEL_LAUNCHES)} levels {E.get_int(C.I_TRANSFER_LEVELS)}", flush=True)
    E.close()
for name, mk in [("siouxfalls", lambda: S.sioux_falls(seed=0)), ("grid11", lambda: S.grid(seed=0)), ("chicago_dn30", lambda: S.chicago_sketch(seed=42)), ("chicago_dn5", lambda: S.chicago_sketch(deltan=5, seed=42))]:
    sc = mk()
    for R in (1, 1, 8, 64):
        if name == "chicago_dn5" and R == 64: R = 32
        probe(name, sc, R)
