"""Ad-hoc GPU probe (not a test): dump the instantaneous link travel times of the C4 grid at its route-update steps,
so that shortest-path kernels can be prototyped on realistic edge costs on the CPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
from uxsim_b200 import _capi as C

wl = sys.argv[1] if len(sys.argv) > 1 else "grid100"
sc = bench.make_scenario(wl, 0)
tab = sc.tables()
E = sc.engine_from_tables(C.load_product(), tab, vehicle_logging_timestep_interval=-1, num_replicas=1, random_seed=0)
E.main_loop()
rts = E.get_int(C.I_ROUTE_UPDATE_STEPS)
L, T = E.get_int(C.I_NUM_LINKS), E.get_int(C.I_TOTAL_TIMESTEPS)
tt = E.get_array(C.A_TRAVELTIME_INSTANT, 0).reshape(L, T)
print("loop ms", E.get_double(C.D_LAST_LOOP_MS), "route steps", rts, "trips", E.get_int(C.I_TRIPS_COMPLETED), "of", E.get_int(C.I_NUM_VEHICLES))
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed(f"gpurun_out/costs_{wl}.npz", tt=tt[:, ::rts].astype(np.float64), start=np.asarray(tab["l_start"]), end=np.asarray(tab["l_end"]))
