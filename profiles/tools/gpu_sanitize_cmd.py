"""Target for compute-sanitizer (not a test): small stochastic scenarios through every kernel variant.
   compute-sanitizer --tool memcheck|racecheck python profiles/tools/gpu_sanitize_cmd.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S

api = C.load_product()
for sssp, upd, R, log in (("wave", "rl", 40, -1), ("wave", "rl", 2, 1), ("frontier", "warp", 3, 1), ("sweep", "warp", 1, -1)):
    os.environ["UXSIM_B200_SSSP"], os.environ["UXSIM_B200_UPDATE"] = sssp, upd
    for sc in (S.grid(n=5, seed=2, tmax=1500, demand_t=600, signals=True), S.sioux_falls(seed=1, tmax=1200)):
        E = sc.to_engine(api, vehicle_logging_timestep_interval=log, num_replicas=R)
        E.main_loop()
        print(sssp, upd, R, sc.name, "updates", E.get_int(C.I_PLATOON_UPDATES), "trips", E.get_int(C.I_TRIPS_COMPLETED), flush=True)
        E.close()
