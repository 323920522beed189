"""Target for compute-sanitizer (not a test): small stochastic scenarios through every kernel variant.
   compute-sanitizer --tool memcheck|racecheck python profiles/tools/gpu_sanitize_cmd.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S

api = C.load_product()
for sssp, upd, node, R, log in (("wave", "rl", "rl", 40, -1), ("wave", "rl", "hybrid", 34, -1), ("wave", "rl", "warp", 2, 1), ("frontier", "warp", "warp", 3, 1), ("sweep", "warp", "warp", 1, -1)):
    os.environ["UXSIM_B200_SSSP"], os.environ["UXSIM_B200_UPDATE"], os.environ["UXSIM_B200_NODE"] = sssp, upd, node
    for sc in (S.grid(n=5, seed=2, tmax=1500, demand_t=600, signals=True), S.sioux_falls(seed=1, tmax=1200), S.short_links_chain()):
        E = sc.to_engine(api, vehicle_logging_timestep_interval=log, num_replicas=R)
        E.main_loop()
        print(sssp, upd, node, R, sc.name, "updates", E.get_int(C.I_PLATOON_UPDATES), "trips", E.get_int(C.I_TRIPS_COMPLETED), flush=True)
        E.close()
