"""Profiling target (not a test): one main_loop of the bench workload, nothing else.  Used under ncu:
   python profiles/tools/gpu_profile_cmd.py [workload] [replicas] [until_t]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from uxsim_b200 import _capi as C

wl = sys.argv[1] if len(sys.argv) > 1 else "chicago"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
until = float(sys.argv[3]) if len(sys.argv) > 3 else -1.0
sc = bench.make_scenario(wl, 0)
E = sc.engine_from_tables(C.load_product(), sc.tables(), vehicle_logging_timestep_interval=-1, num_replicas=R, random_seed=0)
E.main_loop(until_t=until)
print("loop ms", E.get_double(C.D_LAST_LOOP_MS), "steps", E.get_int(C.I_TIMESTEP), "updates(rep0)", E.get_int(C.I_PLATOON_UPDATES), "launches", E.get_int(C.I_KERNEL_LAUNCHES))
