import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ["UXSIM_B200_NODE"] = "rl"
import numpy as np
from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S
from parity import compare_engines, run_pair
gpu = C.load_product(); ora = C.CApi(os.path.join(ROOT, 'oracle', 'liboracle.so'), 'uxo_')
sc = S.chicago_sketch(deltan=30, tmax=10000, seed=42)
Eo = sc.to_engine(ora, vehicle_logging_timestep_interval=1); Eo.main_loop()
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    Eg = sc.to_engine(gpu, vehicle_logging_timestep_interval=1); Eg.main_loop()
    d = compare_engines(Eg, Eo)
    if d: bad += 1; print("iter", it, "DIFF", d[:6], flush=True)
    Eg.close()
print("done, bad =", bad)
