"""Ad-hoc GPU probe (not a test): 32-seed means of the CUDA engine vs the reference C++ (oracle/_ref) on a workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import bench
from uxsim_b200 import _capi as C
from parity import summary_stats
gpu = C.load_product(); ref, kind = bench.load_reference_lib()
wl = sys.argv[1]; R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sc = bench.make_scenario(wl, 0); tab = sc.tables()
Eg = sc.engine_from_tables(gpu, tab, vehicle_logging_timestep_interval=-1, num_replicas=R, random_seed=1000)
Eg.main_loop()
g = [summary_stats(Eg, r) for r in range(R)]
rr = []
for s in range(R):
    Er = sc.engine_from_tables(ref, tab, vehicle_logging_timestep_interval=-1, random_seed=2000 + s, threads=-1)
    Er.main_loop(); rr.append(summary_stats(Er)); Er.close()
gt, ga, gf = np.mean([x[0] for x in g]), np.mean([x[1] for x in g]), np.mean([x[2] for x in g], axis=0)
rt, ra, rf = np.mean([x[0] for x in rr]), np.mean([x[1] for x in rr]), np.mean([x[2] for x in rr], axis=0)
print(f"{wl} R={R}: trips {gt:.1f} vs {rt:.1f} ({gt/rt-1:+.4f}); avg travel time {ga:.2f} vs {ra:.2f} ({ga/ra-1:+.4f}); link-flow L1 {np.abs(gf-rf).sum()/rf.sum():.4f}; "
      f"std trips gpu {np.std([x[0] for x in g]):.1f} ref {np.std([x[0] for x in rr]):.1f}")
