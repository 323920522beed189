"""Ad-hoc probe: one workload, given replica counts."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, bench
from uxsim_b200 import _capi as C
api = C.load_product()
wl = sys.argv[1]; Rs = [int(x) for x in sys.argv[2:]] or [1]
sc = bench.make_scenario(wl, 0); tab = sc.tables()
for R in Rs:
    E = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1, num_replicas=R)
    E.get_array(C.A_LINK_NUM_VEHICLES); E.set_option(C.OPT_KERNEL_TIMING, 1); t0 = time.time(); E.main_loop(); wall = time.time() - t0
    upd = sum(int(E.get_array(C.A_LINK_STAT_COUNT, r).sum()) for r in range(R))
    st = E.get_array(C.A_VEH_STATE)
    ks = " ".join(f"{C.K_NAMES[k]}={E.get_double(C.D_KERNEL_MS_OF+k)/max(1,E.get_int(C.I_KERNEL_LAUNCHES_OF+k))*1e3:.1f}us x{E.get_int(C.I_KERNEL_LAUNCHES_OF+k)}" for k in range(6))
    print(f"{wl} R={R} eager loop {E.get_double(C.D_LAST_LOOP_MS):.1f} ms wall {wall:.2f}s updates {upd:.3e} upd/s {upd/E.get_double(C.D_LAST_LOOP_MS)*1e3:.3e} | states {np.bincount(st, minlength=5)} | {ks}", flush=True)
    E.close()
