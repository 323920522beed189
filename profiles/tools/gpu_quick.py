"""Ad-hoc GPU probe (not a test): per-kernel event timing for a workload at several replica counts."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
from uxsim_b200 import _capi as C

api = C.CApi(os.environ["UX_LIB"], "ux_") if os.environ.get("UX_LIB") else C.load_product()  # UX_LIB: probe a variant build
def probe(wl, R):
    sc = bench.make_scenario(wl, 0); tab = sc.tables()
    E = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=int(os.environ.get("UX_LOG", "-1")), num_replicas=R)
    E.get_array(C.A_LINK_NUM_VEHICLES); E.main_loop(); ms = E.get_double(C.D_LAST_LOOP_MS)
    upd = sum(int(E.get_array(C.A_LINK_STAT_COUNT, r).sum()) for r in range(R)); E.close()
    E = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=int(os.environ.get("UX_LOG", "-1")), num_replicas=R)
    E.set_option(C.OPT_KERNEL_TIMING, 1); E.main_loop()
    ks = " ".join(f"{C.K_NAMES[k]}={E.get_double(C.D_KERNEL_MS_OF+k)/max(1,E.get_int(C.I_KERNEL_LAUNCHES_OF+k))*1e3:.1f}us x{E.get_int(C.I_KERNEL_LAUNCHES_OF+k)}" for k in range(6)).replace("k_link_pre", "k_pre").replace("k_edge_cost", "k_ec")
    print(f"{wl:12s} R={R:3d} graph loop {ms:8.2f} ms  {upd/ms*1e3:.3e} upd/s | eager-timed: {ks}", flush=True)
    E.close()
RS = [int(r) for r in os.environ.get("UX_RS", "1,8,32").split(",")]
for wl in sys.argv[1:] or ["chicago", "chicago_dn5", "grid11"]:
    for R in RS:
        probe(wl, R)
