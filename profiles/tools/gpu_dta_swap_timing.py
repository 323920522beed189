"""Ad-hoc GPU probe (not a test): where a batched route swap spends its wall time."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S
api = C.load_product()
wl, R = sys.argv[1], int(sys.argv[2])
sc = S.dta_grid9() if wl == "dta_grid9" else bench.make_scenario(wl, 0)
tab = sc.tables()
E0 = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1)
E0.dta_enumerate_route_sets(4, 1)
E = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1, num_replicas=R)
E.set_option(C.OPT_RECORD_TRAVERSALS, 1); E.dta_share_route_sets(E0)
for day in range(4):
    if day: E.dta_next_day()
    E.main_loop()
    t0 = time.perf_counter()
    E._check(api.dta_route_swap(E.h, -1, 0, 0.05, -1, 0, 100 + day))
    t1 = time.perf_counter()
    for r in range(R):
        E.get_array(C.A_DTA_COST_ACTUAL, r); E.get_double(C.D_DTA_N_SWAP_OF + r)
    t2 = time.perf_counter()
    for r in range(R):
        E.get_array(C.A_VEH_TRAVEL_TIME, r)
    t3 = time.perf_counter()
    print(f"{wl} R={R} day {day}: swap call {(t1-t0)*1e3:.1f} ms (events {E.get_double(C.D_DTA_SWAP_MS):.1f}), result scalars {(t2-t1)*1e3:.1f} ms, travel times {(t3-t2)*1e3:.1f} ms", flush=True)
