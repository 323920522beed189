"""Ad-hoc GPU probe (not a test): where the wall time of one end-to-end ensemble run goes (host stages vs device loop)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from uxsim_b200 import _capi as C

api = C.load_product()
wl = sys.argv[1] if len(sys.argv) > 1 else "chicago"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sc = bench.make_scenario(wl, 0)
tab = sc.tables()
for it in range(4):
    t = [time.perf_counter()]
    E = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1, num_replicas=R, random_seed=it); t.append(time.perf_counter())
    E.get_array(C.A_LINK_NUM_VEHICLES); t.append(time.perf_counter())
    E.main_loop(); t.append(time.perf_counter())
    st = E.ensemble_link_stats(); t.append(time.perf_counter())
    res = (E.get_array_all(C.A_VEH_ARRIVAL_TS), E.get_array_all(C.A_VEH_TRAVEL_TIME), E.get_array_all(C.A_LINK_STAT_COUNT)); t.append(time.perf_counter())
    curves = (E.get_array(C.A_CUM_ARRIVAL, 0), E.get_array(C.A_CUM_DEPARTURE, 0)); t.append(time.perf_counter())
    eng = "upload %.1f graph-build %.1f loop-wall %.1f loop-device %.1f" % tuple(E.get_double(k) for k in (C.D_UPLOAD_MS, C.D_GRAPH_BUILD_MS, C.D_LAST_LOOP_WALL_MS, C.D_LAST_LOOP_MS))
    E.close(); t.append(time.perf_counter())
    print(f"{wl} R={R} total {(t[-1]-t[0])*1e3:.1f} ms | ingest, upload, main_loop, link stats, readback, curves, close = " + " ".join(f"{(b-a)*1e3:.1f}" for a, b in zip(t, t[1:])) + " | engine: " + eng, flush=True)
