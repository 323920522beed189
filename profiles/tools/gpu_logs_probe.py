"""Ad-hoc GPU probe (not a test): cost of the per-platoon vehicle logs (API default) on the GPU and in the reference C++."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from uxsim_b200 import _capi as C

api = C.load_product()
ref, kind = bench.load_reference_lib()
wl = sys.argv[1] if len(sys.argv) > 1 else "chicago"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sc = bench.make_scenario(wl, 0)
tab = sc.tables()
for log in (-1, 1):
    for it in range(2):
        E = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=log, num_replicas=R, random_seed=it)
        E.get_array(C.A_LINK_NUM_VEHICLES)
        t0 = time.perf_counter(); E.main_loop(); t1 = time.perf_counter()
        upd = sum(int(E.get_array(C.A_LINK_STAT_COUNT, r).sum()) for r in range(R))
        ms = E.get_double(C.D_LAST_LOOP_MS)
        tl = 0.0
        if log == 1:
            t2 = time.perf_counter(); n = len(E.get_array(C.A_LOG_T, 0)); tl = (time.perf_counter() - t2) * 1e3
        E.close()
    print(f"{wl} R={R} logs={'on' if log == 1 else 'off'}: device loop {ms:.1f} ms, {upd/ms*1e3:.3e} upd/s; log hand-back of one replica {tl:.1f} ms", flush=True)
    E = sc.engine_from_tables(ref, tab, vehicle_logging_timestep_interval=log, threads=0)
    t0 = time.perf_counter(); E.main_loop(); dt = time.perf_counter() - t0
    u = E.get_int(C.I_PLATOON_UPDATES)
    t2 = time.perf_counter(); n = len(E.get_array(C.A_LOG_T, 0)) if log == 1 else 0; tl = (time.perf_counter() - t2) * 1e3
    E.close()
    print(f"   reference C++ ({kind}, all cores) logs={'on' if log == 1 else 'off'}: {dt*1e3:.1f} ms, {u/dt:.3e} upd/s; log hand-back {tl:.1f} ms", flush=True)
