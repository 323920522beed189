"""Ad-hoc GPU probe (not a test): do independent replica groups on separate streams overlap each other's kernel tails?"""
import sys, time, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from uxsim_b200 import _capi as C

api = C.load_product()
wl = sys.argv[1] if len(sys.argv) > 1 else "chicago"
Rtot = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sc = bench.make_scenario(wl, 0)
tab = sc.tables()
for G in [int(x) for x in os.environ.get("UX_GROUPS", "1,2,4,8").split(",")]:
    R = Rtot // G
    for it in range(2):
        Es = [sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1, num_replicas=R, random_seed=1000 * it + g * R) for g in range(G)]
        for E in Es:
            E.get_array(C.A_LINK_NUM_VEHICLES)
        th = [threading.Thread(target=E.main_loop) for E in Es]
        t0 = time.perf_counter()
        for t in th:
            t.start(); time.sleep(float(os.environ.get('UX_DELAY_MS', '0')) / 1e3)
        for t in th: t.join()
        dt = (time.perf_counter() - t0) * 1e3
        upd = sum(int(E.get_array(C.A_LINK_STAT_COUNT, r).sum()) for E in Es for r in range(R))
        dev = [E.get_double(C.D_LAST_LOOP_MS) for E in Es]
        for E in Es: E.close()
    print(f"{wl} total R={Rtot} as {G} worlds x {R}: wall {dt:.1f} ms  {upd/dt*1e3:.3e} upd/s  (device per world: {' '.join('%.1f' % d for d in dev)})", flush=True)
