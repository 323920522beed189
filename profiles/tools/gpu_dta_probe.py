"""Ad-hoc GPU probe (not a test): one day of the DUE / DSO day-to-day loop -- enforce routes, simulate, swap routes -- on the GPU
and in the reference C++ (oracle/_ref harness: traffi.cpp + dta_solver.cpp) on the host cores."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S

gpu = C.load_product()
ref, kind = bench.load_reference_lib()
wl = sys.argv[1] if len(sys.argv) > 1 else "chicago"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = S.dta_grid9() if wl == "dta_grid9" else bench.make_scenario(wl, 0)
tab = sc.tables()
out = {"workload": wl, "k": k}
E0 = sc.engine_from_tables(gpu, tab, vehicle_logging_timestep_interval=-1)
t0 = time.perf_counter(); out["n_routes"] = E0.dta_enumerate_route_sets(k, 123); out["enumerate_s_product_host"] = time.perf_counter() - t0
store = E0.dta_route_sets_csr()
E0.close()
out["n_od"] = len(store[0])
prev = {}
for day in range(3):
    for name, api, log in (("gpu", gpu, -1), ("ref", ref, 1)):
        E = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=log, threads=-1, random_seed=day)
        E.set_option(C.OPT_RECORD_TRAVERSALS, 1)
        t = [time.perf_counter()]
        E.dta_set_route_sets(*store); t.append(time.perf_counter())
        if name in prev: E.batch_enforce_routes_csr(*prev[name])
        t.append(time.perf_counter())
        E.main_loop(); t.append(time.perf_counter())
        for mode, label in ((0, "due"), (1, "dso")):
            a = time.perf_counter(); res = E.dta_route_swap(mode, 0.05, 1000 + day); b = time.perf_counter()
            out[f"{name}_swap_{label}_ms_day{day}"] = (b - a) * 1e3
            if name == "gpu": out[f"gpu_swap_{label}_device_ms_day{day}"] = E.get_double(C.D_DTA_SWAP_MS)
            out[f"{name}_{label}_nswap_day{day}"] = res["n_swap"]
            if mode == 0: prev[name] = (res["rs_link_ids"], res["rs_offsets"])
        out[f"{name}_set_store_ms_day{day}"] = (t[1] - t[0]) * 1e3
        out[f"{name}_enforce_ms_day{day}"] = (t[2] - t[1]) * 1e3
        out[f"{name}_sim_ms_day{day}"] = (t[3] - t[2]) * 1e3
        out[f"{name}_trips_day{day}"] = int((E.get_array(C.A_VEH_ARRIVAL_TS) >= 0).sum())
        E.close()
print(json.dumps(out, indent=1))
