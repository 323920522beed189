"""Ad-hoc GPU probe (not a test): replica-batched day-to-day DUE -- R chains per day in one world."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
from uxsim_b200 import scenario as S
from uxsim_b200.dta import DayToDayEnsemble

wl = sys.argv[1] if len(sys.argv) > 1 else "dta_grid9"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 64
days = int(sys.argv[3]) if len(sys.argv) > 3 else 6
k = int(sys.argv[4]) if len(sys.argv) > 4 else 6
sc = S.dta_grid9() if wl == "dta_grid9" else bench.make_scenario(wl, 0)
t0 = time.perf_counter()
ens = DayToDayEnsemble(sc, R, mode="DUE")
ens.solve(max_iter=days, n_routes_per_od=k, swap_prob=0.05)
tot = time.perf_counter() - t0
ds = np.array(ens.day_seconds)
print(json.dumps({"workload": wl, "chains": R, "days": days, "k": k, "total_s": tot,
                  "day_ms_build_enforce": ds[1:, 0].mean() * 1e3, "day_ms_simulation": ds[1:, 1].mean() * 1e3, "day_ms_route_swaps": ds[1:, 2].mean() * 1e3,
                  "chain_days_per_s": R / ds[1:].sum(axis=1).mean(), "swap_device_ms_last": ens.swap_device_ms[-1],
                  "mean_ttt_by_day": ens.ttts.mean(axis=0).tolist(), "mean_gap_by_day": ens.t_gaps.mean(axis=0).tolist(),
                  "trips_last_day_mean": ens.trips[:, -1].mean()}))
