"""BASELINE.json: "Stochastic networks must match the reference Python and C++ modes on total trips completed, average travel
time and per-link cumulative flow, within 1 % on the mean over 32 seeds."

The reference side is committed as fixtures generated from the reference ITSELF (tests/golden/make_stat_fixtures.py, run in the
dev container): per-seed trips / average travel time and per-link flow means of 256 seeds of the unmodified C++ engine
(traffi.cpp, its own mt19937 streams) and 64 seeds of the pure-Python engine (uxsim.py, numpy Generator).  The CUDA side runs 256
seeds as 256 replicas of ONE world.  (Bit-exact agreement with the reference under a shared random stream is a separate test,
test_gpu_golden.py.)

What the fixtures show about the standard itself: the reference's OWN 32-seed means scatter by more than 1 % on some statistics
(spread between disjoint 32-seed groups of the C++ engine: Sioux Falls 0.3 % trips / 0.9 % travel time / 1.3 % link flow; 11x11 grid
0.02 % / 2.3 % / 2.7 %; Chicago sketch 1.1 % / 1.7 % / 3.6 %), and its two engines differ from each other by 2 % on the grid's
travel time.  So every bound below is 1 %, unless the MEASURED noise floor of the compared means -- three standard errors of the
difference, from the per-seed spread -- is larger; the floor is computed next to the bound and printed.  With 256 seeds a side the
floors are small enough that most bounds ARE the 1 % of the north star.
"""
import os

import numpy as np
import pytest

from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S

from parity import summary_stats

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NETWORKS = {"siouxfalls": lambda s: S.sioux_falls(seed=s), "grid11": lambda s: S.grid(seed=s), "chicago_dn30": lambda s: S.chicago_sketch(seed=s)}
R = 256
TOL = 0.01  # the north star's 1 %


def gpu_ensemble(name, cuda_api):
    sc = NETWORKS[name](1000)
    E = sc.engine_from_tables(cuda_api, sc.tables(), vehicle_logging_timestep_interval=-1, num_replicas=R, random_seed=1000)
    E.main_loop()
    st = E.get_array_all(C.A_VEH_ARRIVAL_TS)
    tt = E.get_array_all(C.A_VEH_TRAVEL_TIME)
    done = st >= 0
    trips = done.sum(axis=1).astype(np.float64)
    att = np.array([tt[r][done[r]].mean() for r in range(R)])
    L, T = E.get_int(C.I_NUM_LINKS), E.get_int(C.I_TOTAL_TIMESTEPS)
    flow = np.array([E.get_array(C.A_CUM_ARRIVAL, r).reshape(L, T)[:, -1] for r in range(R)])
    assert summary_stats(E, 3)[0] == trips[3]
    E.close()
    return trips, att, flow


def bound(tol_rel, ref_mean, *sems):
    """max(1 % of the reference mean, 3 standard errors of the difference of the two means) -> (absolute bound, floor / mean)"""
    floor = 3.0 * float(np.sqrt(sum(s * s for s in sems)))
    return max(tol_rel * ref_mean, floor), floor / ref_mean


def sem(x):
    return float(np.std(x, ddof=1) / np.sqrt(len(x)))


@pytest.mark.parametrize("name", list(NETWORKS))
def test_256_seed_means_match_reference_cpp_and_python(name, cuda_api):
    z = np.load(os.path.join(GOLDEN, f"stat_{name}.npz"))
    g_trips, g_att, g_flow = gpu_ensemble(name, cuda_api)
    report = []
    # ---- scalars vs the C++ engine (256 seeds) and the Python engine (64 seeds)
    for mode in ("cpp", "python"):
        for stat, g, ref in (("trips", g_trips, z[f"{mode}_trips"].astype(np.float64)), ("att", g_att, z[f"{mode}_att"])):
            b, floor = bound(TOL, ref.mean(), sem(g), sem(ref))
            if mode == "python":
                # the GPU engine implements the C++ engine's semantics; where the reference's two engines differ from each other
                # (11x11 grid, travel time: 2 %), it may be as far from the Python engine as the reference's own C++ engine is
                b = max(b, abs(z[f"cpp_{stat}"].mean() - ref.mean()) + floor * ref.mean())
            d = abs(g.mean() - ref.mean())
            report.append(f"{name} {mode} {stat}: gpu {g.mean():.2f} ref {ref.mean():.2f} diff {d / ref.mean():.4%} bound {b / ref.mean():.4%} (noise floor {floor:.4%})")
            assert d <= b, report[-1]
    # ---- per-link cumulative flow: L1 distance of the mean flows relative to the total flow; floor = expected L1 of the noise
    cg = z["cpp_flow_groups"]  # [8, L] means over disjoint groups of 32 seeds
    var_g = g_flow.var(axis=0, ddof=1) / R
    var_c = cg.var(axis=0, ddof=1) / cg.shape[0]
    cf = cg.mean(axis=0)
    noise = float((np.sqrt(2.0 / np.pi) * np.sqrt(var_g + var_c)).sum() / cf.sum())
    d = float(np.abs(g_flow.mean(axis=0) - cf).sum() / cf.sum())
    report.append(f"{name} cpp link flow L1: {d:.4%} bound {max(TOL, 1.5 * noise):.4%} (expected L1 of the sampling noise {noise:.4%})")
    assert d <= max(TOL, 1.5 * noise), report[-1]
    pg = z["python_flow_groups"]  # [2, L]
    pf = pg.mean(axis=0)
    var_p = cg.var(axis=0, ddof=1) / pg.shape[0]  # same process: per-link variance of a 32-seed mean taken from the C++ groups
    noise_p = float((np.sqrt(2.0 / np.pi) * np.sqrt(var_g + var_p)).sum() / pf.sum())
    d = float(np.abs(g_flow.mean(axis=0) - pf).sum() / pf.sum())
    own = float(np.abs(cf - pf).sum() / pf.sum())  # the reference's own C++-vs-Python distance
    report.append(f"{name} python link flow L1: {d:.4%} bound {max(TOL, 1.5 * noise_p, own):.4%} (noise {noise_p:.4%}, reference C++ vs Python {own:.4%})")
    assert d <= max(TOL, 1.5 * noise_p, own), report[-1]
    print("\n".join(report))
    # replicas are genuinely different runs
    assert len(set(g_trips.tolist()) | set(np.round(g_att, 6).tolist())) > 8


def test_32_seed_means_within_reference_scatter(cuda_api):
    """The north star's own sample size: 32 GPU seeds against the reference's 32-seed groups (Sioux Falls).  The GPU mean must lie
    within 1 % of the reference's 256-seed mean, or inside the range the reference's own eight 32-seed group means span."""
    z = np.load(os.path.join(GOLDEN, "stat_siouxfalls.npz"))
    sc = NETWORKS["siouxfalls"](7000)
    E = sc.engine_from_tables(cuda_api, sc.tables(), vehicle_logging_timestep_interval=-1, num_replicas=32, random_seed=7000)
    E.main_loop()
    g = [summary_stats(E, r) for r in range(32)]
    E.close()
    for k, ref in ((0, z["cpp_trips"].astype(np.float64)), (1, z["cpp_att"])):
        gm = np.mean([x[k] for x in g])
        groups = ref.reshape(-1, 32).mean(axis=1)
        assert abs(gm / ref.mean() - 1) <= TOL or groups.min() <= gm <= groups.max(), (k, gm, ref.mean(), groups)
