"""Generate the committed STATISTICAL fixtures from the REFERENCE ITSELF (dev container only; needs /root/reference):

    python tests/golden/make_stat_fixtures.py [siouxfalls grid11 chicago_dn30] [--python-seeds 32] [--cpp-seeds 256]

BASELINE.json's stochastic standard: "match the reference Python and C++ modes on total trips completed, average travel time
and per-link cumulative flow, within 1 % on the mean over 32 seeds".  For every network this writes stat_<network>.npz with, for
BOTH reference engines run with their own random streams (Python: numpy Generator seeded by random_seed, uxsim.py; C++: the
unmodified traffi.cpp behind oracle/ref_harness.cpp):

  <mode>_seeds, <mode>_trips, <mode>_att   per-seed trips completed and average travel time of completed trips
  <mode>_flow_groups                       [G, L] mean final cum_arrival per link over disjoint groups of 32 seeds

so that a test can (a) compare the CUDA engine's means with the reference's and (b) read the reference-vs-reference noise
floor of each statistic (spread between disjoint 32-seed groups of the SAME engine) next to any tolerance it uses.
"""
import argparse
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from uxsim_b200 import _capi as C  # noqa: E402
from uxsim_b200 import scenario as S  # noqa: E402

NETWORKS = {
    "siouxfalls": lambda s: S.sioux_falls(seed=s),
    "grid11": lambda s: S.grid(seed=s),
    "chicago_dn30": lambda s: S.chicago_sketch(seed=s),
}
GROUP = 32


def run_python(args):
    name, seed = args
    from ref_python_shim import import_reference

    ux = import_reference()
    sc = NETWORKS[name](seed)
    W = ux.World(**{**sc.world, "vehicle_logging_timestep_interval": -1, "print_mode": 0, "save_mode": 0, "show_mode": 0})
    S.replay_into(sc, W)
    W.exec_simulation()
    tt = np.array([v.travel_time for v in W.VEHICLES.values() if v.state == "end"], dtype=np.float64)
    flow = np.array([l.cum_arrival[-1] for l in W.LINKS], dtype=np.float64)
    return seed, len(tt), float(tt.mean()) if len(tt) else 0.0, flow


def run_cpp(args):
    name, seed = args
    api = C.CApi(os.path.join(ROOT, "oracle", "_ref", "libuxsim_ref.so"), "uxr_")
    sc = NETWORKS[name](seed)
    E = sc.to_engine(api, vehicle_logging_timestep_interval=-1)
    E.main_loop()
    st, tt = E.get_array(C.A_VEH_STATE), E.get_array(C.A_VEH_TRAVEL_TIME)
    done = st == C.VS_END
    L, T = E.get_int(C.I_NUM_LINKS), E.get_int(C.I_TOTAL_TIMESTEPS)
    flow = E.get_array(C.A_CUM_ARRIVAL).reshape(L, T)[:, -1].copy()
    E.close()
    return seed, int(done.sum()), float(tt[done].mean()) if done.any() else 0.0, flow


def collect(pool, fn, name, seeds):
    res = sorted(pool.map(fn, [(name, s) for s in seeds], chunksize=1))
    trips = np.array([r[1] for r in res], np.int64)
    att = np.array([r[2] for r in res], np.float64)
    flows = np.array([r[3] for r in res])
    groups = np.array([flows[g:g + GROUP].mean(axis=0) for g in range(0, len(res) - GROUP + 1, GROUP)])
    return dict(seeds=np.array([r[0] for r in res], np.int64), trips=trips, att=att, flow_groups=groups)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("networks", nargs="*", default=list(NETWORKS))
    ap.add_argument("--python-seeds", type=int, default=32)
    ap.add_argument("--cpp-seeds", type=int, default=256)
    ap.add_argument("--procs", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    a = ap.parse_args()
    with Pool(a.procs) as pool:
        for name in a.networks:
            out = {}
            # seed ranges disjoint from each other and from the ones the CUDA tests use (1000...)
            for mode, fn, seeds in (("cpp", run_cpp, range(2000, 2000 + a.cpp_seeds)), ("python", run_python, range(5000, 5000 + a.python_seeds))):
                for k, v in collect(pool, fn, name, seeds).items():
                    out[f"{mode}_{k}"] = v
                print(name, mode, "trips %.1f att %.2f" % (out[f"{mode}_trips"].mean(), out[f"{mode}_att"].mean()), flush=True)
            np.savez_compressed(os.path.join(HERE, f"stat_{name}.npz"), **out)
