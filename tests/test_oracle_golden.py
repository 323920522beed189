"""Pin the oracle: (1) against golden vectors produced by the reference's pure-Python engine (deterministic
scenarios) and by the reference's C++ engine with Philox draws (stochastic networks) -- runs anywhere;
(2) against the reference's own C++ engine (oracle/_ref) when that library is present -- bit-exact on
deterministic scenarios including full vehicle logs, statistical on stochastic ones."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import PHILOX_CASES, PYTHON_CASES  # noqa: E402
from golden_check import check_against_golden  # noqa: E402
from parity import EXACT_FIELDS, compare_engines, run_pair, summary_stats  # noqa: E402

from uxsim_b200 import _capi as C  # noqa: E402
from uxsim_b200 import scenario as S  # noqa: E402


@pytest.mark.parametrize("name,mk", PYTHON_CASES, ids=[n for n, _ in PYTHON_CASES])
def test_oracle_matches_reference_python_engine(name, mk, oracle_api):
    E = mk().to_engine(oracle_api, vehicle_logging_timestep_interval=-1)
    E.main_loop()
    check_against_golden(E, "python", name)


@pytest.mark.parametrize("name,mk", PHILOX_CASES, ids=[n for n, _ in PHILOX_CASES])
def test_oracle_matches_reference_cpp_engine_with_philox(name, mk, oracle_api):
    E = mk().to_engine(oracle_api, vehicle_logging_timestep_interval=-1)
    E.main_loop()
    check_against_golden(E, "philox", name)


DET = [S.straight_1link, S.straight_2link_bottleneck_capacity_out, S.straight_3link_spillback, S.straight_2link_signal,
       S.straight_node_capacity, S.lane_drop_2to1, S.multilane_3lane, S.merge_y, S.diverge_2route, S.dead_end, S.short_links_chain]
LOG_FIELDS = [(k, getattr(C, k)) for k in dir(C) if k.startswith(("A_LOG_", "A_LTL_", "A_ENTER_"))]


@pytest.mark.parametrize("builder", DET, ids=lambda f: f.__name__)
def test_oracle_bit_exact_vs_reference_cpp(builder, oracle_api, ref_api):
    """Every result array incl. per-platoon logs, FP64 travel times and instantaneous travel times."""
    Eo, Er = run_pair(builder(), oracle_api, ref_api, vehicle_logging_timestep_interval=1)
    fields = [f for f in EXACT_FIELDS if f[0] not in ("route_pref", "route_next", "link_num_vehicles")] + LOG_FIELDS
    bad = []
    for nm, what in fields:
        x, y = Eo.get_array(what), Er.get_array(what)
        if x.shape != y.shape or not np.array_equal(x, y):
            bad.append(nm)
    assert bad == []
    assert Eo.get_int(C.I_PLATOON_UPDATES) == Er.get_int(C.I_PLATOON_UPDATES)


def test_oracle_statistical_parity_sioux_falls(oracle_api, ref_api):
    """BASELINE.json: stochastic networks within 1 % of the reference on the mean over 32 seeds (trips completed,
    average travel time) and per-link cumulative flow (L1 error relative to total flow).  Unmodified reference RNG."""
    res = {}
    for tag, api in (("oracle", oracle_api), ("ref", ref_api)):
        trips, att, flow = [], [], []
        for seed in range(32):
            E = S.sioux_falls(seed=seed).to_engine(api, vehicle_logging_timestep_interval=-1)
            E.main_loop()
            a, b, c = summary_stats(E)
            trips.append(a), att.append(b), flow.append(c)
            E.close()
        res[tag] = (np.mean(trips), np.mean(att), np.mean(flow, axis=0))
    o, r = res["oracle"], res["ref"]
    assert abs(o[0] / r[0] - 1) < 0.01
    assert abs(o[1] / r[1] - 1) < 0.01
    assert np.abs(o[2] - r[2]).sum() / r[2].sum() < 0.02
