"""The hand-back contract ("PyTorch / numpy arrays only to hand results back to analyzer.py"): the reference's OWN,
UNMODIFIED `uxsim.analyzer.Analyzer` runs over a `CudaWorld` and produces what it produces over the reference engine.

Reference side: the reference's pure-Python engine on the same deterministic scenario (the C++ mode cannot be built here --
nanobind is absent -- and on these scenarios the two reference engines agree bit for bit).  The reference package
is imported from /root/reference in the dev container and from its staged copy baseline/_ref (baseline/stage_reference.py)
on the GPU box; the import shim (oracle/ref_python_shim.py, test infrastructure) only stubs matplotlib.

CPU variant: the wrapper over the host emulation of the kernels; `-m gpu` variant: the CUDA engine itself."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = "/root/reference" if os.path.isdir("/root/reference/uxsim") else os.path.join(ROOT, "baseline", "_ref")
os.environ.setdefault("UXSIM_REFERENCE_ROOT", REF_ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from uxsim_b200 import scenario as S  # noqa: E402

SCENARIOS = [
    ("3link_spillback", S.straight_3link_spillback), ("lane_drop", S.lane_drop_2to1), ("2link_signal", S.straight_2link_signal),
    ("merge_unequal", lambda: S.merge_y(priority1=1, priority2=2)), ("node_capacity", S.straight_node_capacity),
]


def reference_package():
    if not os.path.isdir(os.path.join(REF_ROOT, "uxsim")):
        pytest.skip("the reference package is neither at /root/reference nor staged under baseline/_ref")
    pytest.importorskip("pandas")
    from ref_python_shim import import_reference

    return import_reference()


def analyzer_outputs(W):
    """Everything the reference analyzer derives from a finished world, as plain numpy / pandas values."""
    A = W.analyzer
    A.basic_analysis()
    out = {"scalars": (A.trip_all, A.trip_completed, A.total_travel_time, A.average_travel_time, A.total_delay, A.average_delay)}
    A.od_analysis()
    out["od"] = sorted((o.name, d.name, int(A.od_trips[o, d]), int(A.od_trips_comp[o, d]), float(A.od_tt_free[o, d]), float(A.od_tt_ave[o, d]),
                        float(A.od_tt_std[o, d]), float(A.od_dist_min[o, d]), float(A.od_dist_ave[o, d])) for (o, d) in A.od_trips)
    A.compute_edie_state()
    out["edie"] = [(np.array(l.k_mat), np.array(l.q_mat), np.array(l.v_mat)) for l in W.LINKS]
    for name in ("basic_to_pandas", "od_to_pandas", "link_to_pandas", "link_traffic_state_to_pandas", "link_cumulative_to_pandas",
                 "vehicles_to_pandas", "vehicle_trip_to_pandas"):
        if hasattr(A, name):
            out[name] = getattr(A, name)()
    return out


def assert_same(a, b):
    import pandas as pd

    assert a["scalars"] == b["scalars"]
    assert a["od"] == b["od"]
    for (k1, q1, v1), (k2, q2, v2) in zip(a["edie"], b["edie"]):
        assert k1.shape == k2.shape
        np.testing.assert_allclose(k1, k2, rtol=1e-9, atol=1e-12); np.testing.assert_allclose(q1, q2, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(v1, v2, rtol=1e-9, atol=1e-9)
    for name, df in a.items():
        if isinstance(df, pd.DataFrame):
            pd.testing.assert_frame_equal(df.reset_index(drop=True), b[name].reset_index(drop=True), check_dtype=False, rtol=1e-9, atol=1e-9)


def run_both(mk, ux):
    sc = mk()
    Wr = ux.World(**{**sc.world, "print_mode": 0, "save_mode": 0, "show_mode": 0})
    S.replay_into(sc, Wr)
    Wr.exec_simulation()
    Wc = sc.to_world()
    Wc.exec_simulation()
    assert type(Wc.analyzer).__module__ == "uxsim.analyzer"  # the reference's own class, not this package's stand-in
    return analyzer_outputs(Wr), analyzer_outputs(Wc)


@pytest.mark.parametrize("name,mk", SCENARIOS, ids=[c[0] for c in SCENARIOS])
def test_reference_analyzer_over_cuda_world_emulated(name, mk, emu_api, monkeypatch):
    ux = reference_package()
    import uxsim_b200.world as wmod

    monkeypatch.setattr(wmod.C, "load_product", lambda: emu_api)
    a, b = run_both(mk, ux)
    assert_same(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("name,mk", SCENARIOS, ids=[c[0] for c in SCENARIOS])
def test_reference_analyzer_over_cuda_world(name, mk, cuda_api):
    ux = reference_package()
    a, b = run_both(mk, ux)
    assert_same(a, b)
