"""`.uxsim_scenario` files in CUDA mode (scenario_reader_writer.py:152-224).  The fixture was written by the
REFERENCE's own `save_scenario` (tests/golden/make_scenario_file_fixture.py); reading it through `CudaWorld.load_scenario`
(vectorised ingest for the plain adddemand records) must define the same scenario as (a) the same records replayed call by call,
and (b) -- dev container -- the reference's own loader; and the reference's dat/chicago_sketch.uxsim_scenario must give the
scenario bench.py runs (the pre-converted table uxsim_b200/data/chicago_sketch.npz)."""
import os
import pickle
import sys

import numpy as np
import pytest

import uxsim_b200
from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "small_grid.uxsim_scenario")
KW = dict(deltan=5, tmax=2400, print_mode=0, save_mode=0, show_mode=0, random_seed=0)


def veh_table(W):
    return [(v.orig.name, v.dest.name, int(round(v.departure_time_in_second / W.DELTAT))) for v in W.VEHICLES.values()]


def replay_call_by_call(W, dat):
    for n in dat["Nodes"]:
        W.addNode(**n)
    for l in dat["Links"]:
        W.addLink(**l)
    for kind in dat:
        if kind.startswith("adddemand"):
            for dem in dat[kind]:
                getattr(W, kind)(**dem)


@pytest.fixture()
def emu_world(emu_api, monkeypatch):
    import uxsim_b200.world as wmod

    monkeypatch.setattr(wmod.C, "load_product", lambda: emu_api)


def test_load_scenario_equals_call_by_call_replay(emu_world):
    Wa = uxsim_b200.World(cuda=True, **KW)
    Wa.load_scenario(FIXTURE)
    Wb = uxsim_b200.World(cuda=True, **KW)
    replay_call_by_call(Wb, pickle.load(open(FIXTURE, "rb")))
    assert len(Wa.VEHICLES) == len(Wb.VEHICLES) == 222
    assert veh_table(Wa) == veh_table(Wb)
    assert [n.name for n in Wa.NODES] == [n.name for n in Wb.NODES] and [l.name for l in Wa.LINKS] == [l.name for l in Wb.LINKS]
    assert dict(Wa.demand_info).keys() == dict(Wb.demand_info).keys()
    Wa.exec_simulation(), Wb.exec_simulation()
    for la, lb in zip(Wa.LINKS, Wb.LINKS):
        assert np.array_equal(la.cum_arrival, lb.cum_arrival) and np.array_equal(la.cum_departure, lb.cum_departure)
    Wa.analyzer.basic_analysis()
    assert Wa.analyzer.trip_all == 222 * 5 and Wa.analyzer.trip_completed > 0


def test_load_scenario_partial(emu_world):
    W = uxsim_b200.World(cuda=True, **KW)
    W.load_scenario(FIXTURE, demand=False)
    assert len(W.NODES) == 16 and len(W.LINKS) == 48 and len(W.VEHICLES) == 0
    W.load_scenario(FIXTURE, network=False)
    assert len(W.VEHICLES) == 222


def test_load_scenario_matches_reference_loader(emu_world):
    """Same file through the reference's own `World.load_scenario` (pure-Python engine): same platoons, same order."""
    if not os.path.isdir("/root/reference/uxsim") and not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "uxsim")):
        pytest.skip("reference package not available")
    os.environ.setdefault("UXSIM_REFERENCE_ROOT", "/root/reference" if os.path.isdir("/root/reference/uxsim") else os.path.join(ROOT, "baseline", "_ref"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from ref_python_shim import import_reference

    ux = import_reference()
    Wr = ux.World(name="", **KW)
    Wr.load_scenario(FIXTURE)
    Wc = uxsim_b200.World(cuda=True, **KW)
    Wc.load_scenario(FIXTURE)
    ref = [(v.orig.name, v.dest.name, int(v.departure_time)) for v in Wr.VEHICLES.values()]
    assert veh_table(Wc) == ref
    for a, b in zip(Wr.LINKS, Wc.LINKS):
        assert (a.name, a.length, a.u, a.number_of_lanes, a.merge_priority, a.capacity_out, a.capacity_in, list(a.signal_group)) == \
               (b.name, b.length, b.u, b.number_of_lanes, b.merge_priority, b.capacity_out, b.capacity_in, list(b.signal_group))
    assert [list(n.signal) for n in Wr.NODES] == [list(n.signal) for n in Wc.NODES]


def test_chicago_sketch_file_gives_the_bench_scenario(emu_world):
    """dat/chicago_sketch.uxsim_scenario (example_23en) through load_scenario == the pre-converted table the bench uses."""
    path = "/root/reference/dat/chicago_sketch.uxsim_scenario"
    if not os.path.exists(path):
        pytest.skip("reference data not present (dev container only)")
    W = uxsim_b200.World(cuda=True, deltan=30, tmax=10000, print_mode=0, save_mode=0, show_mode=0, random_seed=42)
    W.load_scenario(path)
    sc = S.chicago_sketch()
    tab = sc.tables()
    assert len(W.NODES) == 546 and len(W.LINKS) == 2176 and len(W.VEHICLES) == 32284
    E = sc.engine_from_tables(W._E.api, tab, vehicle_logging_timestep_interval=-1)
    for what in (C.A_VEH_ORIG, C.A_VEH_DEST, C.A_VEH_DEPART_TS):
        assert np.array_equal(W._E.get_array(what), E.get_array(what))
    assert np.allclose([l.length for l in W.LINKS], tab["l_length"]) and np.array_equal([l.number_of_lanes for l in W.LINKS], tab["l_lanes"])
    E.close()
