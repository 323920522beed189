"""GPU parity tests proper: the CUDA engine, called through the C-ABI, must be bit-identical to the oracle.

Bar (BASELINE.json north_star): bit-exact integer cumulative counts and arrival timesteps on deterministic
scenarios.  Because the oracle and the kernels draw from the same counter-based Philox stream, the same
bit-exactness is demanded here on STOCHASTIC networks too, for every result array (FP64 included)."""
import numpy as np
import pytest

from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S

from parity import compare_engines, compare_logs, run_pair, summary_stats

pytestmark = pytest.mark.gpu

DETERMINISTIC = [
    S.straight_1link, S.straight_2link_bottleneck_capacity_out, S.straight_2link_bottleneck_speed,
    S.straight_3link_spillback, S.straight_2link_signal, S.straight_node_capacity, S.lane_drop_2to1,
    S.multilane_3lane, S.merge_y, S.diverge_2route, S.dead_end, S.short_links_chain,
]


@pytest.mark.parametrize("builder", DETERMINISTIC, ids=lambda f: f.__name__)
def test_deterministic_bit_exact(builder, cuda_api, oracle_api):
    Eg, Eo = run_pair(builder(), cuda_api, oracle_api)
    assert compare_engines(Eg, Eo) == []
    assert Eg.get_int(C.I_PLATOON_UPDATES) == Eo.get_int(C.I_PLATOON_UPDATES)


def test_short_links_order_matters(cuda_api, oracle_api):
    """Serial id-ordered Node.transfer visibility: node ids decreasing downstream on links
    shorter than one platoon spacing; needs >1 dependency level and must still be exact."""
    Eg, Eo = run_pair(S.short_links_chain(reverse_ids=True), cuda_api, oracle_api)
    assert Eg.get_int(C.I_TRANSFER_LEVELS) > 1
    assert compare_engines(Eg, Eo) == []


STOCHASTIC = [
    ("merge_stochastic", lambda: S.merge_y(hard_deterministic_mode=False, random_seed=3)),
    ("siouxfalls", lambda: S.sioux_falls(seed=1)),
    ("grid5", lambda: S.grid(n=5, seed=2, tmax=3000, demand_t=1000)),
    ("grid6_signals", lambda: S.grid(n=6, seed=5, tmax=3000, demand_t=1200, signals=True)),
    ("grid11", lambda: S.grid(seed=2)),
    ("dta_grid9", lambda: S.dta_grid9(seed=1)),
    ("chicago_dn30", lambda: S.chicago_sketch(seed=0)),
    ("gradual_duo", lambda: S.grid(n=5, seed=4, tmax=3000, demand_t=1000, route_choice_update_gradual=True)),
]


@pytest.mark.parametrize("name,mk", STOCHASTIC, ids=[n for n, _ in STOCHASTIC])
def test_stochastic_bit_exact(name, mk, cuda_api, oracle_api):
    Eg, Eo = run_pair(mk(), cuda_api, oracle_api)
    assert compare_engines(Eg, Eo) == []
    assert Eg.get_int(C.I_PLATOON_UPDATES) == Eo.get_int(C.I_PLATOON_UPDATES)
    assert Eg.get_int(C.I_TRIPS_COMPLETED) == Eo.get_int(C.I_TRIPS_COMPLETED)


@pytest.mark.parametrize("mode", ["wave", "frontier", "sweep"])
@pytest.mark.parametrize("name", ["siouxfalls", "grid6_signals", "chicago_dn30"])
def test_shortest_path_kernel_variants_bit_exact(name, mode, cuda_api, oracle_api, monkeypatch):
    """Every route-search kernel (whole-edge sweeps, shared-memory worklist, L2-resident near-far wavefront) reaches the same
    fixed point: forced on networks of any size with UXSIM_B200_SSSP, results stay bit-identical to the oracle."""
    monkeypatch.setenv("UXSIM_B200_SSSP", mode)
    Eg, Eo = run_pair(dict(STOCHASTIC)[name](), cuda_api, oracle_api)
    assert compare_engines(Eg, Eo) == []


def test_staged_route_search_equals_unstaged(cuda_api, monkeypatch):
    """k_sssp_wave<smem, warp, STAGE> (the in-edge tables of a replica staged in shared memory; off by default, measured slower)
    against the default kernel on an ensemble large enough to take it (>= 16384 destination rows per launch)."""
    sc = dict(STOCHASTIC)["chicago_dn30"]()
    out = []
    for stage in ("0", "1"):
        monkeypatch.setenv("UXSIM_B200_SSSP_STAGE", stage)
        E = sc.to_engine(cuda_api, vehicle_logging_timestep_interval=-1, num_replicas=56)
        E.main_loop(until_t=1500.0)
        out.append([(E.get_array(C.A_ROUTE_NEXT, r), E.get_array(C.A_ROUTE_DIST, r), E.get_array(C.A_CUM_ARRIVAL, r)) for r in (0, 31, 55)])
        E.close()
    for a, b in zip(*out):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_chunked_execution_equals_single_run(cuda_api, oracle_api):
    """exec_simulation(until_t=...) in chunks (test_verification_straight_road.py:940-1113) == one run."""
    sc = S.grid(n=5, seed=9, tmax=3000, demand_t=1000)
    Ea, Eo = run_pair(sc, cuda_api, oracle_api, until=[300, 305, 1234, 2000, 3000])
    assert compare_engines(Ea, Eo) == []
    Eb = sc.to_engine(cuda_api, vehicle_logging_timestep_interval=-1)
    Eb.main_loop()
    assert compare_engines(Ea, Eb) == []


def test_replicas_match_independent_seeds(cuda_api, oracle_api):
    """R replicas in one world == R single-seed oracle runs with seeds seed + r."""
    sc = S.grid(n=5, seed=11, tmax=3000, demand_t=1000)
    R = 4
    Eg = sc.to_engine(cuda_api, vehicle_logging_timestep_interval=-1, num_replicas=R)
    Eg.main_loop()
    for r in range(R):
        Eo = sc.to_engine(oracle_api, vehicle_logging_timestep_interval=-1, random_seed=11 + r)
        Eo.main_loop()
        assert compare_engines(Eg, Eo, replica_a=r) == [], f"replica {r}"
    ens = Eg.ensemble_link_stats()
    vol = sum(summary_stats(Eg, r)[2] for r in range(R))
    assert np.array_equal(ens[:, 0], vol)


@pytest.mark.parametrize("name,R,force", [("chicago_dn30", 40, True), ("grid6_signals", 33, True), ("siouxfalls", 8, True), ("grid6_signals", 64, False)])
def test_replica_lane_update_kernel(name, R, force, cuda_api, oracle_api, monkeypatch):
    """Large ensembles run the link update with one thread per (link, replica) (k_update_rl; picked by itself from 64 replicas
    on lightly loaded networks, forced here for ragged replica counts): every replica still equals the single-seed oracle run."""
    if force:
        monkeypatch.setenv("UXSIM_B200_UPDATE", "rl")
    sc = dict(STOCHASTIC)[name]()
    Eg = sc.to_engine(cuda_api, vehicle_logging_timestep_interval=-1, num_replicas=R)
    Eg.main_loop()
    for r in (0, 7, R - 1):
        Eo = sc.to_engine(oracle_api, vehicle_logging_timestep_interval=-1, random_seed=sc.world.get("random_seed", 0) + r)
        Eo.main_loop()
        assert compare_engines(Eg, Eo, replica_a=r) == [], f"replica {r}"


@pytest.mark.parametrize("group", ["16", "32"])
@pytest.mark.parametrize("name", ["3lane", "grid6_signals", "chicago_dn30"])
def test_update_kernel_lane_groups_bit_exact(name, group, cuda_api, oracle_api, monkeypatch):
    """k_update with one (link, part) work item per warp or two (16-lane groups; picked by itself for large, lightly loaded
    worlds): forced both ways, logs on -- identical results."""
    monkeypatch.setenv("UXSIM_B200_UPDATE_GROUP", group)
    mk = S.multilane_3lane if name == "3lane" else dict(STOCHASTIC)[name]
    Eg, Eo = run_pair(mk(), cuda_api, oracle_api, vehicle_logging_timestep_interval=1)
    assert compare_engines(Eg, Eo) == [] and compare_logs(Eg, Eo) == []


@pytest.mark.parametrize("mode", ["rl", "warp"])
@pytest.mark.parametrize("name,mk", [("3lane", S.multilane_3lane), ("short_chain", S.short_links_chain), ("grid6_signals", dict(STOCHASTIC)["grid6_signals"]),
                                     ("chicago_dn30", dict(STOCHASTIC)["chicago_dn30"]), ("chicago_dn5", lambda: S.chicago_sketch(deltan=5, seed=0, tmax=2500))],
                         ids=["3lane", "short_chain", "grid6_signals", "chicago_dn30", "chicago_dn5"])
def test_update_kernel_variants_bit_exact(name, mk, mode, cuda_api, oracle_api, monkeypatch):
    """Both forms of the link update (warp per link / thread per link and replica) forced on single scenarios, logs on:
    long queues (Chicago at deltan=5) exercise the warp-cooperative tail of k_update_rl."""
    monkeypatch.setenv("UXSIM_B200_UPDATE", mode)
    log = 1 if name != "chicago_dn5" else -1
    Eg, Eo = run_pair(mk(), cuda_api, oracle_api, vehicle_logging_timestep_interval=log)
    assert compare_engines(Eg, Eo) == []
    if log == 1:
        assert compare_logs(Eg, Eo) == []


def test_c4_full_size_route_search_and_conservation(cuda_api, monkeypatch):
    """BASELINE configs[3] at full size (100x100 signalised grid, 1 M platoons = 5 M vehicles, 10 000 destinations), first three
    route updates: the L2-resident near-far route search and the shared-memory FIFO worklist reach the same fixed point (dist and
    next hop of all 10^8 pairs identical, hence identical traffic), and the size-independent properties hold."""
    sc = S.grid_random_od(n=100, n_platoons=1_000_000, deltan=5, tmax=7200, seed=0)
    tab = sc.tables()
    res = {}
    for mode in ("wave", "frontier"):
        monkeypatch.setenv("UXSIM_B200_SSSP", mode)
        E = sc.engine_from_tables(cuda_api, tab, vehicle_logging_timestep_interval=-1, random_seed=0)
        E.main_loop(until_t=1300)
        L, T = E.get_int(C.I_NUM_LINKS), E.get_int(C.I_TOTAL_TIMESTEPS)
        res[mode] = (E.get_array(C.A_ROUTE_NEXT), E.get_array(C.A_ROUTE_DIST), E.get_array(C.A_CUM_ARRIVAL).reshape(L, T)[:, :261].copy(),
                     E.get_array(C.A_CUM_DEPARTURE).reshape(L, T)[:, :261].copy(), E.get_array(C.A_LINK_NUM_VEHICLES), E.get_int(C.I_PLATOON_UPDATES))
        E.close()
    a, b = res["wave"], res["frontier"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), "dist / next differ between the route-search kernels"
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]) and a[5] == b[5] > 4e7
    arr, dep, nveh = a[2], a[3], a[4]
    assert (np.diff(arr, axis=1) >= 0).all() and (np.diff(dep, axis=1) >= 0).all() and (arr >= dep).all()
    assert np.array_equal((arr[:, -1] - dep[:, -1]) / 5.0, nveh)  # platoons on a link = entered - left


@pytest.mark.parametrize("split", ["0", "1"])
@pytest.mark.parametrize("name", ["grid6_signals", "dta_grid9", "chicago_dn30"])
def test_scalar_bookkeeping_kernel_split_bit_exact(name, split, cuda_api, oracle_api, monkeypatch):
    """The per-link / per-node scalar step (tokens, curve carries, signals) either runs inside k_node (small worlds) or as its
    own thread-per-link kernel k_link_pre (large worlds): forced both ways, results are bit-identical to the oracle."""
    monkeypatch.setenv("UXSIM_B200_PRE_SPLIT", split)
    Eg, Eo = run_pair(dict(STOCHASTIC)[name](), cuda_api, oracle_api)
    assert compare_engines(Eg, Eo) == []


def test_run_to_run_determinism(cuda_api):
    sc = S.sioux_falls(seed=5)
    Ea = sc.to_engine(cuda_api, vehicle_logging_timestep_interval=-1)
    Eb = sc.to_engine(cuda_api, vehicle_logging_timestep_interval=-1)
    Ea.main_loop(), Eb.main_loop()
    assert compare_engines(Ea, Eb) == []


def test_conservation_at_full_size(cuda_api):
    """Size-independent properties at a BASELINE size (Chicago sketch, deltan=5, ~1M vehicles): flow conservation
    per link, monotone curves, arrivals bounded by departures upstream, every finished trip has a valid step."""
    sc = S.chicago_sketch(deltan=5, seed=0)
    E = sc.to_engine(cuda_api, vehicle_logging_timestep_interval=-1)
    E.main_loop()
    L, T = E.get_int(C.I_NUM_LINKS), E.get_int(C.I_TOTAL_TIMESTEPS)
    arr = E.get_array(C.A_CUM_ARRIVAL).reshape(L, T)
    dep = E.get_array(C.A_CUM_DEPARTURE).reshape(L, T)
    nveh = E.get_array(C.A_LINK_NUM_VEHICLES)
    assert (np.diff(arr, axis=1) >= 0).all() and (np.diff(dep, axis=1) >= 0).all()
    assert (arr >= dep).all()
    assert np.array_equal(arr[:, -1] - dep[:, -1], nveh * 5.0)
    st = E.get_array(C.A_VEH_STATE)
    ats = E.get_array(C.A_VEH_ARRIVAL_TS)
    dts = E.get_array(C.A_VEH_DEPART_TS)
    done = st == C.VS_END
    assert done.sum() > 0.8 * len(st)
    assert (ats[done] > dts[done]).all() and (ats[~done] == -1).all()
    assert E.get_int(C.I_TRIPS_COMPLETED) == int(done.sum() + (st == C.VS_ABORT).sum())
    assert int(dep[:, -1].sum() / 5.0) == E.get_int(C.I_LINK_EVENTS)


def test_error_codes(cuda_api):
    E = S.straight_1link().to_engine(cuda_api)
    with pytest.raises(RuntimeError):
        E.main_loop(duration_t=10, until_t=10)
    with pytest.raises(IndexError):
        E.set_link_param(99, C.LINK_VMAX, 10.0)


LOG_CASES = [("spillback", S.straight_3link_spillback), ("3lane", S.multilane_3lane), ("lane_drop", S.lane_drop_2to1), ("dead_end", S.dead_end),
             ("short_chain", S.short_links_chain), ("merge_stochastic", lambda: S.merge_y(hard_deterministic_mode=False, random_seed=3)),
             ("grid6_signals", lambda: S.grid(n=6, seed=5, tmax=3000, demand_t=1200, signals=True)), ("siouxfalls", lambda: S.sioux_falls(seed=1)),
             ("chicago_dn30", lambda: S.chicago_sketch(seed=0))]


@pytest.mark.parametrize("name,mk", LOG_CASES, ids=[c[0] for c in LOG_CASES])
def test_vehicle_logs_bit_exact(name, mk, cuda_api, oracle_api):
    """vehicle_logging_timestep_interval=1 (the API default): every per-platoon log array in the reference's compact flat
    layout (traffi.cpp:1990-2081), log_t_link and the enter log, mid-run and at the end."""
    sc = mk()
    Eg, Eo = run_pair(sc, cuda_api, oracle_api, until=[0.4 * sc.tmax()], vehicle_logging_timestep_interval=1)
    assert compare_logs(Eg, Eo) == []
    Eg.main_loop(), Eo.main_loop()
    assert compare_logs(Eg, Eo) == [] and compare_engines(Eg, Eo) == []
    assert Eg.get_int(C.I_LOG_ENTRIES) == Eo.get_int(C.I_LOG_ENTRIES)


def test_grid_random_od_with_signals(cuda_api, oracle_api):
    """Small variant of BASELINE config C4 (grid, signal at every node, random OD, addVehicle departures)."""
    sc = S.grid_random_od(n=12, n_platoons=20000, tmax=4000, demand_t=1500, seed=3)
    Eg, Eo = run_pair(sc, cuda_api, oracle_api)
    assert compare_engines(Eg, Eo) == []


def test_c4_full_size_properties(cuda_api):
    """BASELINE config C4 at full size (100x100 grid, 1M platoons = 5M vehicles, 1440 steps): conservation properties."""
    sc = S.grid_random_od(n=100, n_platoons=1_000_000, seed=0)
    E = sc.engine_from_tables(cuda_api, sc.tables(), vehicle_logging_timestep_interval=-1)
    E.main_loop()
    L, T = E.get_int(C.I_NUM_LINKS), E.get_int(C.I_TOTAL_TIMESTEPS)
    arr = E.get_array(C.A_CUM_ARRIVAL).reshape(L, T)
    dep = E.get_array(C.A_CUM_DEPARTURE).reshape(L, T)
    assert (np.diff(arr, axis=1) >= 0).all() and (arr >= dep).all()
    assert np.array_equal(arr[:, -1] - dep[:, -1], E.get_array(C.A_LINK_NUM_VEHICLES) * 5.0)
    st = E.get_array(C.A_VEH_STATE)  # (the scenario is far over capacity by design: most platoons are still queued at the end)
    assert (st == C.VS_END).sum() + (st == C.VS_RUN).sum() + (st == C.VS_WAIT).sum() == len(st)
    assert (st == C.VS_END).sum() > 0
    assert E.get_int(C.I_TRIPS_COMPLETED) == int((st == C.VS_END).sum())
    assert int(arr[:, -1].sum() / 5.0) - int(dep[:, -1].sum() / 5.0) == int((st == C.VS_RUN).sum())


@pytest.mark.parametrize("log", [-1, 1])
def test_demand_added_between_chunks(log, cuda_api, oracle_api):
    """adddemand / addVehicle after the simulation started (the reference rebuilds its departure buckets on every
    main_loop and clamps past-due departures to the chunk start, traffi.cpp:1663-1676)."""
    sc = S.grid(n=5, seed=9, tmax=3000, demand_t=1000)
    nidx = sc.node_index()
    Es = [sc.to_engine(api, vehicle_logging_timestep_interval=log) for api in (cuda_api, oracle_api)]
    for E in Es:
        E.main_loop(until_t=700)
        E.add_demand(nidx["n(0, 0)"], nidx["n(4, 4)"], 300, 1500, 0.3)
        E.add_vehicles([nidx["n(4, 0)"], nidx["n(0, 2)"]], [nidx["n(0, 4)"], nidx["n(4, 2)"]], [10, 400])
        E.main_loop(until_t=1400)
        E.add_demand(nidx["n(2, 0)"], nidx["n(2, 4)"], 1500, 2000, 0.2)
        E.main_loop()
    assert compare_engines(Es[0], Es[1]) == []
    if log > 0:
        assert compare_logs(Es[0], Es[1]) == []


def test_new_destination_midrun_needs_all_destinations(cuda_api, oracle_api):
    sc = S.straight_3link_spillback()
    E = sc.to_engine(cuda_api, vehicle_logging_timestep_interval=-1)
    E.main_loop(until_t=300)
    E.add_vehicles([0], [1], [100])  # node 1 was never a destination
    with pytest.raises(RuntimeError):
        E.main_loop()
    kw = sc.engine_kwargs(vehicle_logging_timestep_interval=-1)
    Es = []
    for api in (cuda_api, oracle_api):
        E = sc.to_engine(api, vehicle_logging_timestep_interval=-1)
        Es.append(E)
    # with the option the same sequence is exact
    E2 = C.EngineWorld(cuda_api, **kw)
    E2.set_option(C.OPT_ALL_DESTINATIONS, 1)
    for n in sc.nodes:
        E2.add_node(n["x"], n["y"], n["signal"], n["signal_offset"])
    nidx = sc.node_index()
    for l in sc.links:
        E2.add_link(nidx[l["start_node"]], nidx[l["end_node"]], l["free_flow_speed"], l["jam_density"], l["length"], l["number_of_lanes"],
                    l["merge_priority"], -1.0 if l["capacity_out"] is None else l["capacity_out"], -1.0 if l["capacity_in"] is None else l["capacity_in"])
    for d in sc.demands:
        E2.add_demand(nidx[d[1]], nidx[d[2]], d[3], d[4], d[5])
    E2.initialize_adj_matrix()
    Eo = Es[1]
    for E in (E2, Eo):
        E.main_loop(until_t=300)
        E.add_vehicles([0], [1], [100])
        E.main_loop()
    assert compare_engines(E2, Eo, fields=[f for f in __import__("parity").EXACT_FIELDS if f[0] not in ("route_pref", "route_next")]) == []


def test_large_graph_worklist_shortest_paths(cuda_api, oracle_api):
    """N >= 2048 nodes switches route search to the frontier (worklist) kernel: same fixed point, same results."""
    sc = S.grid_random_od(n=46, n_platoons=6000, tmax=2400, demand_t=900, seed=5)
    Eg, Eo = run_pair(sc, cuda_api, oracle_api)
    assert compare_engines(Eg, Eo) == []


def test_bench_shape_chicago_128_replicas(cuda_api, oracle_api):
    """The configuration bench.py measures: Chicago sketch, 128 replicas in one world, ingested from flat tables, kernels picked
    by the engine itself (ensemble forms of the step kernels, warp-per-destination route search) -- replicas 0, 63 and 127
    equal the single-seed oracle runs bit for bit."""
    sc = S.chicago_sketch()
    tab = sc.tables()
    R, seed0 = 128, 1000
    Eg = sc.engine_from_tables(cuda_api, tab, vehicle_logging_timestep_interval=-1, num_replicas=R, random_seed=seed0)
    Eg.main_loop()
    for r in (0, 63, 127):
        Eo = sc.engine_from_tables(oracle_api, tab, vehicle_logging_timestep_interval=-1, random_seed=seed0 + r)
        Eo.main_loop()
        assert compare_engines(Eg, Eo, replica_a=r) == [], f"replica {r}"
        assert summary_stats(Eg, r)[0] == summary_stats(Eo)[0]
        Eo.close()
    Eg.close()


def test_bench_shape_chicago_2048_replicas(cuda_api, oracle_api):
    """bench.py's default since round 2: 2048 replicas in one world (68 GB of state) -- the node step with a thread per (node,
    replica), the link update without helper blocks, the warp-per-destination route search, all picked by the engine itself.
    Replicas 0, 1023 and 2047 equal the single-seed oracle runs bit for bit; the replicas are genuinely different runs."""
    sc = S.chicago_sketch()
    tab = sc.tables()
    R, seed0 = 2048, 300000
    Eg = sc.engine_from_tables(cuda_api, tab, vehicle_logging_timestep_interval=-1, num_replicas=R, random_seed=seed0)
    Eg.main_loop()
    for r in (0, 1023, 2047):
        Eo = sc.engine_from_tables(oracle_api, tab, vehicle_logging_timestep_interval=-1, random_seed=seed0 + r)
        Eo.main_loop()
        assert compare_engines(Eg, Eo, replica_a=r) == [], f"replica {r}"
        Eo.close()
    arr = Eg.get_array_all(C.A_VEH_ARRIVAL_TS)
    assert arr.shape == (R, 32284) and len({int((a >= 0).sum()) for a in arr}) > 32  # genuinely different runs
    Eg.close()


def test_c4_full_size_against_oracle(cuda_api, oracle_api):
    """BASELINE configs[3] at FULL size (100x100 signalised grid, 1 M platoons, 10 000 destinations) against the oracle for the
    first 600 s of simulated time (two route updates, 10 M platoon-updates; about a minute and 6.5 GB on the CPU side)."""
    sc = S.grid_random_od(n=100, n_platoons=1_000_000, deltan=5, tmax=7200, seed=0)
    tab = sc.tables()
    Eg = sc.engine_from_tables(cuda_api, tab, vehicle_logging_timestep_interval=-1, random_seed=0)
    Eo = sc.engine_from_tables(oracle_api, tab, vehicle_logging_timestep_interval=-1, random_seed=0)
    Eg.main_loop(until_t=600), Eo.main_loop(until_t=600)
    assert Eg.get_int(C.I_PLATOON_UPDATES) == Eo.get_int(C.I_PLATOON_UPDATES) > 9e6
    assert compare_engines(Eg, Eo) == []
    Eg.close(); Eo.close()


RL_NODE_CASES = [("short_chain", S.short_links_chain, 1), ("3lane", S.multilane_3lane, 1), ("dead_end", S.dead_end, 1),
                 ("grid6_signals", dict(STOCHASTIC)["grid6_signals"], 1), ("siouxfalls", dict(STOCHASTIC)["siouxfalls"], 1),
                 ("chicago_dn30", dict(STOCHASTIC)["chicago_dn30"], 1), ("chicago_dn5", lambda: S.chicago_sketch(deltan=5, seed=0, tmax=2500), -1),
                 ("gradual_duo", dict(STOCHASTIC)["gradual_duo"], 1)]


@pytest.mark.parametrize("name,mk,log", RL_NODE_CASES, ids=[c[0] for c in RL_NODE_CASES])
def test_replica_lane_node_kernel_forced(name, mk, log, cuda_api, oracle_api, monkeypatch):
    """k_link_pre<true> + k_node_rl (thread per node and replica) forced on single scenarios (one live lane per warp), logs on:
    bit-identical to the oracle incl. the multi-level Chicago sketch and its long queues at deltan = 5."""
    monkeypatch.setenv("UXSIM_B200_NODE", "rl")
    Eg, Eo = run_pair(mk(), cuda_api, oracle_api, vehicle_logging_timestep_interval=log)
    assert compare_engines(Eg, Eo) == []
    if log == 1:
        assert compare_logs(Eg, Eo) == []


@pytest.mark.parametrize("name,R,mode", [("chicago_dn30", 35, "rl"), ("grid6_signals", 64, "hybrid"), ("siouxfalls", 33, "hybrid"), ("grid11", 32, "rl"),
                                         ("chicago_dn30", 40, "hybrid"), ("grid6_signals", 256, None)])
def test_replica_lane_node_kernel_ensembles(name, R, mode, cuda_api, oracle_api, monkeypatch):
    """The node step with a thread per (node, replica) (k_node_h; picked by itself from 256 replicas, forced below that) on
    ensembles with ragged lane groups, in its pure form and with the heavy nodes stepped by warps: every checked replica equals
    the single-seed oracle run."""
    if mode:
        monkeypatch.setenv("UXSIM_B200_NODE", mode)
        monkeypatch.setenv("UXSIM_B200_NODE_HEAVY", "4" if mode == "hybrid" else "1000000")
    sc = dict(STOCHASTIC)[name]()
    seed0 = sc.world.get("random_seed", 0) or 0
    Eg = sc.to_engine(cuda_api, vehicle_logging_timestep_interval=-1, num_replicas=R)
    Eg.main_loop()
    for r in sorted({0, 1, 31, R - 1}):
        Eo = sc.to_engine(oracle_api, vehicle_logging_timestep_interval=-1, random_seed=seed0 + r)
        Eo.main_loop()
        assert compare_engines(Eg, Eo, replica_a=r) == [], f"replica {r}"
        Eo.close()
    Eg.close()


def test_node_kernel_forms_agree_on_an_ensemble(cuda_api, monkeypatch):
    sc = dict(STOCHASTIC)["chicago_dn30"]()
    Es = []
    for mode in ("rl", "warp"):
        monkeypatch.setenv("UXSIM_B200_NODE", mode)
        E = sc.to_engine(cuda_api, vehicle_logging_timestep_interval=-1, num_replicas=40)
        E.main_loop()
        Es.append(E)
    for r in (0, 17, 39):
        assert compare_engines(Es[0], Es[1], replica_a=r, replica_b=r) == []
