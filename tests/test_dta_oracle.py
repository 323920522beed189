"""DTA solver batch operations: the oracle's restatement of dta_solver.cpp against the reference's own
code (oracle/_ref, built from /root/reference in the dev container) -- bit-exact for every output, since these functions
use the reference's own mt19937 streams."""
import numpy as np
import pytest

from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S

SCENARIOS = {
    "diverge": lambda: S.diverge_2route(),
    "grid5": lambda: S.grid(5, seed=3),
    "siouxfalls": lambda: S.sioux_falls(),
    "dta_grid9": lambda: S.dta_grid9(),
}


def first_route_csr(E, pick=0):
    """Per-vehicle CSR enforcing route number `pick` (mod the set size) of the vehicle's OD pair from the store."""
    o, d, od_off, r_off, links = E.dta_route_sets_csr()
    key = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(o, d))}
    vo, vd = E.get_array(C.A_VEH_ORIG), E.get_array(C.A_VEH_DEST)
    ids, off = [], [0]
    for a, b in zip(vo, vd):
        p = key.get((int(a), int(b)))
        if p is not None and od_off[p + 1] > od_off[p]:
            r = od_off[p] + pick % (od_off[p + 1] - od_off[p])
            ids.extend(links[r_off[r]:r_off[r + 1]])
        off.append(len(ids))
    return np.asarray(ids, np.int32), np.asarray(off, np.int64)


def run(api, name, seed=0, k=6, enum_seed=99, pick=0, **kw):
    """A finished run in which every platoon follows an enforced route (as in DTA iterations >= 1), so that the oracle and
    the reference hold the same state whatever their shortest-path tie-breaking is."""
    sc = SCENARIOS[name]()
    E = sc.to_engine(api, vehicle_logging_timestep_interval=1, random_seed=seed, **kw)
    E.set_option(C.OPT_RECORD_TRAVERSALS, 1)  # the CUDA engine records link entries instead of reading logs; a no-op elsewhere
    E.dta_enumerate_route_sets(k, enum_seed)
    E.batch_enforce_routes_csr(*first_route_csr(E, pick))
    E.main_loop()
    return E


def assert_swap_equal(a, b):
    for k in ("rs_offsets", "rs_link_ids", "ra_offsets", "ra_link_ids"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["cost_actual"], b["cost_actual"])
    for k in ("n_swap", "potential_n_swap", "total_t_gap"):
        assert a[k] == b[k], (k, a[k], b[k])


@pytest.mark.parametrize("name", ["diverge", "grid5", "siouxfalls"])
def test_enumerate_route_sets_matches_reference(oracle_api, ref_api, name):
    sc = SCENARIOS[name]()
    Eo, Er = sc.to_engine(oracle_api), sc.to_engine(ref_api)
    for k, seed in ((3, 1), (10, 12345)):
        no, nr = Eo.dta_enumerate_route_sets(k, seed), Er.dta_enumerate_route_sets(k, seed)
        assert no == nr and no > 0
        for x, y in zip(Eo.dta_route_sets_csr(), Er.dta_route_sets_csr()):
            assert np.array_equal(x, y)
    Eo.close(); Er.close()


@pytest.mark.parametrize("name", ["diverge", "grid5", "siouxfalls", "dta_grid9"])
@pytest.mark.parametrize("mode", [0, 1])
def test_route_swap_matches_reference(oracle_api, ref_api, name, mode):
    # deterministic simulation (hard_deterministic_mode) so that both engines hold the same finished run
    Eo, Er = (run(api, name, hard_deterministic_mode=True) for api in (oracle_api, ref_api))
    assert np.array_equal(Eo.get_array(C.A_VEH_ARRIVAL_TS), Er.get_array(C.A_VEH_ARRIVAL_TS))
    for swap_prob, rng_seed in ((0.05, 7), (0.5, 8), (1.0, 9)):
        a, b = Eo.dta_route_swap(mode, swap_prob, rng_seed), Er.dta_route_swap(mode, swap_prob, rng_seed)
        assert_swap_equal(a, b)
        assert np.array_equal(Eo.get_array(C.A_DTA_RA_ENTRY_T), Er.get_array(C.A_DTA_RA_ENTRY_T))
    Eo.close(); Er.close()


# ---- the product's kernels (host emulation build of engine.cu, CPU only) against the oracle ------------------------------
@pytest.mark.parametrize("name", ["diverge", "grid5", "siouxfalls"])
def test_emu_enumerate_matches_oracle(oracle_api, emu_api, name):
    sc = SCENARIOS[name]()
    Eo, Ee = sc.to_engine(oracle_api), sc.to_engine(emu_api)
    assert Eo.dta_enumerate_route_sets(5, 4242) == Ee.dta_enumerate_route_sets(5, 4242)
    for x, y in zip(Eo.dta_route_sets_csr(), Ee.dta_route_sets_csr()):
        assert np.array_equal(x, y)
    Eo.close(); Ee.close()


def check_swap_against_oracle(oracle_api, api, name, mode, **kw):
    Eo, Ee = (run(a, name, hard_deterministic_mode=True, **kw) for a in (oracle_api, api))
    assert np.array_equal(Eo.get_array(C.A_VEH_ARRIVAL_TS), Ee.get_array(C.A_VEH_ARRIVAL_TS))
    moved = 0
    for swap_prob, rng_seed in ((0.05, 7), (0.5, 8), (1.0, 9)):
        a, b = Eo.dta_route_swap(mode, swap_prob, rng_seed), Ee.dta_route_swap(mode, swap_prob, rng_seed)
        assert_swap_equal(a, b)
        for arr in (C.A_DTA_RA_ENTRY_T, C.A_DTA_T_GAP, C.A_DTA_CHOICE):
            assert np.array_equal(Eo.get_array(arr), Ee.get_array(arr)), arr
        moved += a["n_swap"]
    Eo.close(); Ee.close()
    return moved


@pytest.mark.parametrize("name", ["diverge", "grid5", "siouxfalls", "dta_grid9"])
@pytest.mark.parametrize("mode", [0, 1])
def test_emu_route_swap_matches_oracle(oracle_api, emu_api, name, mode):
    moved = check_swap_against_oracle(oracle_api, emu_api, name, mode)
    if name in ("siouxfalls", "dta_grid9"):
        assert moved > 0  # congested networks: some platoons do find a cheaper route


def check_batched_swap_equals_per_replica(api, mode):
    sc = SCENARIOS["dta_grid9"]()
    E = sc.to_engine(api, vehicle_logging_timestep_interval=-1, random_seed=5, num_replicas=3)
    E.set_option(C.OPT_RECORD_TRAVERSALS, 1)
    E.dta_enumerate_route_sets(5, 77)
    E.main_loop()
    single = [E.dta_route_swap(mode, 0.3, 100 + r, replica=r) for r in range(3)]
    gaps = [E.get_array(C.A_DTA_T_GAP, r) for r in range(3)]
    batched = E.dta_route_swap_all(mode, 0.3, 100)
    for r in range(3):
        assert_swap_equal(single[r], batched[r])
        assert np.array_equal(gaps[r], E.get_array(C.A_DTA_T_GAP, r))
    assert not np.array_equal(batched[0]["ra_link_ids"], batched[1]["ra_link_ids"])  # replicas are different runs
    E.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_emu_batched_swap_equals_per_replica(emu_api, mode):
    check_batched_swap_equals_per_replica(emu_api, mode)


# ---- aborted trips: a platoon whose enforced route ends in a dead end aborts; dta_solver.cpp:168 / 308 skip every
#      state other than END, so it keeps no route for the next day and consumes no draw of the swap stream ----------------------
def run_with_aborts(api, **kw):
    """dead_end scenario (links l1 orig->mid, l2 mid->dest, l3 mid->deadend): every 4th platoon bound for `dest` is forced onto
    l1, l3 and aborts at the dead end; the route store knows the pair (orig, dest), so a backend that treated aborts like
    finished trips would draw for them and re-enforce their route."""
    sc = S.dead_end()
    E = sc.to_engine(api, vehicle_logging_timestep_interval=1, hard_deterministic_mode=True, **kw)
    E.set_option(C.OPT_RECORD_TRAVERSALS, 1)
    E.dta_enumerate_route_sets(3, 5)
    dest = E.get_array(C.A_VEH_DEST)
    d_main = int(np.bincount(dest).argmax())
    ids, off, k = [], [0], 0
    for d in dest:
        if int(d) == d_main:
            ids.extend([0, 2] if k % 4 == 0 else [0, 1])
            k += 1
        else:
            ids.extend([0, 2])
        off.append(len(ids))
    E.batch_enforce_routes_csr(np.asarray(ids, np.int32), np.asarray(off, np.int64))
    E.main_loop()
    assert (E.get_array(C.A_VEH_STATE) == C.VS_ABORT).sum() >= 10
    return E


def check_swap_with_aborts(api_a, api_b, mode):
    Ea, Eb = run_with_aborts(api_a), run_with_aborts(api_b)
    assert np.array_equal(Ea.get_array(C.A_VEH_STATE), Eb.get_array(C.A_VEH_STATE))
    for swap_prob, rng_seed in ((0.3, 7), (1.0, 9)):
        a, b = Ea.dta_route_swap(mode, swap_prob, rng_seed), Eb.dta_route_swap(mode, swap_prob, rng_seed)
        assert_swap_equal(a, b)
    aborted = np.flatnonzero(Ea.get_array(C.A_VEH_STATE) == C.VS_ABORT)
    assert np.all(np.diff(a["rs_offsets"])[aborted] == 0)  # no route kept for an aborted platoon
    Ea.close(); Eb.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_route_swap_with_aborts_oracle_matches_reference(oracle_api, ref_api, mode):
    check_swap_with_aborts(oracle_api, ref_api, mode)


@pytest.mark.parametrize("mode", [0, 1])
def test_emu_route_swap_with_aborts(oracle_api, emu_api, mode):
    check_swap_with_aborts(oracle_api, emu_api, mode)
