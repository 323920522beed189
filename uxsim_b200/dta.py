"""Day-to-day DTA solvers over the CUDA engine.

Counterparts of ``CppSolverDUE`` / ``CppSolverDSO_D2D`` (DTAsolvers/DTAsolvers_cpp_adapter.py:249-493): same
``solve()`` signatures, the same convergence traces (``ttts, n_swaps, potential_swaps, t_gaps``), the same use of the
global ``random`` module for the enumeration / swap seeds (so a seeded script draws the same seeds in both), and the
same native calls in the same order -- enumerate route sets once, then per day: enforce yesterday's routes, simulate,
swap routes.  Here the simulation and the swap run on the GPU (``ux_main_loop``, ``ux_dta_route_swap``); the route
logs stay in the flat id-CSR form the engine returns and are converted to names only when asked for.
"""
from __future__ import annotations

import copy
import random
import time
import warnings

import numpy as np

from . import _capi as C


class _RouteLog:
    """Per-iteration traversed routes, ``{vehicle name: [link name, ...]}``, materialised on first access."""

    def __init__(self, link_ids, offsets, veh_names, link_names):
        self._csr, self._veh_names, self._link_names, self._d = (link_ids, offsets), veh_names, link_names, None

    def _build(self):
        if self._d is None:
            ids, off = self._csr
            ln = self._link_names
            self._d = {v: [ln[i] for i in ids[off[k]:off[k + 1]]] for k, v in enumerate(self._veh_names)}
        return self._d

    def __getitem__(self, k):
        return self._build()[k]

    def __iter__(self):
        return iter(self._build())

    def __len__(self):
        return len(self._veh_names)

    def items(self):
        return self._build().items()

    def keys(self):
        return self._build().keys()

    def values(self):
        return self._build().values()


class _CostLog:
    """One day's actual route costs: the array in platoon order, also addressable by vehicle name (the reference keys a dict by name)."""

    def __init__(self, values, veh_names):
        self.values_array, self._names, self._idx = values, veh_names, None

    def __getitem__(self, k):
        if isinstance(k, str):
            if self._idx is None:
                self._idx = {n: i for i, n in enumerate(self._names)}
            k = self._idx[k]
        return self.values_array[k]

    def __len__(self):
        return len(self.values_array)

    def __iter__(self):
        return iter(self._names)

    def keys(self):
        return list(self._names)

    def values(self):
        return list(self.values_array)

    def items(self):
        return list(zip(self._names, self.values_array))

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values_array, dtype=dtype)


class _LazyODRoutes(dict):
    """``{(o_name, d_name): [[link_name, ...], ...]}`` built from the store's CSR on first access
    (DTAsolvers_cpp_adapter.py:120-170 does the same for the C++ store)."""

    def __init__(self, csr, node_names, link_names):
        super().__init__()
        self._csr, self._nn, self._ln, self._built = csr, node_names, link_names, False

    def _build(self):
        if not self._built:
            o, d, od_off, r_off, links = self._csr
            for p in range(len(o)):
                super().__setitem__((self._nn[o[p]], self._nn[d[p]]),
                                    [[self._ln[i] for i in links[r_off[r]:r_off[r + 1]]] for r in range(od_off[p], od_off[p + 1])])
            self._built = True

    def __getitem__(self, k):
        self._build()
        return super().__getitem__(k)

    def __iter__(self):
        self._build()
        return super().__iter__()

    def __len__(self):
        self._build()
        return super().__len__()

    def items(self):
        self._build()
        return super().items()

    def keys(self):
        self._build()
        return super().keys()

    def values(self):
        self._build()
        return super().values()


def _ensure_cuda_world(W, solver_name):
    if not hasattr(W, "_E") or W._E.api.prefix != "ux_":
        raise ValueError(f"{solver_name} requires func_World to return World(cuda=True, ...); got {type(W).__name__}.")


def _route_sets_to_csr(W, route_sets):
    """User-supplied ``{(o, d): [[link, ...], ...]}`` (names or objects) -> the store's CSR, pairs sorted by (orig id, dest id)."""
    rows = sorted(((W.get_node(k[0]).id, W.get_node(k[1]).id), [[W.get_link(l).id for l in r] for r in routes]) for k, routes in route_sets.items())
    o, d, od_off, r_off, links = [], [], [0], [0], []
    for (a, b), routes in rows:
        o.append(a), d.append(b)
        for r in routes:
            links.extend(r)
            r_off.append(len(links))
        od_off.append(len(r_off) - 1)
    return o, d, od_off, r_off, links


class _CudaSolverBase:
    _is_cuda_solver = True
    _mode = 0      # 0 = DUE (travel time + tolls), 1 = DSO (travel time + congestion externality)
    _label = "DUE"

    def __init__(self, func_World, **kwargs):
        self.func_World = func_World
        self.W_sol = None
        self.W_intermid_solution = None
        self.dfs_link = []
        self.iteration_seconds = []   # per day: (simulation, route swap) wall seconds
        self.swap_device_ms = []      # per day: device time of the route-swap kernels

    def _solve(self, max_iter, n_routes_per_od, swap_prob, swap_num, route_sets, print_progress):
        s = self
        s.start_time = time.time()
        W_orig = s.func_World()
        _ensure_cuda_world(W_orig, type(s).__name__)
        if print_progress:
            W_orig.print_scenario_stats()
        if not W_orig.finalized:
            W_orig.finalize_scenario()
        node_names = [n.name for n in W_orig.NODES]
        link_names = [l.name for l in W_orig.LINKS]
        # same global-RNG consumption order as the reference solvers (one draw for the enumeration, one per day)
        enum_seed = random.randint(0, 2**31 - 1)
        if route_sets is not None:
            W_orig._E.dta_set_route_sets(*_route_sets_to_csr(W_orig, route_sets))
        else:
            W_orig._E.dta_enumerate_route_sets(n_routes_per_od, enum_seed)
        store = W_orig._E.dta_route_sets_csr()  # W_orig keeps the one store of this solve; every day's world shares it
        dict_od_to_routes = _LazyODRoutes(store, node_names, link_names)
        if print_progress:
            print(f"number of OD pairs: {len(store[0])}, number of routes: {len(store[3]) - 1}")
        s.ttts, s.n_swaps, s.potential_swaps, s.t_gaps, s.route_log, s.cost_log = [], [], [], [], [], []
        prev_rs = None
        prev_W = None
        best_avg_tt = None
        print(f"solving {s._label}...")
        for i in range(max_iter):
            W = s.func_World()
            _ensure_cuda_world(W, type(s).__name__)
            W._skip_log_on_terminate = i != max_iter - 1
            if not W.finalized:
                W.finalize_scenario()
            E = W._E
            E.set_option(C.OPT_RECORD_TRAVERSALS, 1)
            E.dta_share_route_sets(W_orig._E)
            if prev_rs is not None:
                E.batch_enforce_routes_csr(*prev_rs)
            t0 = time.perf_counter()
            W.exec_simulation()
            t1 = time.perf_counter()
            W.analyzer.print_simple_stats()
            unfinished = W.analyzer.trip_all - W.analyzer.trip_completed
            if unfinished > 0:
                warnings.warn(f"Warning: {unfinished} / {W.analyzer.trip_all} vehicles have not finished their trips. The {s._label} solver assumes "
                              "that all vehicles finish their trips during the simulation duration. Consider increasing the simulation time limit "
                              "or checking the network configuration.", UserWarning)
            W.dict_od_to_routes = dict_od_to_routes
            if s._mode == 0:
                s.W_intermid_solution = W
            elif unfinished == 0:
                avg_tt = W.analyzer.average_travel_time
                if s.W_intermid_solution is None or avg_tt < best_avg_tt:
                    best_avg_tt = avg_tt
                    W.analyzer = copy.copy(W.analyzer)
                    if s.W_intermid_solution is not None:
                        s.W_intermid_solution._E.close()  # the superseded best day gives its device slabs back now, not at GC
                    s.W_intermid_solution = W
            s.dfs_link.append(W.analyzer.link_to_pandas() if hasattr(W.analyzer, "link_to_pandas") else W.analyzer.link_summary())
            rng_seed = random.randint(0, 2**31 - 1)
            t2 = time.perf_counter()
            res = E.dta_route_swap(s._mode, swap_prob, rng_seed, swap_num=-1 if swap_num is None else swap_num,
                                   has_external_route_sets=route_sets is not None)
            t3 = time.perf_counter()
            s.iteration_seconds.append((t1 - t0, t3 - t2))
            s.swap_device_ms.append(E.get_double(C.D_DTA_SWAP_MS))
            prev_rs = (res["rs_link_ids"], res["rs_offsets"])
            veh_names = W._veh_by_index.names()
            t_gap_per_vehicle = res["total_t_gap"] / len(W.VEHICLES)
            if print_progress:
                gap = "time gap" if s._mode == 0 else "marginal time gap"
                print(f" iter {i}: {gap}: {t_gap_per_vehicle:.1f}, potential route change: {res['potential_n_swap']}, route change: {res['n_swap']}, "
                      f"total travel time: {W.analyzer.total_travel_time: .1f}, delay ratio: {W.analyzer.average_delay / W.analyzer.average_travel_time: .3f}")
            s.route_log.append(_RouteLog(res["ra_link_ids"], res["ra_offsets"], veh_names, link_names))
            s.cost_log.append(_CostLog(res["cost_actual"], veh_names))  # per platoon in creation order, and by vehicle name as in the reference
            s.ttts.append(int(W.analyzer.total_travel_time))
            s.n_swaps.append(res["n_swap"])
            s.potential_swaps.append(res["potential_n_swap"])
            s.t_gaps.append(t_gap_per_vehicle)
            if prev_W is not None and prev_W is not s.W_intermid_solution:
                prev_W._E.close()   # device slabs go back to the engine's cache for the next day
            prev_W = W
        s.end_time = time.time()
        print(f"{s._label} summary:")
        last_iters = max(1, int(max_iter / 4))
        if s._mode == 0:
            print(f" total travel time: initial {s.ttts[0]:.1f} -> average of last {last_iters} iters {np.average(s.ttts[-last_iters:]):.1f}")
            print(f" number of potential route changes: initial {s.potential_swaps[0]:.1f} -> average of last {last_iters} iters {np.average(s.potential_swaps[-last_iters:]):.1f}")
            print(f" route travel cost gap: initial {s.t_gaps[0]:.1f} -> average of last {last_iters} iters {np.average(s.t_gaps[-last_iters:]):.1f}")
        else:
            print(f" total travel time: initial {s.ttts[0]:.1f} -> minimum {np.min(s.ttts):.1f}")
            print(f" number of potential route changes: initial {s.potential_swaps[0]:.1f} -> minimum {np.min(s.potential_swaps):.1f}")
            print(f" route marginal travel time gap: initial {s.t_gaps[0]:.1f} -> minimum {np.min(s.t_gaps):.1f}")
        print(f" computation time: {s.end_time - s.start_time:.1f} seconds")
        s.W_sol = W if s._mode == 0 else (s.W_intermid_solution or W)
        return s.W_sol


    # ---- the three convergence plots of the reference's solvers (DTAsolvers.py:267-372); drawing is outside the loop (out of
    #      scope), so they are thin: without matplotlib they warn and return, like the reference's catch_exceptions_and_warn helpers
    def _plot(self, draw):
        try:
            import matplotlib.pyplot as plt

            draw(plt)
            plt.show()
        except Exception as e:  # noqa: BLE001
            warnings.warn(f"{type(self).__name__}: plotting skipped ({type(e).__name__}: {e})")

    def plot_convergence(self):
        def draw(plt):
            for title, ys in (("total travel time", self.ttts), ("number of route change", self.n_swaps),
                              ("travel cost difference between chosen route and minimum cost route", self.t_gaps)):
                plt.figure(figsize=(6, 2)); plt.title(title); plt.plot(ys); plt.xlabel("iter")
        self._plot(draw)

    def plot_link_stats(self):
        def draw(plt):
            for title, col, ylabel in (("traffic volume", "traffic_volume", "volume (veh)"), ("average travel time", "average_travel_time", "time (s)")):
                plt.figure(); plt.title(title)
                for i, name in enumerate(self.dfs_link[0]["link"]):
                    plt.plot([df[col][i] for df in self.dfs_link], label=name)
                plt.xlabel("iteration"); plt.ylabel(ylabel); plt.legend(loc="center left", bbox_to_anchor=(1, 0.5))
        self._plot(draw)

    def plot_vehicle_stats(self, orig=None, dest=None):
        def draw(plt):
            half = range(len(self.cost_log) // 2, len(self.cost_log))
            sel = [v for v in self.W_sol.VEHICLES.values() if orig in (None, v.orig.name) and dest in (None, v.dest.name)]
            costs = np.array([[self.cost_log[day][v.id] for day in half] for v in sel])
            plt.figure(); plt.title(f"orig: {orig or 'any'}, dest: {dest or 'any'}")
            plt.errorbar(x=[v.departure_time_in_second for v in sel], y=costs.mean(axis=1), yerr=costs.std(axis=1), fmt="bx", ecolor="#aaaaff", capsize=0)
            plt.xlabel("departure time of vehicle"); plt.ylabel("travel time")
        self._plot(draw)


class CudaSolverDUE(_CudaSolverBase):
    """Dynamic user equilibrium by day-to-day route swapping on the GPU (CppSolverDUE, DTAsolvers_cpp_adapter.py:249-366)."""

    _mode, _label = 0, "DUE"

    def solve(self, max_iter, n_routes_per_od=10, swap_prob=0.05, route_sets=None, print_progress=True):
        return self._solve(max_iter, n_routes_per_od, swap_prob, None, route_sets, print_progress)


class CudaSolverDSO_D2D(_CudaSolverBase):
    """Dynamic system optimum by day-to-day marginal-cost route swapping on the GPU (CppSolverDSO_D2D, :369-493)."""

    _mode, _label = 1, "DSO"

    def solve(self, max_iter, n_routes_per_od=10, swap_prob=0.05, swap_num=None, route_sets=None, print_progress=True):
        return self._solve(max_iter, n_routes_per_od, swap_prob, swap_num, route_sets, print_progress)


def SolverDUE(func_World, cuda=True, **kwargs):
    """``SolverDUE(func_World, cpp=True)`` selects the C++ solver in the reference (DTAsolvers.py); ``cuda=True`` here."""
    if not cuda:
        raise NotImplementedError("uxsim_b200 only provides the CUDA solvers; use the reference package for the Python / C++ ones.")
    return CudaSolverDUE(func_World, **kwargs)


def SolverDSO_D2D(func_World, cuda=True, **kwargs):
    if not cuda:
        raise NotImplementedError("uxsim_b200 only provides the CUDA solvers; use the reference package for the Python / C++ ones.")
    return CudaSolverDSO_D2D(func_World, **kwargs)


class DayToDayEnsemble:
    """R independent day-to-day chains advanced together (seeded replicas of a DUE / DSO solve).

    Every day is ONE world with R replicas (replica r = seed ``base_seed + r`` with its own enforced routes): one
    ``main_loop`` simulates all chains' day, then each chain swaps routes on its own finished run.  Works on the arrays of a
    :class:`uxsim_b200.scenario.Scenario` (no per-vehicle Python objects).  Ranks of a ``torch.distributed`` job give each
    rank its own block of chains (``uxsim_b200.ensemble.shard_seeds``); there is no exchange between chains.
    """

    def __init__(self, scenario, n_replicas, mode="DUE", base_seed=0, device=-1, api=None, **world_over):
        from .ensemble import shard_seeds

        self.n_global = int(n_replicas)
        self.rank, self.world_size = 0, 1
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                self.rank, self.world_size = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass
        self.local = shard_seeds(self.n_global, self.rank, self.world_size)   # this rank's block of chains
        if len(self.local) == 0:
            raise ValueError(f"rank {self.rank} of {self.world_size} has no chain: use at least as many chains as ranks")
        self.scenario, self.R, self.mode = scenario, len(self.local), {"DUE": 0, "DSO": 1}[mode]
        self.base_seed, self.device, self.api, self.world_over = int(base_seed) + self.local.start, device, api or C.load_product(), world_over
        self.link_stats_sum = None  # [L, 4] summed over the chains of ALL ranks after the last day (one all-reduce)
        self.ttts = self.n_swaps = self.potential_swaps = self.t_gaps = self.trips = None
        self.day_seconds = []      # per day: (build or rewind, simulation, route swaps) wall seconds
        self.swap_device_ms = []   # per day: device time of the batched route-swap kernels
        self.routes = None         # per chain: (link ids, offsets) enforced on the next day

    def solve(self, max_iter, n_routes_per_od=10, swap_prob=0.05, enum_seed=12345, swap_seed=0, keep_last_world=False, route_sets_from=None):
        """``route_sets_from``: an engine world that already holds the OD route sets (one enumeration can serve several
        ensembles of the same scenario, as one ``ODRouteSetStore`` serves every day of a reference solve); it is left open."""
        sc, api, R = self.scenario, self.api, self.R
        tab = sc.tables()
        if route_sets_from is not None:
            E0 = route_sets_from
        else:
            E0 = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1, device=self.device, **self.world_over)
            E0.dta_enumerate_route_sets(n_routes_per_od, enum_seed)
        dn = sc.deltan
        shape = (R, max_iter)
        self.ttts, self.n_swaps, self.potential_swaps, self.t_gaps, self.trips = (np.zeros(shape) for _ in range(5))
        E = None
        for day in range(max_iter):
            t0 = time.perf_counter()
            if E is None:   # day 0: build the world once; later days rewind it in place (routes stay on the device)
                E = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1, num_replicas=R, device=self.device,
                                          random_seed=self.base_seed, **self.world_over)
                E.set_option(C.OPT_RECORD_TRAVERSALS, 1)
                E.dta_share_route_sets(E0)
            else:
                E.dta_next_day()
            t1 = time.perf_counter()
            E.main_loop()
            t2 = time.perf_counter()
            day_seed = ((swap_seed * 1000003 + day * 65537) & 0x3FFFFFFF) + self.local.start   # chain c draws its swap candidates from seed + c
            results = E.dta_route_swap_all(self.mode, swap_prob, day_seed, routes=False)   # every chain's swap in the same launches
            tt_all = E.get_array_all(C.A_VEH_TRAVEL_TIME)
            for r, res in enumerate(results):
                tt = tt_all[r]
                done = tt >= 0
                self.ttts[r, day] = float(tt[done].sum()) * dn
                self.trips[r, day] = int(done.sum()) * dn
                self.n_swaps[r, day], self.potential_swaps[r, day] = res["n_swap"], res["potential_n_swap"]
                self.t_gaps[r, day] = res["total_t_gap"] / max(1, len(tt))
            t3 = time.perf_counter()
            self.day_seconds.append((t1 - t0, t2 - t1, t3 - t2))
            self.swap_device_ms.append(E.get_double(C.D_DTA_SWAP_MS))
        self.routes = [(E.get_array(C.A_DTA_RS_LINKS, r), E.get_array(C.A_DTA_RS_OFFSETS, r)) for r in range(R)]  # the solution: handed back once
        self.link_stats_sum = self._reduce_link_stats(E)
        if not keep_last_world:
            E.close()
            E = None
        if route_sets_from is None:
            E0.close()
        return E if keep_last_world else None

    def _reduce_link_stats(self, E):
        """Per-link {volume, sum tt, sum tt^2, platoons left} of the last day, summed over every chain of every rank: the
        path's only exchange (NCCL all-reduce on the device buffer; gloo / no process group: host tensor)."""
        import torch

        from .ensemble import _CudaArrayView, allreduce_link_stats

        L = E.get_int(C.I_NUM_LINKS)
        if self.api.prefix == "ux_" and torch.cuda.is_available():
            ptr = C.C.c_void_p(0)
            n = E._check(self.api.ensemble_link_stats_device(E.h, C.C.byref(ptr)))
            t = torch.as_tensor(_CudaArrayView(ptr.value, int(n), np.float64, E), device="cuda").clone()
        else:
            t = torch.from_numpy(E.ensemble_link_stats().reshape(-1).copy())
        return allreduce_link_stats(t).cpu().numpy().reshape(L, 4)

    def gather_traces(self):
        """Total-travel-time and gap traces of all chains of all ranks, ``[n_global, days]`` (rank order = chain order)."""
        if self.world_size == 1:
            return self.ttts, self.t_gaps
        import torch.distributed as dist

        parts = [None] * self.world_size
        dist.all_gather_object(parts, (self.ttts, self.t_gaps))
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
