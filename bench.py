#!/usr/bin/env python
"""bench.py -- platoon-updates/sec of the simulation loop (World.exec_simulation) on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload chicago|chicago_dn5|grid11|siouxfalls]
                    [--replicas R] [--impl ours|reference]

A "step" = one full run of the workload's simulation loop (all T timesteps) for a batch of R seeded replicas per
GPU.  `value` = RUN platoon-updates of all replicas on all GPUs / device time of main_loop (inputs resident in
HBM, CUDA events on the engine's launch stream, max over ranks).  `e2e` = the same metric through the C-ABI from
HOST arrays: create world + vectorised ingest + upload (H2D) + main_loop + read-back (D2H) inside the timed
region.  `--impl reference` times the reference's own C++ engine (oracle/_ref, built from the reference sources;
else the oracle port) on the host cores for the same workload / metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (builder kwargs, description)
    "chicago": ("chicago_sketch", dict(deltan=30, tmax=10000), "Chicago sketch, deltan=30, tmax=10000 (example_23en): 546 nodes, 2176 links, 32284 platoons = 968520 veh, 333 steps"),
    "chicago_dn5": ("chicago_sketch", dict(deltan=5, tmax=10000), "Chicago sketch, deltan=5: 211782 platoons = 1058910 veh, 2000 steps"),
    "grid11": ("grid", dict(n=11, deltan=5, tmax=7200), "11x11 grid (example_04en): 121 nodes, 440 links, 12100 platoons = 60500 veh, 1440 steps"),
    "siouxfalls": ("sioux_falls", dict(deltan=5, tmax=7200), "Sioux Falls: 24 nodes, 76 links, 6938 platoons, 1440 steps"),
    "grid100": ("grid_random_od", dict(n=100, n_platoons=1_000_000, deltan=5, tmax=7200), "100x100 grid, signals at every node, random OD: 10000 nodes, 39600 links, 1M platoons = 5M veh, 1440 steps"),
}
GRID100_CPU_SAMPLE_T = 300.0  # simulated seconds of C4 the CPU legs run (one route search + 60 steps: ~30 s on one core)
ALG_BYTES = {  # algorithmic bytes (SURVEY.md 8d / DESIGN.md)
    "per_platoon_update": 32, "per_link_step": 40, "per_node_step": 20, "per_transfer": 160,
}


def make_scenario(workload, seed):
    from uxsim_b200 import scenario as S

    fn, kw, _ = WORKLOADS[workload]
    return getattr(S, fn)(seed=seed, **kw)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def load_reference_lib():
    from uxsim_b200._capi import CApi

    ref = os.path.join(ROOT, "oracle", "_ref", "libuxsim_ref.so")
    if os.path.exists(ref):
        return CApi(ref, "uxr_"), "reference"
    ora = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(ora):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, capture_output=True)
    return CApi(ora, "uxo_"), "port"


def time_reference(workload, seeds, threads):
    """Run the reference C++ engine (or the oracle port) on the host cores; returns (updates, seconds)."""
    from uxsim_b200 import _capi as C

    api, kind = load_reference_lib()
    upd, secs = 0, 0.0
    for s in seeds:
        sc = make_scenario(workload, s)
        E = sc.engine_from_tables(api, sc.tables(), vehicle_logging_timestep_interval=-1, threads=threads)
        t0 = time.perf_counter()
        E.main_loop()
        secs += time.perf_counter() - t0
        upd += E.get_int(C.I_PLATOON_UPDATES)
        E.close()
    return upd, secs, kind


_GUARDED_CHILD = r"""
import json, os, resource, sys, time
sys.path.insert(0, sys.argv[1])
workload, seed, until_t, lib, mem_gb = sys.argv[2], int(sys.argv[3]), float(sys.argv[4]), sys.argv[5], int(sys.argv[6])
resource.setrlimit(resource.RLIMIT_AS, (mem_gb << 30, mem_gb << 30))  # fail with bad_alloc instead of meeting the OOM killer
import bench
from uxsim_b200 import _capi as C
api = C.CApi(os.path.join(sys.argv[1], "oracle", "_ref", "libuxsim_ref.so"), "uxr_") if lib == "reference" else C.CApi(os.path.join(sys.argv[1], "oracle", "liboracle.so"), "uxo_")
sc = bench.make_scenario(workload, seed)
E = sc.engine_from_tables(api, sc.tables(), vehicle_logging_timestep_interval=-1, threads=-1 if lib == "reference" else 1)
t0 = time.perf_counter()
E.main_loop(until_t=until_t)
print(json.dumps({"updates": E.get_int(C.I_PLATOON_UPDATES), "seconds": time.perf_counter() - t0}))
"""


def time_cpu_guarded(workload, seed, until_t):
    """CPU leg for scenarios that may not fit the host (C4): a child process with an address-space limit runs the first
    `until_t` simulated seconds on the reference C++ engine; if that fails (it does on C4: the engine keeps a route-preference
    vector over all links per vehicle, ~320 GB for 1M platoons x 39 600 links -> std::bad_alloc) the oracle port runs the same
    sample on one core.  Returns (updates, seconds, kind, note)."""
    try:
        mem_gb = max(8, min(48, int(os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") * 0.6) >> 30))
    except (ValueError, OSError):
        mem_gb = 32
    note = ""
    for lib in ("reference", "port"):
        if lib == "reference" and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libuxsim_ref.so")):
            continue
        r = subprocess.run([sys.executable, "-c", _GUARDED_CHILD, ROOT, workload, str(seed), str(until_t), lib, str(mem_gb)], capture_output=True, text=True)
        if r.returncode == 0 and r.stdout.strip():
            d = json.loads(r.stdout.strip().splitlines()[-1])
            return d["updates"], d["seconds"], lib, note
        err = (r.stderr.strip().splitlines() or ["killed"])[-1][:120]
        note += f"{lib} engine failed under a {mem_gb} GB address-space limit ({err}); "
    raise RuntimeError("no CPU engine could run the sample: " + note)


def _ref_worker(job):
    workload, seeds = job
    os.environ["OMP_NUM_THREADS"] = "1"
    upd, secs, kind = time_reference(workload, seeds, 1)
    return upd, secs, kind


def time_reference_ensemble(workload, seeds, procs):
    """The CPU's best arrangement for an ensemble: `procs` independent single-thread scenarios side by side (one process
    each, spawned so no CUDA state is inherited).  Returns (updates, seconds = the slowest worker's main_loop time, kind)."""
    import multiprocessing as mp

    jobs = [(workload, seeds[i::procs]) for i in range(procs) if seeds[i::procs]]
    with mp.get_context("spawn").Pool(len(jobs)) as pool:
        res = pool.map(_ref_worker, jobs)
    return sum(r[0] for r in res), max(r[1] for r in res), res[0][2]


_PYMODE_CHILD = r"""
import json, os, sys, time
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "oracle"))
os.environ["UXSIM_REFERENCE_ROOT"] = os.path.join(root, "baseline", "_ref")
import numpy as np
import bench
from ref_python_shim import import_reference  # test infrastructure: stub modules for matplotlib, nothing else
from uxsim_b200 import scenario as S
ux = import_reference()
workload, seed, log = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
sc = bench.make_scenario(workload, seed)
W = ux.World(**{**sc.world, "vehicle_logging_timestep_interval": log, "print_mode": 0, "save_mode": 0, "show_mode": 0})
S.replay_into(sc, W)
t0 = time.perf_counter()
W.exec_simulation()
secs = time.perf_counter() - t0
upd = int(round(sum(float((np.asarray(l.cum_arrival) - np.asarray(l.cum_departure)).sum()) for l in W.LINKS) / W.DELTAN))  # platoons on links, summed over steps
print(json.dumps({"workload": workload, "seconds": secs, "platoon_updates": upd, "logging": log > 0}))
"""


def time_python_mode(jobs, timeout=600):
    """The reference's pure-Python engine (uxsim/uxsim.py, staged UNMODIFIED under baseline/_ref by baseline/stage_reference.py),
    timed around exec_simulation() as example_28en_benchmark_cpp_mode.py:44-46 does; `jobs` = [(workload, logging interval)], one
    single-thread process each, side by side.  Returns {key: {...}} or {"unavailable": why}."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "uxsim")):
        return {"unavailable": "baseline/_ref/uxsim is not staged (python baseline/stage_reference.py needs /root/reference)"}
    procs = [(w, lg, subprocess.Popen([sys.executable, "-c", _PYMODE_CHILD, ROOT, w, "0", str(lg)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)) for w, lg in jobs]
    out = {"engine": "reference pure-Python engine (uxsim.py), one single-thread process per run, timing around exec_simulation() only", "cores_used": len(procs)}
    for w, lg, pr in procs:
        key = f"{w}{'_logging_on' if lg > 0 else ''}"
        try:
            so, se = pr.communicate(timeout=timeout)
            d = json.loads(so.strip().splitlines()[-1])
            out[key] = {"seconds": d["seconds"], "platoon_updates": d["platoon_updates"], "updates_per_sec": d["platoon_updates"] / d["seconds"], "vehicle_logging": lg > 0}
        except Exception as e:  # a missing dependency of the reference on this box must not take the bench line down
            pr.kill()
            out[key] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    return out


def time_reference_logging(workload, cores):
    """The reference C++ engine with vehicle logging ON (the API default; Vehicle::log_data traffi.cpp:1060-1110): one seed per
    core side by side, single-threaded each.  Returns updates/s."""
    import multiprocessing as mp

    n = max(1, min(cores, 16))
    with mp.get_context("spawn").Pool(n) as pool:
        res = pool.map(_ref_log_worker, [(workload, 4000 + k) for k in range(n)])
    return sum(r[0] for r in res) / max(r[1] for r in res), n


def _ref_log_worker(job):
    from uxsim_b200 import _capi as C

    workload, seed = job
    os.environ["OMP_NUM_THREADS"] = "1"
    api, _ = load_reference_lib()
    sc = make_scenario(workload, seed)
    E = sc.engine_from_tables(api, sc.tables(), vehicle_logging_timestep_interval=1, threads=1)
    t0 = time.perf_counter()
    E.main_loop()
    secs = time.perf_counter() - t0
    upd = E.get_int(C.I_PLATOON_UPDATES)
    E.close()
    return upd, secs


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = -1 if cores > 1 else 1
    if args.workload == "grid100":  # bounded sample in a guarded child (see time_cpu_guarded)
        ms, upd_total, t_total, kind, note = [], 0, 0.0, "port", ""
        for i in range(args.warmup + args.steps):
            u, s_, kind, note = time_cpu_guarded(args.workload, 1000 + i, GRID100_CPU_SAMPLE_T)
            if i >= args.warmup:
                ms.append(s_ * 1e3); upd_total += u; t_total += s_
        value = upd_total / t_total
        ncores = cores if kind == "reference" else 1
        sample = f"{note}first {GRID100_CPU_SAMPLE_T:.0f} s of simulated time (route search at t=0 + {int(GRID100_CPU_SAMPLE_T / 5)} steps) per step, " + ("reference C++ engine" if kind == "reference" else "oracle port (plain C, 1 thread)")
        print(json.dumps({"impl": "reference", "metric": "platoon_updates_per_sec", "value": value, "unit": "platoon-updates/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": {"workload": WORKLOADS[args.workload][2], "sample": sample},
                          "cpu_baseline": {"value": value, "unit": "platoon-updates/s", "cores": ncores, "kind": kind, "sample": sample},
                          "e2e": {"value": value, "unit": "platoon-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    # Two arrangements of the same engine on the same cores; the line reports the faster one:
    #  "omp"      one scenario at a time, OpenMP threads inside it (what cpp=True does for a single run)
    #  "ensemble" `cores` single-thread scenarios side by side (what a CPU user would do for a seed ensemble)
    per_core = {"chicago": 1, "chicago_dn5": 1, "grid11": 2, "siouxfalls": 4, "grid100": 1}.get(args.workload, 1)
    sample_seeds = cores * per_core if cores > 1 and args.workload != "grid100" else 1
    omp_seeds = 2 if args.workload == "chicago" else 1
    ms, upd_total, t_total = [], 0, 0.0
    omp_upd, omp_t = 0, 0.0
    kind = "reference"
    for i in range(args.warmup + args.steps):
        seeds = [1000 + i * sample_seeds + k for k in range(sample_seeds)]
        if sample_seeds > 1:
            upd, secs, kind = time_reference_ensemble(args.workload, seeds, cores)
        else:
            upd, secs, kind = time_reference(args.workload, seeds, threads)
        u2, s2, _ = time_reference(args.workload, seeds[:omp_seeds], threads)
        if i >= args.warmup:
            ms.append(secs * 1e3)
            upd_total += upd
            t_total += secs
            omp_upd += u2
            omp_t += s2
    value = upd_total / t_total
    arrangement = "ensemble" if sample_seeds > 1 else "omp"
    if omp_upd / omp_t > value:
        value, arrangement, ms = omp_upd / omp_t, "omp", [1e3 * omp_t / args.steps]
    line = {
        "impl": "reference", "metric": "platoon_updates_per_sec", "value": value, "unit": "platoon-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "fixture+synthetic-seeds",
        "config": {"workload": WORKLOADS[args.workload][2], "replicas_per_step": sample_seeds if arrangement == "ensemble" else omp_seeds, "arrangement": arrangement,
                   "omp_value": omp_upd / omp_t, "ensemble_value": upd_total / t_total if sample_seeds > 1 else None,
                   "engine": "reference trafficpp C++ (traffi.cpp via oracle/ref_harness.cpp)" if kind == "reference" else "oracle port (plain C, 1 thread)"},
        "cpu_baseline": {"value": value, "unit": "platoon-updates/s", "cores": cores if kind == "reference" else 1, "kind": kind,
                         "sample": (f"{sample_seeds} seeds per step as {cores} single-thread processes side by side" if arrangement == "ensemble" else f"{omp_seeds} seed(s) per step, one scenario at a time with {cores} OpenMP threads")
                                   + ", vehicle logging off, timing around main_loop only (slowest worker)"},
        "e2e": {"value": value, "unit": "platoon-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_extras:
        # the same engine with vehicle logging ON (the API default), and the reference's OTHER mode: the pure-Python engine
        if kind == "reference":
            v_log, n_log = time_reference_logging(args.workload, cores)
            line["logging_on"] = {"value": v_log, "unit": "platoon-updates/s", "sample": f"{n_log} seeds as {n_log} single-thread processes side by side, vehicle_log_mode = 1"}
        line["python_mode"] = time_python_mode([(args.workload, -1), (args.workload, 1)])
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="chicago", choices=sorted(WORKLOADS))
    ap.add_argument("--replicas", type=int, default=0, help="seeded replicas per GPU per step (0: the workload's default)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs (logging on, other workloads, DUE ensemble, Python mode)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from uxsim_b200 import _capi as C
    from uxsim_b200.ensemble import _CudaArrayView

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL may print its version banner on stdout; keep stdout for the single JSON line
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.all_reduce(torch.zeros(1, device="cuda"))
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    api = C.load_product()  # raises without the CUDA library: no fallback
    # replicas per GPU: throughput keeps growing with the ensemble (Chicago sketch: 1.20e9 platoon-updates/s at 128, 2.28e9 at 1024,
    # 2.46e9 at 2048 = 68 GB of state -- from 256 on the node step runs with a thread per node and replica); memory bounds the
    # fine-platoon workloads
    R = args.replicas or {"chicago": 2048, "chicago_dn5": 32, "grid11": 512, "siouxfalls": 512, "grid100": 1}[args.workload]
    sc = make_scenario(args.workload, 0)
    tab = sc.tables()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def build(step, **over):
        seed = 100000 * step + rank * R  # distinct seeds per step, rank, replica
        return sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1, num_replicas=R, device=local_rank, random_seed=seed, **over)

    def reduce_stats(E):
        """The path's only exchange step: all-reduce of per-link ensemble statistics [L,4] (NCCL, N>1 only)."""
        ptr = C.C.c_void_p(0)
        n = E._check(api.ensemble_link_stats_device(E.h, C.C.byref(ptr)))
        t = torch.as_tensor(_CudaArrayView(ptr.value, int(n), np.float64, E), device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return t

    def updates_of(E):
        return int(E.get_array_all(C.A_LINK_STAT_COUNT).sum())  # RUN platoon-updates of every replica (one gather + one transfer)

    # ------------------------------------------------------------------ value: device-timed, inputs resident
    sampler = ClockSampler(local_rank) if rank == 0 else None
    step_ms, upd_total, launches = [], 0, 0
    for i in range(args.warmup + args.steps):
        E = build(i)
        E.get_array(C.A_LINK_NUM_VEHICLES)  # forces allocation + upload: inputs are resident before timing
        flush_buf.zero_()
        barrier()
        if i == args.warmup and sampler:
            sampler.start()
        l0 = E.get_int(C.I_KERNEL_LAUNCHES)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        E.main_loop()  # CUDA graphs on the engine's stream; returns after the stream is drained
        ms = E.get_double(C.D_LAST_LOOP_MS)  # CUDA events on the launching stream, recorded inside the engine
        if world > 1:
            reduce_stats(E)
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            ms = max(ms, ev0.elapsed_time(ev1))  # includes the all-reduce
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        if i >= args.warmup:
            step_ms.append(ms)
            upd = updates_of(E)
            launches += E.get_int(C.I_KERNEL_LAUNCHES) - l0
            if world > 1:
                t = torch.tensor([upd], dtype=torch.float64, device="cuda")
                dist.all_reduce(t)
                upd = int(t.item())
            upd_total += upd
        E.close()
    clocks = sampler.stop() if sampler else None
    value = upd_total / (sum(step_ms) / 1e3)

    # ------------------------------------------------------------------ e2e: host arrays -> C-ABI -> host results
    e2e_s, e2e_upd, h2d, d2h = 0.0, 0, 0, 0
    pinned = {}
    for i in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        E = build(1000 + i)
        E.main_loop()
        stats = reduce_stats(E)
        host_stats = stats.cpu()
        if not pinned:  # page-locked result buffers, allocated once and reused by every step (what a production caller does)
            for what in (C.A_VEH_ARRIVAL_TS, C.A_VEH_TRAVEL_TIME, C.A_LINK_STAT_COUNT):
                probe = E.get_array(what, 0)
                pinned[what] = torch.empty((R, probe.size), dtype=torch.from_numpy(probe[:0].copy()).dtype, pin_memory=True).numpy()
        res = tuple(E.get_array_all(what, out=pinned[what]) for what in (C.A_VEH_ARRIVAL_TS, C.A_VEH_TRAVEL_TIME, C.A_LINK_STAT_COUNT))  # [R, n] each, one transfer each
        curves = (E.get_array(C.A_CUM_ARRIVAL, 0), E.get_array(C.A_CUM_DEPARTURE, 0))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if i >= args.warmup:
            upd = int(res[2].sum())
            if world > 1:
                t = torch.tensor([upd], dtype=torch.float64, device="cuda")
                dist.all_reduce(t)
                upd = int(t.item())
            e2e_s += dt
            e2e_upd += upd
            h2d = E.get_int(C.I_H2D_BYTES)
            d2h = E.get_int(C.I_D2H_BYTES) + host_stats.numel() * 8
        del curves
        E.close()
    e2e_value = e2e_upd / e2e_s

    # ------------------------------------------------------------------ roofline of the dominant kernel (rank 0)
    roofline, kernels, cpu_baseline, single = None, None, None, None
    if rank == 0:
        E = build(5000)
        E.set_option(C.OPT_KERNEL_TIMING, 1)  # eager launches with a CUDA-event pair around each kernel
        E.main_loop()
        T, L, N = E.get_int(C.I_TOTAL_TIMESTEPS), E.get_int(C.I_NUM_LINKS), E.get_int(C.I_NUM_NODES)
        upd = updates_of(E)
        events = E.get_int(C.I_LINK_EVENTS) * R
        kms = {C.K_NAMES[k]: E.get_double(C.D_KERNEL_MS_OF + k) for k in range(len(C.K_NAMES))}
        kl = {C.K_NAMES[k]: E.get_int(C.I_KERNEL_LAUNCHES_OF + k) for k in range(len(C.K_NAMES))}
        tot_ms = sum(kms.values())
        alg = {
            "k_update": ALG_BYTES["per_platoon_update"] * upd,
            "k_node": (ALG_BYTES["per_link_step"] * L + ALG_BYTES["per_node_step"] * N) * T * R + ALG_BYTES["per_transfer"] * events,
        }
        D = E.get_int(C.I_NUM_ACTIVE_DESTS)
        n_route = kl.get("k_sssp", 0)  # route updates in the run; SURVEY 8(d): 12*D*N bytes of dist/next out, 16*D*L for the DUO rows
        alg["k_sssp"] = 12 * D * N * R * n_route
        alg["k_duo"] = 16 * D * L * R * kl.get("k_duo", 0)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        kernels = {}
        for k in kms:
            if kl[k] == 0:
                continue
            avg_us = kms[k] / kl[k] * 1e3
            kernels[k] = {"share": kms[k] / tot_ms, "launches": kl[k], "avg_us": avg_us}
            if k in alg:
                per_launch = alg[k] / kl[k]
                kernels[k]["alg_bytes_per_launch"] = per_launch
                kernels[k]["GBps"] = per_launch / (avg_us * 1e-6) / 1e9
        dom = max((k for k in kernels if k in alg), key=lambda k: kernels[k]["share"])
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch from the committed ncu capture of this workload / replica count, if there is one
            ent = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}:{R}")
            if ent and dom in ent:
                traffic, traffic_src = ent[dom], ent.get("source")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["GBps"] / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)",
                    "alg_bytes_per_launch": kernels[dom]["alg_bytes_per_launch"], "avg_launch_us": kernels[dom]["avg_us"],
                    "note": "every kernel of this path is latency/launch-bound at this size (no tensor cores); see DESIGN.md"}
        E.close()
        # single-scenario latency (the "Chicago-sketch sim wall time" half of the metric)
        lat = []
        for i in range(3):
            E1 = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1, num_replicas=1, device=local_rank, random_seed=42)
            E1.get_array(C.A_LINK_NUM_VEHICLES)
            E1.main_loop()
            lat.append(E1.get_double(C.D_LAST_LOOP_MS))
            u1 = int(E1.get_array(C.A_LINK_STAT_COUNT, 0).sum())
            E1.close()
        single = {"scenario_wall_ms": float(np.median(lat)), "platoon_updates": u1, "updates_per_sec": u1 / (np.median(lat) / 1e3)}
        lat8 = []
        for i in range(3):  # ... and a small ensemble (8 replicas: what one GPU holds when 64 chains are spread over 8)
            E8 = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1, num_replicas=8, device=local_rank, random_seed=42)
            E8.get_array(C.A_LINK_NUM_VEHICLES)
            E8.main_loop()
            lat8.append(E8.get_double(C.D_LAST_LOOP_MS))
            u8 = int(E8.get_array_all(C.A_LINK_STAT_COUNT).sum())
            E8.close()
        single["replicas_8"] = {"wall_ms": float(np.median(lat8)), "updates_per_sec": u8 / (np.median(lat8) / 1e3)}
        # CPU baseline on this box's host cores: bounded sample of the same workload
        if world == 1 and args.workload == "grid100":
            u_cpu, t_cpu, kind, note = time_cpu_guarded(args.workload, 2000, GRID100_CPU_SAMPLE_T)
            cpu_baseline = {"value": u_cpu / t_cpu, "unit": "platoon-updates/s", "cores": (os.cpu_count() or 1) if kind == "reference" else 1, "kind": kind,
                            "sample": f"{note}first {GRID100_CPU_SAMPLE_T:.0f} s of simulated time (route search at t=0 + {int(GRID100_CPU_SAMPLE_T / 5)} steps), {t_cpu:.1f} s, "
                                      + ("reference C++ engine" if kind == "reference" else "oracle port (plain C, 1 thread)")}
        elif world == 1:
            cores = os.cpu_count() or 1
            t_cpu, u_cpu, n_seed, kind = 0.0, 0, 0, "reference"
            while t_cpu < args.cpu_seconds / 2 and n_seed < 64:  # one scenario at a time, OpenMP inside it
                u, s_, kind = time_reference(args.workload, [2000 + n_seed], -1 if cores > 1 else 1)
                u_cpu += u
                t_cpu += s_
                n_seed += 1
            omp_v, best_v = u_cpu / t_cpu, u_cpu / t_cpu
            sample = f"{n_seed} seeds of the workload, one scenario at a time with {cores} OpenMP threads, {t_cpu:.1f} s of main_loop, vehicle logging off"
            ens_v = None
            if cores > 1 and kind == "reference" and args.workload != "grid100":
                per_proc = max(1, int(args.cpu_seconds / 2 / max(1e-3, (t_cpu / n_seed) * 4.5)))  # a single thread is ~4.5x slower
                u_e, t_e, _ = time_reference_ensemble(args.workload, [3000 + k for k in range(cores * min(per_proc, 8))], cores)
                ens_v = u_e / t_e
                if ens_v > best_v:
                    best_v = ens_v
                    sample = (f"{cores * min(per_proc, 8)} seeds of the workload as {cores} single-thread processes side by side, {t_e:.1f} s (slowest worker's "
                              "main_loop time), vehicle logging off")
            cpu_baseline = {"value": best_v, "unit": "platoon-updates/s", "cores": cores if kind == "reference" else 1, "kind": kind,
                            "sample": sample, "omp_value": omp_v, "ensemble_value": ens_v, "scenario_wall_ms": 1e3 * t_cpu / n_seed}

    # ------------------------------------------------------------------ extra legs (BASELINE.md section 3 / VERDICT r1)
    logging_on, other, due, python_mode = None, None, None, None
    if not args.no_extras:
        flush_buf = None
        torch.cuda.empty_cache()
        if rank == 0 and args.workload != "grid100":
            # (1) vehicle logging ON (vehicle_logging_timestep_interval=1, the API default): every RUN platoon-update also appends a
            #     36-byte record (Vehicle::log_data traffi.cpp:1060-1110).  Bounded replica count: the log arena is ~350 MB per replica.
            R_log = min(R, 128 if args.workload == "chicago" else 16 if args.workload == "chicago_dn5" else 64)
            res = {}
            for tag, interval in (("off", -1), ("on", 1)):
                ms_l, upd_l = [], 0
                for i in range(3):
                    El = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=interval, num_replicas=R_log, device=local_rank, random_seed=7000 + i)
                    El.get_array(C.A_LINK_NUM_VEHICLES)
                    El.main_loop()
                    if i > 0:
                        ms_l.append(El.get_double(C.D_LAST_LOOP_MS))
                        upd_l += int(El.get_array_all(C.A_LINK_STAT_COUNT).sum())
                    if tag == "on" and i == 2:  # hand-back of ONE replica's logs in the reference's compact flat layout (what the analyzer asks for)
                        t0 = time.perf_counter()
                        n_rec = El.get_array(C.A_LOG_T, 0).size
                        res["handback_ms"], res["log_entries"] = 1e3 * (time.perf_counter() - t0), int(n_rec)
                    El.close()
                res[tag] = upd_l / (sum(ms_l) / 1e3)
            logging_on = {"value": res["on"], "unit": "platoon-updates/s", "replicas": R_log, "logging_off_same_replicas": res["off"], "on_over_off_time": res["off"] / res["on"],
                          "handback_one_replica_ms": res["handback_ms"], "log_entries_one_replica": res["log_entries"],
                          "note": "device loop with the per-step RUN records written; hand-back = scatter by platoon + D2H + host assembly of log_t/x/v/s/lane/link/state for one replica"}
        if rank == 0:
            # (2) the other named sizes, device-timed like `value`: C3' (Chicago at deltan=5, 1.06 M vehicles) and C4 (100x100 grid, 5 M vehicles)
            other = {}
            for wl, Rw in (("chicago_dn5", 16), ("grid100", 1)):
                if wl == args.workload:
                    continue
                scw = make_scenario(wl, 0)
                tabw = scw.tables()
                ms_w, upd_w = [], 0
                for i in range(2):
                    Ew = scw.engine_from_tables(api, tabw, vehicle_logging_timestep_interval=-1, num_replicas=Rw, device=local_rank, random_seed=9000 + i)
                    Ew.get_array(C.A_LINK_NUM_VEHICLES)
                    Ew.main_loop()
                    if i > 0:
                        ms_w.append(Ew.get_double(C.D_LAST_LOOP_MS))
                        upd_w += int(Ew.get_array_all(C.A_LINK_STAT_COUNT).sum())
                    Ew.close()
                other[wl] = {"workload": WORKLOADS[wl][2], "replicas": Rw, "ms_per_step": float(np.mean(ms_w)), "value": upd_w / (sum(ms_w) / 1e3), "unit": "platoon-updates/s",
                             "scenario_wall_ms": float(np.mean(ms_w)) if Rw == 1 else None}
                del scw, tabw
        if args.workload in ("chicago", "grid11", "siouxfalls"):
            # (3) BASELINE configs[4]: day-to-day DUE (route swap between simulated days) as seeded chains sharded over the ranks; ALL ranks.
            #     "fixed": 64 chains in total (BASELINE's wording: strong scaling); "weak": 64 chains per GPU.  One route-set enumeration serves both.
            from uxsim_b200.dta import DayToDayEnsemble

            days, k_routes = 4, 4
            E0 = sc.engine_from_tables(api, tab, vehicle_logging_timestep_interval=-1, device=local_rank)
            t0 = time.perf_counter()
            E0.dta_enumerate_route_sets(k_routes, 12345)
            enum_s = time.perf_counter() - t0
            due = {"config": {"mode": "DUE", "days": days, "routes_per_od": k_routes, "swap_prob": 0.05, "workload": args.workload}, "route_enumeration_s": enum_s,
                   "unit": "chain-days/s (days 1.. : rewind + simulation + route swap; day 0 builds the world and the graphs)"}
            for tag, chains in (("fixed_64_chains", 64), ("weak_64_per_gpu", 64 * world)):
                if chains < world:
                    continue
                ens = DayToDayEnsemble(sc, chains, mode="DUE", device=local_rank)
                barrier()
                ens.solve(max_iter=days, n_routes_per_od=k_routes, swap_prob=0.05, route_sets_from=E0)
                torch.cuda.synchronize()
                ds = np.array(ens.day_seconds)
                steady = float(ds[1:].sum())
                if world > 1:
                    t = torch.tensor([steady], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    steady = float(t.item())
                due[tag] = {"chains": chains, "chains_per_gpu": len(ens.local), "value": chains * (days - 1) / steady,
                            "day_ms": {"rewind": float(ds[1:, 0].mean() * 1e3), "simulation": float(ds[1:, 1].mean() * 1e3), "route_swaps": float(ds[1:, 2].mean() * 1e3)},
                            "link_volume_mean": float(ens.link_stats_sum[:, 0].sum() / chains)}
                del ens
            if rank == 0:  # the same day in the reference C++ (traffi.cpp + dta_solver.cpp via oracle/_ref) on one host core, logs on (it reads routes from them)
                try:
                    ref, kind = load_reference_lib()
                    if kind == "reference":
                        store = E0.dta_route_sets_csr()
                        Er = sc.engine_from_tables(ref, tab, vehicle_logging_timestep_interval=1, threads=1)
                        a = time.perf_counter(); Er.dta_set_route_sets(*store); Er.main_loop(); Er.dta_route_swap(0, 0.05, 1); b = time.perf_counter()
                        Er.close()
                        cores = os.cpu_count() or 1
                        due["reference_cpu"] = {"kind": kind, "one_core_day_ms": (b - a) * 1e3, "cores": cores, "chain_days_per_sec_all_cores": cores / (b - a)}
                except Exception as e:
                    due["reference_cpu"] = {"unavailable": str(e)[:200]}
            E0.close()
        if rank == 0 and world == 1:
            # (4) the reference's pure-Python mode on this box's host cores (C1, C2, C3; logging off, and on for the benched workload)
            jobs = [("siouxfalls", -1), ("grid11", -1), ("chicago", -1)]
            if args.workload in ("chicago", "grid11", "siouxfalls"):
                jobs.append((args.workload, 1))
            python_mode = time_python_mode(jobs)

    if rank == 0:
        line = {
            "metric": "platoon_updates_per_sec", "value": value, "unit": "platoon-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(step_ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "fixture network+demand (converted from the reference's dat/), synthetic seeds",
            "config": {"workload": WORKLOADS[args.workload][2], "replicas_per_gpu": R, "global_replicas": R * world,
                       "l2_flush": "256 MiB write between timed steps; every step builds a new world (state re-initialised; device slabs recycled through the engine's cache)",
                       "timing": "CUDA events on the engine's launch stream around all graph launches of main_loop; max over ranks",
                       "e2e_readback": "per replica: arrival step + travel time per platoon, per-link RUN counts (into page-locked buffers reused across steps); ensemble link stats [L,4]; cumulative curves of replica 0",
                       "parallelism": f"replicas x{R} per GPU (blockIdx.y), ranks x{world}, NCCL all-reduce of link stats only"},
            "e2e": {"value": e2e_value, "unit": "platoon-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels, "single_scenario": single,
            "cpu_baseline": cpu_baseline, "logging_on": logging_on, "workloads": other, "due_ensemble": due, "python_mode": python_mode,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
