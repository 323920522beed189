#!/usr/bin/env python
"""Supplementary benchmark for BASELINE.json configs[4]: day-to-day DUE / DSO solves as seeded replica chains sharded over
the GPUs of one node (one process per GPU; the only exchange is one NCCL all-reduce of the per-link statistics at the end).

  python bench_dta.py [--workload chicago|dta_grid9] [--chains 64] [--days 6] [--k 4] [--mode DUE|DSO]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_dta.py --gpus N ...

Prints one JSON line on rank 0.  `chain_days_per_sec` = chains x (days - 1) / wall time of days 1.. (day 0 builds the world and
the graphs); the reference figure next to it is one day of the same loop in the reference C++ (oracle/_ref: traffi.cpp +
dta_solver.cpp) on one host core, scaled by the core count (chains are independent, so cores run them side by side).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--workload", default="chicago", choices=["chicago", "dta_grid9"])
    ap.add_argument("--chains", type=int, default=64)
    ap.add_argument("--days", type=int, default=6)
    ap.add_argument("--k", type=int, default=4, help="routes per OD pair")
    ap.add_argument("--mode", default="DUE", choices=["DUE", "DSO"])
    ap.add_argument("--swap-prob", type=float, default=0.05)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import bench
    from uxsim_b200 import _capi as C
    from uxsim_b200 import scenario as S
    from uxsim_b200.dta import DayToDayEnsemble

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        saved = os.dup(1)
        os.dup2(2, 1)  # NCCL prints its banner on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.all_reduce(torch.zeros(1, device="cuda"))
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    sc = S.dta_grid9() if args.workload == "dta_grid9" else bench.make_scenario(args.workload, 0)
    ens = DayToDayEnsemble(sc, args.chains, mode=args.mode, device=local_rank)
    t0 = time.perf_counter()
    ens.solve(max_iter=args.days, n_routes_per_od=args.k, swap_prob=args.swap_prob)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    ds = np.array(ens.day_seconds)
    steady = float(ds[1:].sum())  # days 1..: rewind + simulation + route swaps
    if world > 1:
        t = torch.tensor([steady, total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        steady, total = float(t[0]), float(t[1])
    ttts, gaps = ens.gather_traces()
    if rank == 0:
        line = {"metric": "chain_days_per_sec", "value": args.chains * (args.days - 1) / steady, "unit": "chain-days/s", "n_gpus": world,
                "config": {"workload": args.workload, "chains": args.chains, "chains_per_gpu": len(ens.local), "days": args.days, "routes_per_od": args.k,
                           "mode": args.mode, "swap_prob": args.swap_prob},
                "total_seconds_incl_enumeration": total,
                "day_ms": {"rewind": float(ds[1:, 0].mean() * 1e3), "simulation": float(ds[1:, 1].mean() * 1e3), "route_swaps": float(ds[1:, 2].mean() * 1e3)},
                "mean_total_travel_time_by_day": ttts.mean(axis=0).tolist(), "mean_gap_by_day": gaps.mean(axis=0).tolist(),
                "link_volume_mean": float(ens.link_stats_sum[:, 0].sum() / args.chains)}
        # the same day in the reference C++ on one host core (logs on: dta_get_traveled_route reads them)
        try:
            ref, kind = bench.load_reference_lib()
            if kind == "reference":
                tab = sc.tables()
                E0 = sc.engine_from_tables(C.load_product(), tab, vehicle_logging_timestep_interval=-1, device=local_rank)
                E0.dta_enumerate_route_sets(args.k, 12345)
                store = E0.dta_route_sets_csr()
                E0.close()
                E = sc.engine_from_tables(ref, tab, vehicle_logging_timestep_interval=1, threads=1)
                a = time.perf_counter(); E.dta_set_route_sets(*store); b = time.perf_counter()
                E.main_loop(); c = time.perf_counter()
                E.dta_route_swap(0 if args.mode == "DUE" else 1, args.swap_prob, 1); d = time.perf_counter()
                E.close()
                cores = os.cpu_count() or 1
                line["reference_cpu"] = {"kind": kind, "one_core_day_ms": {"store": (b - a) * 1e3, "simulation": (c - b) * 1e3, "route_swap": (d - c) * 1e3},
                                         "cores": cores, "chain_days_per_sec_all_cores": cores / (d - a)}
        except Exception as e:  # the reference harness is optional here
            line["reference_cpu"] = {"unavailable": str(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
