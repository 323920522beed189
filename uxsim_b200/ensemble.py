"""Seeded-replica ensembles: the natural data-parallel shard of the simulation loop.

One scenario, many seeds.  Inside one GPU the replicas of a world run in the same kernel launches
(``blockIdx.y`` = replica); across GPUs each rank (one process per GPU, ``torch.distributed``) owns a
contiguous block of seeds and there is NO data-path collective: the only exchange is one all-reduce (NCCL over
NVLink; gloo in the CPU tests) of the per-link statistics tensor ``[L, 4]`` =
{traffic volume, sum of actual travel time, sum of squares, platoons left on the link} at the end of a run,
which is what a DUE/DSO solver iteration or a 32-seed ensemble consumes (DTAsolvers.py:157-253 averages the
same per-link quantities over days / seeds on the host).

PyTorch is used here for what it is good at: wrapping the engine's device buffers as tensors without a copy
(``__cuda_array_interface__``), pinned host staging, and ``torch.distributed``.  It does no simulation work.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _capi as C


class _CudaArrayView:
    """Zero-copy view of an engine-owned device buffer through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, n: int, dtype, owner):
        self._owner = owner  # keeps the world alive while the tensor exists
        self.__cuda_array_interface__ = {
            "shape": (n,), "typestr": np.dtype(dtype).str, "data": (ptr, False), "version": 3, "strides": None,
        }


def device_tensor(E: C.EngineWorld, what: int, replica: int = 0):
    """Result array of the CUDA engine as a torch CUDA tensor borrowing the engine's memory (valid until the
    next main_loop / getter that reuses the scratch buffer / destroy)."""
    import torch

    ptr, n, dt = E.get_device_array(what, replica)
    if n == 0:
        return torch.empty(0, dtype=torch.from_numpy(np.empty(0, dt)).dtype, device="cuda")
    return torch.as_tensor(_CudaArrayView(ptr, n, dt, E), device="cuda")


def shard_seeds(n_replicas: int, rank: int, world_size: int) -> range:
    """Contiguous block of replica indices owned by ``rank`` (balanced to within one replica)."""
    base, extra = divmod(n_replicas, world_size)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


@dataclass
class EnsembleResult:
    n_replicas: int                 # global number of replicas
    link_stats_sum: np.ndarray      # [L, 4] summed over ALL replicas of all ranks
    trips_completed: List[int]      # per local replica
    average_travel_time: List[float]
    platoon_updates: int            # local replicas
    loop_ms: float                  # device time of the local main_loop

    @property
    def mean_link_volume(self) -> np.ndarray:
        return self.link_stats_sum[:, 0] / self.n_replicas

    def mean_link_travel_time(self, total_timesteps: int) -> np.ndarray:
        return self.link_stats_sum[:, 1] / (self.n_replicas * total_timesteps)


def allreduce_link_stats(stats, group=None):
    """Sum a per-link statistics tensor over all ranks.  ``stats``: torch tensor (CUDA for NCCL, CPU for gloo)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


class Ensemble:
    """Run ``n_replicas`` seeded replicas of ``scenario`` (seeds ``base_seed + i``), sharded over the ranks of the
    default process group (if any), and reduce the per-link statistics."""

    def __init__(self, scenario, n_replicas: int, base_seed: int = 0, device: int = -1, api: Optional[C.CApi] = None, **world_over):
        import torch.distributed as dist

        self.scenario = scenario
        self.n_replicas = int(n_replicas)
        self.base_seed = int(base_seed)
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.local = shard_seeds(self.n_replicas, self.rank, self.world_size)
        self.api = api or C.load_product()
        self.device = device
        self.world_over = world_over
        self.engine: Optional[C.EngineWorld] = None

    def build(self):
        kw = dict(vehicle_logging_timestep_interval=-1, num_replicas=max(1, len(self.local)), device=self.device,
                  random_seed=self.base_seed + (self.local.start if len(self.local) else 0))
        kw.update(self.world_over)
        self.engine = self.scenario.to_engine(self.api, **kw)
        return self.engine

    def run(self) -> EnsembleResult:
        import torch

        E = self.engine or self.build()
        E.main_loop()
        nloc = len(self.local)
        L = E.get_int(C.I_NUM_LINKS)
        on_gpu = self.api.prefix == "ux_" and torch.cuda.is_available()
        if nloc == 0:
            stats = torch.zeros(L * 4, dtype=torch.float64, device="cuda" if on_gpu else "cpu")
        elif on_gpu:
            ptr = C.C.c_void_p(0)
            n = E._check(self.api.ensemble_link_stats_device(E.h, C.C.byref(ptr)))
            stats = torch.as_tensor(_CudaArrayView(ptr.value, int(n), np.float64, E), device="cuda").clone()
        else:
            stats = torch.from_numpy(E.ensemble_link_stats().reshape(-1).copy())
        stats = allreduce_link_stats(stats)
        trips, att, updates = [], [], 0
        for r in range(nloc):
            updates += int(E.get_array(C.A_LINK_STAT_COUNT, r).sum())
            st = E.get_array(C.A_VEH_STATE, r)
            tt = E.get_array(C.A_VEH_TRAVEL_TIME, r)
            done = st == C.VS_END
            trips.append(int(done.sum()))
            att.append(float(tt[done].mean()) if done.any() else 0.0)
        return EnsembleResult(self.n_replicas, stats.cpu().numpy().reshape(L, 4), trips, att,
                              updates, E.get_double(C.D_LAST_LOOP_MS))
