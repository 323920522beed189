"""Multi-process path on CPU: replica sharding + the all-reduce of per-link statistics, world_size 2, gloo.
The ranks drive the ORACLE library through the same Ensemble class (the CUDA engine needs a GPU); what is under
test is the host logic: seed sharding, the reduction, and that 2 ranks x 2 replicas == 4 single runs."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from uxsim_b200 import _capi as C
from uxsim_b200 import scenario as S
from uxsim_b200.ensemble import shard_seeds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_seeds_partition():
    for n in (1, 5, 8, 64):
        for ws in (1, 2, 3, 8):
            parts = [shard_seeds(n, r, ws) for r in range(ws)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from uxsim_b200 import _capi as C, scenario as S
    from uxsim_b200.ensemble import Ensemble, shard_seeds, allreduce_link_stats
    dist.init_process_group("gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    api = C.CApi(os.path.join({root!r}, "oracle", "liboracle.so"), "uxo_")
    sc = S.grid(n=4, seed=0, tmax=2400, demand_t=800)
    local = shard_seeds(4, rank, ws)
    L = None
    acc = None
    for i in local:   # the oracle runs one replica per world: loop over the local seeds
        E = sc.to_engine(api, vehicle_logging_timestep_interval=-1, random_seed=100 + i)
        E.main_loop()
        st = E.ensemble_link_stats()
        acc = st if acc is None else acc + st
    t = torch.from_numpy(acc.reshape(-1).copy())
    allreduce_link_stats(t)
    if rank == 0:
        np.save(sys.argv[1], t.numpy())
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_allreduce_equals_serial(tmp_path, oracle_api):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    out = tmp_path / "stats.npy"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", str(port), str(script), str(out)], check=True, timeout=300, capture_output=True)
    got = np.load(out)
    sc = S.grid(n=4, seed=0, tmax=2400, demand_t=800)
    ref = None
    for i in range(4):
        E = sc.to_engine(oracle_api, vehicle_logging_timestep_interval=-1, random_seed=100 + i)
        E.main_loop()
        st = E.ensemble_link_stats()
        ref = st if ref is None else ref + st
    assert np.array_equal(got.reshape(ref.shape), ref)


DTA_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch.distributed as dist
    from uxsim_b200 import _capi as C, scenario as S
    from uxsim_b200.dta import DayToDayEnsemble
    dist.init_process_group("gloo")
    api = C.CApi(os.path.join({root!r}, "tests", "emu", "_build", "libuxsim_emu.so"), "ux_")   # host emulation of the CUDA engine
    ens = DayToDayEnsemble(S.dta_grid9(), 4, mode="DUE", base_seed=20, api=api)
    assert len(ens.local) == 2
    ens.solve(max_iter=3, n_routes_per_od=5, swap_prob=0.2)
    ttts, gaps = ens.gather_traces()
    if dist.get_rank() == 0:
        np.savez(sys.argv[1], ttts=ttts, gaps=gaps, stats=ens.link_stats_sum)
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_day_to_day_ensemble_equals_single_process(tmp_path, emu_api):
    """4 DUE chains over 2 ranks (2 replicas each) == the same 4 chains in one process: chains are independent, seeds and
    swap seeds follow the global chain index, and the link statistics are all-reduced once at the end."""
    from uxsim_b200.dta import DayToDayEnsemble

    script = tmp_path / "dta_worker.py"
    script.write_text(DTA_WORKER.format(root=ROOT))
    out = tmp_path / "dta.npz"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", str(port), str(script), str(out)], check=True, timeout=600, capture_output=True)
    got = np.load(out)
    ens = DayToDayEnsemble(S.dta_grid9(), 4, mode="DUE", base_seed=20, api=emu_api)
    ens.solve(max_iter=3, n_routes_per_od=5, swap_prob=0.2)
    assert np.array_equal(got["ttts"], ens.ttts) and np.array_equal(got["gaps"], ens.t_gaps)
    assert np.allclose(got["stats"], ens.link_stats_sum, rtol=1e-12)   # sum over ranks re-associates the replica sum
