"""Prototype (not a test): order of the label-correcting worklist on the C4 grid's own edge costs.
Counts rounds, chunks (128 nodes per pass), pops, re-circulations and L2 requests per destination for plain FIFO order
(delta = 1e9) and for the near-far order of k_sssp_wave at several bucket widths.  The costs are the instantaneous link
travel times of one C4 run at its 12 route-update steps (tests/micro/gpu_dump_costs.py on a B200).
    python tests/micro/sssp_order_proto.py"""
import os
import numpy as np, sys, time
d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'costs_grid100.npz'))
tt, st, en = d['tt'], d['start'].astype(np.int64), d['end'].astype(np.int64)
L, K = tt.shape
N = int(max(st.max(), en.max())) + 1
# in-CSR by end node
order = np.argsort(en, kind='stable')
ein_off = np.zeros(N + 1, np.int64); np.add.at(ein_off, en + 1, 1); ein_off = np.cumsum(ein_off)
ein_from = st[order]

def nearfar2(c, k, delta, nth=128):
    cin = c[order]
    dist = np.full(N, np.inf); dist[k] = 0
    dirty = np.zeros(N, bool); dirty[k] = True
    tau = delta; rounds = pops = relax = recirc = chunks = succ = 0
    while True:
        Q = np.nonzero(dirty)[0]
        if len(Q) == 0: break
        m = dist[Q].min()
        if m >= tau: tau = m + delta
        rounds += 1; chunks += -(-len(Q) // nth)
        near = Q[dist[Q] < tau]
        recirc += len(Q) - len(near)
        F = near; pops += len(F); dirty[F] = False
        cnt = ein_off[F + 1] - ein_off[F]
        idx = np.repeat(ein_off[F], cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        src = np.repeat(F, cnt)
        nd = cin[idx] + dist[src]
        i = ein_from[idx]
        relax += len(idx); succ += int((nd < dist[i]).sum())
        old = dist.copy()
        np.minimum.at(dist, i, nd)
        dirty[dist < old] = True
    return rounds, chunks, pops, recirc, recirc + pops * 5 + succ
rng = np.random.default_rng(0)
dests = rng.integers(0, N, 6)
for k in (0, 2, 4, 8, 11):
    c = tt[:, k] * (1 + 0.01 * rng.random(L))
    for dst in dests[:2]:
        out = f"snap {k} mean {c.mean():.0f} dest {dst}:"
        for delta in (1e9, c.mean()/4, c.mean()/2, c.mean(), 2*c.mean(), 4*c.mean()):
            r, ch, p, rc, req = nearfar2(c, dst, delta)
            out += f" | d={delta:.0f}: rounds {r} chunks {ch} pops {p} recirc {rc} L2req {req}"
        print(out, flush=True)
