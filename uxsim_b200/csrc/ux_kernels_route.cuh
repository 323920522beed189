// ux_kernels_route.cuh -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// Route choice: edge costs, per-destination shortest paths (k_sssp, k_sssp_frontier), DUO preference update.

// ---------------------------------------------------------------------------------------------------
// Route choice: update_adj_time_matrix (traffi.cpp:1267-1306), route_search_all (1317-1385, as the
// per-destination fixed point of oracle R2), route_choice_duo (1486-1553).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_edge_cost(const DevWorld *worlds, int step_off) {
    LOAD_WORLD(W, worlds)
    int e = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (e >= W.n_edges) return;
    int g = W.e_link[e], t = *W.t_base + step_off;
    double cost = 0.0;
    int m0 = W.l_pair_off[g], m1 = W.l_pair_off[g + 1], cntr = 0;
    for (int m = m0; m < m1; m++) {  // members of the node pair in link-id order, running average
        int l = W.l_pair_list[m];
        double tt = W.tt_inst[TLI(W, t, l)];
        if (tt <= 0.0) tt = W.l_length[l] / W.l_vmax[l];
        double pen = W.l_toll ? W.l_toll[(size_t)t * W.L + l] : W.l_penalty[l];
        double c;
        if (W.hard_det) c = tt + pen;
        else c = tt * (1.0 + W.noise * ux_rng_uniform(W.seed, (uint32_t)t, (uint32_t)l, UX_RNG_LINKNOISE, 0u)) + pen;
        double nn = (double)cntr;
        cost = cost * nn / (nn + 1.0) + c / (nn + 1.0);
        cntr++;
        if (W.l_cap_in[l] == 0.0) cost = INFINITY;
    }
    W.pair_cost[e] = cost;
    if (W.cost_hist) W.cost_hist[(size_t)(t / W.route_ts) * (size_t)W.n_edges + (size_t)e] = cost;
    double cf = (cost > 0.0) ? cost : INFINITY;  // adj > 0 filter (traffi.cpp:1331) folded in: +inf never relaxes anything
    W.cost_in[W.e_inpos[e]] = cf;
    W.cost_out[W.e_outpos[e]] = cf;
}

// One block per active destination: label-correcting relaxation to the fixed point dist[i] = min_j c_ij + dist[j]
// (oracle R2).  The destination's dist / next rows live in shared memory for the whole search (12 B per node, up to
// ~19k nodes in 227 KB; larger graphs fall back to the rows in global memory), so a sweep over the edges costs
// shared-memory atomics only; the fixed point does not depend on the relaxation order, so the result is deterministic.
#define SSSP_KE 8
#define SSSP_THREADS 256
template <bool USE_SMEM>
__global__ void __launch_bounds__(SSSP_THREADS) k_sssp(const DevWorld *worlds) {
    LOAD_WORLD(W, worlds)
    int row = (int)blockIdx.x;
    if (row >= W.D) return;
    int N = W.N, k = W.row_dest[row], E = W.n_edges;
    double *gdist = W.rdist + (size_t)row * N;
    int *gnext = W.rnext + (size_t)row * N;
    const double *cost = W.pair_cost;
    const int *e_from = W.e_from, *e_to = W.e_to;
#ifdef UX_EMU
    if (threadIdx.x != 0) return;
    double *dist = gdist; int *next = gnext;
    for (int i = 0; i < N; i++) { dist[i] = INFINITY; next[i] = 0x7fffffff; }
    dist[k] = 0.0;
    for (bool changed = true; changed;) {
        changed = false;
        for (int e = 0; e < E; e++) {
            double c = cost[e];
            if (!(c > 0.0) || is_inf(c)) continue;
            double nd = c + dist[e_to[e]];
            if (nd < dist[e_from[e]]) { dist[e_from[e]] = nd; changed = true; }
        }
    }
    int tid = 0, nth = 1;
#else
    extern __shared__ unsigned long long sssp_smem[];
    // compile-time address space: with a run-time select the loads / atomics below become generic and crawl
    double *dist; int *next; unsigned long long *du;
    if (USE_SMEM) { du = sssp_smem; dist = reinterpret_cast<double *>(sssp_smem); next = reinterpret_cast<int *>(sssp_smem + N); }
    else { dist = gdist; next = gnext; du = reinterpret_cast<unsigned long long *>(gdist); }
    int tid = (int)threadIdx.x, nth = (int)blockDim.x;
    // labels in global memory are relaxed with L2 atomics by other threads of the block: plain loads could keep hitting a stale L1 line
#define SSSP_LD(i_) (USE_SMEM ? du[i_] : __ldcg(du + (i_)))
    for (int i = tid; i < N; i += nth) { dist[i] = (i == k) ? 0.0 : INFINITY; next[i] = 0x7fffffff; }
    // the first SSSP_KE edges of every thread stay in registers for all sweeps; only larger graphs re-read the rest
    double rc[SSSP_KE]; int rf[SSSP_KE], rt[SSSP_KE];
#pragma unroll
    for (int q = 0; q < SSSP_KE; q++) {
        int e = tid + q * nth;
        double c = e < E ? __ldg(cost + e) : INFINITY;
        if (!(c > 0.0)) c = INFINITY;  // adj > 0 filter (traffi.cpp:1331): such edges never relax anything
        rc[q] = c; rf[q] = e < E ? __ldg(e_from + e) : 0; rt[q] = e < E ? __ldg(e_to + e) : 0;
    }
    __syncthreads();
    while (true) {
        int changed = 0;
        // two relaxation passes per barrier: the fixed point does not depend on the order or staleness of the reads
        for (int rep = 0; rep < 2; rep++) {
#pragma unroll
        for (int q = 0; q < SSSP_KE; q++) {
            double nd = rc[q] + __longlong_as_double((long long)SSSP_LD(rt[q]));  // inf + x = inf: never smaller
            if (nd < __longlong_as_double((long long)SSSP_LD(rf[q]))) {  // non-negative doubles order like their bit patterns
                atomicMin(&du[rf[q]], (unsigned long long)__double_as_longlong(nd));
                changed = 1;
            }
        }
        for (int e = tid + SSSP_KE * nth; e < E; e += nth) {
            double c = __ldg(cost + e);
            if (!(c > 0.0) || is_inf(c)) continue;
            double nd = c + __longlong_as_double((long long)SSSP_LD(__ldg(e_to + e)));
            int i = __ldg(e_from + e);
            if (nd < __longlong_as_double((long long)SSSP_LD(i))) {
                atomicMin(&du[i], (unsigned long long)__double_as_longlong(nd));
                changed = 1;
            }
        }
        }
        if (!__syncthreads_or(changed)) break;
    }
#endif
    // next hop: smallest node j with c_ij + dist[j] == dist[i]
#ifdef UX_EMU
#define SSSP_DIST(i_) dist[i_]
#define SSSP_NEXT(i_) next[i_]
#else
#define SSSP_DIST(i_) __longlong_as_double((long long)SSSP_LD(i_))
#define SSSP_NEXT(i_) (USE_SMEM ? next[i_] : __ldcg(next + (i_)))
#endif
    int any = 0;
    for (int e = tid; e < E; e += nth) {
        double c = cost[e];
        if (!(c > 0.0) || is_inf(c)) continue;
        int i = e_from[e], j = e_to[e];
        if (i == k) continue;
        double dj = SSSP_DIST(j);
        if (is_inf(dj)) continue;
        if (c + dj == SSSP_DIST(i)) { atomicMin(&next[i], j); any = 1; }
    }
#ifndef UX_EMU
    any = __syncthreads_or(any);
#endif
    for (int i = tid; i < N; i += nth) {
        int nx = SSSP_NEXT(i);
        if (i == k) nx = k; else if (nx == 0x7fffffff) nx = -1;
        gnext[i] = nx;
        gdist[i] = SSSP_DIST(i);
    }
#undef SSSP_DIST
#undef SSSP_NEXT
#ifndef UX_EMU
#undef SSSP_LD
#endif
    if (tid == 0) W.has_path[row] = any ? 1 : 0;
}

// Frontier (label-correcting worklist) variant for graphs whose rows fit in shared memory: only the in-edges of nodes
// whose distance just improved are relaxed, so the work is ~E per destination instead of (sweeps x E).  Same fixed
// point, hence the same dist / next as k_sssp and the oracle.  Shared memory: dist 8N + next 4N + queued flag N +
// two uint16 worklists 4N = 17N bytes (N < 65536).
__global__ void __launch_bounds__(SSSP_THREADS) k_sssp_frontier(const DevWorld *worlds) {
#ifndef UX_EMU
    LOAD_WORLD(W, worlds)
    int row = (int)blockIdx.x;
    if (row >= W.D) return;
    int N = W.N, k = W.row_dest[row], E = W.n_edges;
    extern __shared__ unsigned long long sssp_smem[];
    unsigned long long *du = sssp_smem;
    double *dist = reinterpret_cast<double *>(sssp_smem);
    int *next = reinterpret_cast<int *>(sssp_smem + N);
    unsigned short *fa = reinterpret_cast<unsigned short *>(next + N), *fb = fa + N;
    unsigned char *inq = reinterpret_cast<unsigned char *>(fb + N);
    __shared__ int nf_s[2];
    const double *cost = W.pair_cost;
    int tid = (int)threadIdx.x, nth = (int)blockDim.x;
    for (int i = tid; i < N; i += nth) { dist[i] = (i == k) ? 0.0 : INFINITY; next[i] = 0x7fffffff; inq[i] = 0; }
    if (tid == 0) { fa[0] = (unsigned short)k; nf_s[0] = 1; nf_s[1] = 0; }
    __syncthreads();
    int cur = 0;
    while (true) {
        int nf = nf_s[cur];
        if (nf == 0) break;
        unsigned short *F = cur ? fb : fa, *G = cur ? fa : fb;
        for (int base = 0; base < nf; base += nth) {
            int idx = base + tid, j = -1;
            double dj = 0.0;
            if (idx < nf) { j = F[idx]; dj = dist[j]; inq[j] = 0; }
            __syncthreads();  // every queued flag of this chunk is cleared before anyone can re-queue
            if (j >= 0) {
                for (int q = __ldg(W.ein_off + j); q < __ldg(W.ein_off + j + 1); q++) {
                    double c = __ldg(cost + __ldg(W.ein_eid + q));
                    if (!(c > 0.0) || is_inf(c)) continue;
                    int i = __ldg(W.ein_from + q);
                    double nd = c + dj;
                    if (nd < __longlong_as_double((long long)du[i])) {
                        unsigned long long old = atomicMin(&du[i], (unsigned long long)__double_as_longlong(nd));
                        if ((unsigned long long)__double_as_longlong(nd) < old) {
                            unsigned int *w32 = reinterpret_cast<unsigned int *>(inq + (i & ~3));
                            unsigned int bit = 1u << ((i & 3) * 8);
                            if (!(atomicOr(w32, bit) & bit)) G[atomicAdd(&nf_s[cur ^ 1], 1)] = (unsigned short)i;
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (tid == 0) nf_s[cur] = 0;
        cur ^= 1;
        __syncthreads();
    }
    int any = 0;
    for (int e = tid; e < E; e += nth) {  // next hop: smallest node j with c_ij + dist[j] == dist[i]
        double c = __ldg(cost + e);
        if (!(c > 0.0) || is_inf(c)) continue;
        int i = __ldg(W.e_from + e), j = __ldg(W.e_to + e);
        if (i == k) continue;
        double dj = dist[j];
        if (is_inf(dj)) continue;
        if (c + dj == dist[i]) { atomicMin(&next[i], j); any = 1; }
    }
    any = __syncthreads_or(any);
    double *gdist = W.rdist + (size_t)row * N;
    int *gnext = W.rnext + (size_t)row * N;
    for (int i = tid; i < N; i += nth) {
        int nx = next[i];
        if (i == k) nx = k; else if (nx == 0x7fffffff) nx = -1;
        gnext[i] = nx;
        gdist[i] = dist[i];
    }
    if (tid == 0) W.has_path[row] = any ? 1 : 0;
#endif
}

// Wavefront variant for large sparse graphs (N >= 2048): a label-correcting worklist like k_sssp_frontier, but
//  * the destination's dist row stays where it will end up -- in its global row, i.e. in the 126 MB L2 -- and only the
//    worklist lives in shared memory: one circular queue of node ids (a node is queued at most once, so N entries hold
//    the current and the next wavefront together) plus the queued bitmap, 2N + N/8 bytes.  k_sssp_frontier keeps 17N
//    bytes per destination in shared memory and runs ONE destination per SM; here 7-10 destinations per SM are in flight
//    (the launcher pads shared memory so that the rows of the resident blocks, 8N bytes each, stay L2-resident);
//  * the worklist is processed near-far (Delta-stepping with a single queue): only nodes with label < tau relax their
//    in-edges, the others go round once more; when no queued label is below tau, tau jumps to (smallest queued label +
//    Delta), Delta = 2 x mean edge cost (k_cost_stats).  On congested networks (edge costs spread over two orders of
//    magnitude) plain FIFO order re-relaxes every node 5-10 times; this order brings it to ~1.1 and the L2 requests per
//    destination from ~450k to ~105k (tests/micro/sssp_order_proto.py, measured on the C4 grid's own costs).
// tau and the queue minimum are kept on the grid of the labels' upper 32 bits (monotone for non-negative doubles): the
// order only steers the work, the fixed point -- hence dist / next -- is the one of k_sssp and the oracle.
// Relaxation = red.min on the bit pattern (no return value needed: whoever saw a larger label queues the node, the
// bitmap removes duplicates).
#ifndef SSSP_WAVE_THREADS
#define SSSP_WAVE_THREADS 256
#endif
__global__ void __launch_bounds__(256) k_cost_stats(const DevWorld *worlds, double factor) {
#ifndef UX_EMU
    const DevWorld &W = worlds[blockIdx.y];
    __shared__ double ssum[256];
    __shared__ int scnt[256];
    int tid = (int)threadIdx.x;
    double s = 0.0; int c = 0;
    for (int e = tid; e < W.n_edges; e += 256) { double v = W.cost_in[e]; if (v < INFINITY) { s += v; c++; } }
    ssum[tid] = s; scnt[tid] = c;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) { if (tid < h) { ssum[tid] += ssum[tid + h]; scnt[tid] += scnt[tid + h]; } __syncthreads(); }
    if (tid == 0) *W.sssp_delta = scnt[0] > 0 ? factor * ssum[0] / (double)scnt[0] : 1.0;
#endif
}

// SMEM_DIST: graphs small enough for the label row to sit in shared memory next to the queue (8N more bytes) keep it there --
// same near-far order, shared-memory atomics instead of L2 round trips -- and copy it out at the end.
// WARP: one destination per WARP instead of per block (graphs whose wavefront rarely exceeds 32 nodes): every barrier becomes a
// __syncwarp, no warp idles through the rounds of a narrow wavefront, and the block's warps share one staged DevWorld.
#ifndef SSSP_WARPS_PER_BLOCK
#define SSSP_WARPS_PER_BLOCK 4
#endif
// STAGE (with WARP): the block's warps -- up to SSSP_STAGE_MAX_WARPS destinations of ONE replica -- first copy what every
// relaxation reads into shared memory: the in-edge CSR (offsets, start nodes as uint16) and the replica's edge costs in that
// order.  A round's chain "pop -> label -> CSR offsets -> start nodes / costs -> neighbour labels" then has no global round
// trip left; `stage_words` 8-byte words precede the per-warp regions.
#define SSSP_STAGE_MAX_WARPS 16
template <bool SMEM_DIST, bool WARP, bool STAGE = false>
__global__ void __launch_bounds__(STAGE ? 32 * SSSP_STAGE_MAX_WARPS : SSSP_WAVE_THREADS) k_sssp_wave(const DevWorld *worlds, int smem_words_per_warp, int stage_words = 0) {
#ifndef UX_EMU
    LOAD_WORLD(W, worlds)
    const int wib = WARP ? (int)(threadIdx.x >> 5) : 0;
    int row = WARP ? (int)blockIdx.x * (int)(blockDim.x >> 5) + wib : (int)blockIdx.x;
    extern __shared__ unsigned long long sssp_smem_all[];
    const int *s_off = nullptr; const unsigned short *s_from = nullptr; const double *s_cin = nullptr;
    if (STAGE) {
        const int N_ = W.N, E_ = W.n_edges;
        double *c_ = reinterpret_cast<double *>(sssp_smem_all);
        int *o_ = reinterpret_cast<int *>(c_ + E_);
        unsigned short *f_ = reinterpret_cast<unsigned short *>(o_ + N_ + 1);
        for (int e = (int)threadIdx.x; e < E_; e += (int)blockDim.x) { c_[e] = W.cost_in[e]; f_[e] = (unsigned short)__ldg(W.ein_from + e); }
        for (int i = (int)threadIdx.x; i <= N_; i += (int)blockDim.x) o_[i] = __ldg(W.ein_off + i);
        __syncthreads();
        s_off = o_; s_from = f_; s_cin = c_;
    }
    if (row >= W.D) return;  // WARP: whole warps leave; only warp-level synchronisation follows
    int N = W.N, k = W.row_dest[row];
    int tid = WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x, nth = WARP ? 32 : (int)blockDim.x, cap = N + 2 * nth;
    unsigned long long *sssp_smem = sssp_smem_all + (STAGE ? (size_t)stage_words : 0) + (WARP ? (size_t)wib * smem_words_per_warp : 0);
#define WAVE_SYNC() do { if (WARP) __syncwarp(); else __syncthreads(); } while (0)
    unsigned long long *du_s = sssp_smem;  // [N] when SMEM_DIST
    unsigned short *Q = reinterpret_cast<unsigned short *>(sssp_smem + (SMEM_DIST ? N : 0));
    unsigned int *inq = reinterpret_cast<unsigned int *>(Q + ((cap + 3) & ~3));
    __shared__ int q_tail_all[SSSP_STAGE_MAX_WARPS];
    __shared__ unsigned int lo_min_all[SSSP_STAGE_MAX_WARPS][2];  // per round parity: lower bound of the queued labels (upper 32 bits)
    int &q_tail = q_tail_all[wib];
    unsigned int *lo_min = lo_min_all[wib];
    unsigned long long *du_g = reinterpret_cast<unsigned long long *>(W.rdist + (size_t)row * N);
#define WAVE_LD(i) (SMEM_DIST ? du_s[i] : __ldcg(du_g + (i)))
    int *gnext = W.rnext + (size_t)row * N;
    const double *cin = W.cost_in;
    const int *ein_off = W.ein_off, *ein_from = W.ein_from;
    const unsigned long long INF_BITS = 0x7ff0000000000000ull;
    double delta = *W.sssp_delta;
    for (int i = tid; i < N; i += nth) { if (SMEM_DIST) du_s[i] = i == k ? 0ull : INF_BITS; else __stcg(du_g + i, i == k ? 0ull : INF_BITS); }
    for (int i = tid; i < (N + 31) / 32; i += nth) inq[i] = 0u;
    if (tid == 0) { Q[0] = (unsigned short)k; q_tail = 1; inq[k >> 5] = 1u << (k & 31); lo_min[0] = 0u; lo_min[1] = 0xffffffffu; }
    WAVE_SYNC();
    int head = 0;  // monotone counters; slot = counter % cap
    // (counter % cap without a division: umulhi(x, ceil(2^32 / cap)) is floor(x / cap) or one more, for any x < 2^31)
    const unsigned cap_m = (unsigned)((0x100000000ull + (unsigned long long)cap - 1ull) / (unsigned long long)cap);
    auto wave_slot = [cap, cap_m](int x) { int r = x - (int)__umulhi((unsigned)x, cap_m) * cap; return r < 0 ? r + cap : r; };
#define WAVE_SLOT(x) wave_slot(x)
    unsigned int tau_hi = 0u;
    for (int r = 0;; r++) {
        int tail = q_tail, p = r & 1;
        if (tail == head) break;
        unsigned int m_hi = lo_min[p];
        if (m_hi >= tau_hi) {  // nothing queued is below tau: jump
            double tnew = __longlong_as_double((long long)((unsigned long long)m_hi << 32)) + delta;
            unsigned int th = (unsigned int)((unsigned long long)__double_as_longlong(tnew) >> 32);
            tau_hi = th > m_hi ? th : m_hi + 1u;
        }
        unsigned int mymin = 0xffffffffu;
        for (int base = head; base < tail; base += nth) {
            int idx = base + tid, j = -1;
            unsigned long long dj = 0ull;
            bool near = false, far = false;
            if (idx < tail) {
                j = Q[WAVE_SLOT(idx)]; dj = WAVE_LD(j);
                near = (unsigned int)(dj >> 32) < tau_hi;
                far = !near;
                if (near) atomicAnd(&inq[j >> 5], ~(1u << (j & 31)));
            }
            // every queued bit of this chunk is cleared before anyone can re-queue, and everybody has read q_tail / lo_min[p]
            // before the first append of the round
            WAVE_SYNC();
            if (base == head && tid == 0) lo_min[p] = 0xffffffffu;  // round r+1 collects into it
            if (far) {  // stays queued (its bit is still set): goes round once more
                unsigned int hj = (unsigned int)(dj >> 32);
                Q[WAVE_SLOT(atomicAdd(&q_tail, 1))] = (unsigned short)j;
                if (hj < mymin) mymin = hj;
            }
            if (WARP) {
                // Warp per destination: the in-edges of the chunk's near nodes form ONE list (prefix sum of the in-degrees over the
                // lanes) that the lanes share out edge by edge.  A wavefront here is ~10 nodes of in-degree ~4: one lane per NODE keeps a
                // third of the warp busy for max-degree/4 serial batches, one lane per EDGE fills the warp for one or two passes -- the
                // kernel is bound by instruction issue (74 % of the issue slots at 1024 replicas), so the instruction count is the time.
                int q0 = 0, dg = 0;
                if (near) { q0 = STAGE ? s_off[j] : __ldg(ein_off + j); dg = (STAGE ? s_off[j + 1] : __ldg(ein_off + j + 1)) - q0; }
                int incl = dg;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= d) incl += v; }
                const int total = __shfl_sync(0xffffffffu, incl, 31), pre = incl - dg;
                for (int eb = 0; eb < total; eb += 32) {
                    const bool in = eb + tid < total;
                    const int e = in ? eb + tid : 0;
                    int lo = 0;  // owner: the largest lane whose list offset is <= e (lanes without edges share their successor's offset and lose)
#pragma unroll
                    for (int st = 16; st >= 1; st >>= 1) { int v = __shfl_sync(0xffffffffu, pre, lo + st); if (v <= e) lo += st; }
                    const unsigned long long djo = __shfl_sync(0xffffffffu, dj, lo);
                    const int q = __shfl_sync(0xffffffffu, q0, lo) + (e - __shfl_sync(0xffffffffu, pre, lo));
                    if (in) {
                        const int i = STAGE ? (int)s_from[q] : __ldg(ein_from + q);
                        const unsigned long long nd = (unsigned long long)__double_as_longlong((STAGE ? s_cin[q] : __ldg(cin + q)) + __longlong_as_double((long long)djo));
                        if (nd < WAVE_LD(i)) {
                            atomicMin((SMEM_DIST ? du_s : du_g) + i, nd);
                            unsigned int hn = (unsigned int)(nd >> 32), bit = 1u << (i & 31);
                            if (hn < mymin) mymin = hn;
                            if (!(atomicOr(&inq[i >> 5], bit) & bit)) Q[WAVE_SLOT(atomicAdd(&q_tail, 1))] = (unsigned short)i;
                        }
                    }
                }
            } else if (near) {
                double djd = __longlong_as_double((long long)dj);
                int q0 = STAGE ? s_off[j] : __ldg(ein_off + j), q1 = STAGE ? s_off[j + 1] : __ldg(ein_off + j + 1);
                for (int q = q0; q < q1; q += 4) {  // loads of four in-edges in flight together
                    int ii[4]; unsigned long long nd[4], cu[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        bool in = q + u < q1;
                        ii[u] = in ? (STAGE ? (int)s_from[q + u] : __ldg(ein_from + q + u)) : -1;
                        nd[u] = in ? (unsigned long long)__double_as_longlong((STAGE ? s_cin[q + u] : __ldg(cin + q + u)) + djd) : INF_BITS;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) cu[u] = ii[u] >= 0 ? WAVE_LD(ii[u]) : 0ull;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (ii[u] >= 0 && nd[u] < cu[u]) {
                            atomicMin((SMEM_DIST ? du_s : du_g) + ii[u], nd[u]);
                            unsigned int hn = (unsigned int)(nd[u] >> 32), bit = 1u << (ii[u] & 31);
                            if (hn < mymin) mymin = hn;
                            if (!(atomicOr(&inq[ii[u] >> 5], bit) & bit)) Q[WAVE_SLOT(atomicAdd(&q_tail, 1))] = (unsigned short)ii[u];
                        }
                    }
                }
            }
            if (base + nth < tail) WAVE_SYNC();  // the next chunk's pops (label read + bit clear) must not overlap these relaxations
        }
        unsigned int wm = __reduce_min_sync(0xffffffffu, mymin);
        if ((tid & 31) == 0 && wm != 0xffffffffu) atomicMin(&lo_min[p ^ 1], wm);
        head = tail;
        WAVE_SYNC();  // this round's red.min / queue appends are visible before the next round reads them
    }
    // next hop: smallest end node j of an out-edge with c_ij + dist[j] == dist[i] (out-edges are stored ends ascending)
    int any = 0;
    const double *cout_ = W.cost_out;
    for (int i = tid; i < N; i += nth) {
        int nx = -1;
        if (i == k) { nx = k; if (SMEM_DIST) du_g[i] = 0ull; }
        else {
            unsigned long long dib = WAVE_LD(i);
            double di = __longlong_as_double((long long)dib);
            if (SMEM_DIST) du_g[i] = dib;
            if (!is_inf(di)) {
                for (int q = __ldg(W.eout_off + i); q < __ldg(W.eout_off + i + 1); q++) {
                    int j = __ldg(W.eout_to + q);
                    if (__ldg(cout_ + q) + __longlong_as_double((long long)WAVE_LD(j)) == di) { nx = j; any = 1; break; }
                }
            }
        }
        gnext[i] = nx;
    }
    any = WARP ? __any_sync(0xffffffffu, any) : __syncthreads_or(any);
    if (tid == 0) W.has_path[row] = any ? 1 : 0;
#undef WAVE_LD
#undef WAVE_SYNC
#undef WAVE_SLOT
#endif
}

// grid = (link tiles, destination rows, replicas): no index arithmetic beyond one add, coalesced rows of `pref`.
// The kernel is a pure stream (16 B per element in and out): a thread takes DUO_PER_THREAD consecutive links with 16-byte loads and stores, so
// that enough bytes are in flight per SM to approach the HBM rate.
#define DUO_PER_THREAD 4
UX_DEV double duo_one(double p, bool on_path, double wt) { return on_path ? (1.0 - wt) * p + wt : (1.0 - wt) * p; }
__global__ void __launch_bounds__(256) k_duo(const DevWorld *worlds, int gradual_step) {
    const DevWorld &W = worlds[blockIdx.z];  // a handful of block-uniform field loads; nothing here aliases them
    int l0 = (int)(blockIdx.x * blockDim.x + threadIdx.x) * DUO_PER_THREAD, r = (int)blockIdx.y, L = W.L;
    if (l0 >= L || r >= W.D) return;
    double wt = gradual_step ? W.duo_w * (W.dt / W.duo_time) : W.duo_w;
    if (!W.pref_active[r]) wt = 1.0;
    double *pp = W.pref + (size_t)r * (size_t)L + (size_t)l0;
    const int *rn = W.rnext + (size_t)r * W.N;
#ifndef UX_EMU
    if (l0 + DUO_PER_THREAD <= L && (reinterpret_cast<unsigned long long>(pp) & 15ull) == 0ull) {
        double2 a = *reinterpret_cast<const double2 *>(pp), b = *reinterpret_cast<const double2 *>(pp + 2);
        int nx[4], en[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { nx[u] = rn[W.l_start[l0 + u]]; en[u] = W.l_end[l0 + u]; }
        a.x = duo_one(a.x, nx[0] == en[0], wt); a.y = duo_one(a.y, nx[1] == en[1], wt);
        b.x = duo_one(b.x, nx[2] == en[2], wt); b.y = duo_one(b.y, nx[3] == en[3], wt);
        *reinterpret_cast<double2 *>(pp) = a; *reinterpret_cast<double2 *>(pp + 2) = b;
        return;
    }
#endif
    for (int u = 0; u < DUO_PER_THREAD && l0 + u < L; u++) pp[u] = duo_one(pp[u], rn[W.l_start[l0 + u]] == W.l_end[l0 + u], wt);
}

__global__ void k_duo_finish(const DevWorld *worlds) {
    const DevWorld &W = worlds[blockIdx.y];
    int r = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (r < W.D && W.has_path[r]) W.pref_active[r] = 1;
}

__global__ void k_advance(const DevWorld *worlds, int nsteps) {
    const DevWorld &W = worlds[blockIdx.y];
    if (blockIdx.x == 0 && threadIdx.x == 0) *W.t_base += nsteps;
}

