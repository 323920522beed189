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
            double nd = rc[q] + __longlong_as_double((long long)du[rt[q]]);  // inf + x = inf: never smaller
            if (nd < __longlong_as_double((long long)du[rf[q]])) {  // non-negative doubles order like their bit patterns
                atomicMin(&du[rf[q]], (unsigned long long)__double_as_longlong(nd));
                changed = 1;
            }
        }
        for (int e = tid + SSSP_KE * nth; e < E; e += nth) {
            double c = __ldg(cost + e);
            if (!(c > 0.0) || is_inf(c)) continue;
            double nd = c + __longlong_as_double((long long)du[__ldg(e_to + e)]);
            int i = __ldg(e_from + e);
            if (nd < __longlong_as_double((long long)du[i])) {
                atomicMin(&du[i], (unsigned long long)__double_as_longlong(nd));
                changed = 1;
            }
        }
        }
        if (!__syncthreads_or(changed)) break;
    }
#endif
    // next hop: smallest node j with c_ij + dist[j] == dist[i]
    int any = 0;
    for (int e = tid; e < E; e += nth) {
        double c = cost[e];
        if (!(c > 0.0) || is_inf(c)) continue;
        int i = e_from[e], j = e_to[e];
        if (i == k) continue;
        double dj = dist[j];
        if (is_inf(dj)) continue;
        if (c + dj == dist[i]) { atomicMin(&next[i], j); any = 1; }
    }
#ifndef UX_EMU
    any = __syncthreads_or(any);
#endif
    for (int i = tid; i < N; i += nth) {
        int nx = next[i];
        if (i == k) nx = k; else if (nx == 0x7fffffff) nx = -1;
        gnext[i] = nx;
        gdist[i] = dist[i];
    }
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

__global__ void __launch_bounds__(256) k_duo(const DevWorld *worlds, int gradual_step) {
    LOAD_WORLD(W, worlds)
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)W.D * W.L;
    if (idx >= total) return;
    int r = (int)(idx / W.L), l = (int)(idx % W.L);
    double wt = gradual_step ? W.duo_w * (W.dt / W.duo_time) : W.duo_w;
    if (!W.pref_active[r]) wt = 1.0;
    double p = W.pref[idx];
    if (W.rnext[(size_t)r * W.N + W.l_start[l]] == W.l_end[l]) p = (1.0 - wt) * p + wt;
    else p = (1.0 - wt) * p;
    W.pref[idx] = p;
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

