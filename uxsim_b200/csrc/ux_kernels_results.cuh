// ux_kernels_results.cuh -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// Kernels off the hot path: clock advance, transposes, traveltime_actual, log scatter, vehicle snapshot, ensemble statistics.

// ---------------------------------------------------------------------------------------------------
// result kernels (off the hot path)
// ---------------------------------------------------------------------------------------------------
// [T][L] -> [L][T]
template <typename T>
__global__ void k_transpose(const T *src, T *dst, int rows, int cols) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * cols) return;
    int r = (int)(idx / cols), c = (int)(idx % cols);
    dst[(size_t)c * rows + r] = src[idx];
}

// Link::ensure_traveltime_real (traffi.cpp:631-648): position i takes the value of the latest event whose entry step <= i.
UX_DEV void tt_actual_link(const DevWorld &W, int l, double *out) {
    double cur = W.l_length[l] / W.l_vmax[l];
    int best = 0;
#pragma unroll 8
    for (int i = 0; i < W.T; i++) {  // the loads do not depend on the running arg-max, so they pipeline
        int ord = W.ev_ord[TLI(W, i, l)];
        double v = W.ev_tt[TLI(W, i, l)];
        if (ord > best) { best = ord; cur = v; }
        out[(size_t)l * W.T + i] = cur;
    }
}
__global__ void k_tt_actual(const DevWorld *worlds, int replica, double *out) {
    const DevWorld &W = worlds[replica];
    int l = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (l < W.L) tt_actual_link(W, l, out);
}

// order the RUN records per platoon and by time: record (vid, t) goes to off[vid] + (t - gen_ts[vid])
__global__ void k_log_scatter(const DevWorld *worlds, int replica, long long n, const long long *off, int *o_t, int *o_link, int *o_lane,
                              double *o_x, double *o_v, double *o_s) {
    const DevWorld &W = worlds[replica];
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int vid = W.lg_vid[i], t = W.lg_t[i];
    long long d = off[vid] + (t - W.v_gen_ts[vid]);
    o_t[d] = t; o_link[d] = W.lg_link[i]; o_lane[d] = W.lg_lane[i]; o_x[d] = W.lg_x[i]; o_v[d] = W.lg_v[i]; o_s[d] = W.lg_s[i];
}

// scatter ring state to per-vehicle arrays (x, v, lane)
__global__ void k_veh_snapshot(const DevWorld *worlds, int replica, double *vx, double *vv, int *vlane) {
    const DevWorld &W = worlds[replica];
    int gw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = (int)(threadIdx.x & 31);
    if (gw >= W.nseg) return;
    int l = W.seg_link[gw], part = W.seg_start[gw] & 0xffff, nparts = W.seg_start[gw] >> 16, cap = W.l_rcap[l];
    int cur = *W.t_base & 1;
    int hd = W.head[cur * W.L + l], n = W.tail[l] - hd;
    long long o = W.l_roff[l];
    for (int pos = part * 32 + lane; pos < n; pos += nparts * 32) {
        int p = ring_slot(hd, pos, cap);
        int vid = W.VID[o + p];
        vx[vid] = W.X[(size_t)cur * W.C + o + p]; vv[vid] = W.V[o + p]; vlane[vid] = W.LANE[o + p];
    }
}

// per-link ensemble statistics over this world's replicas: {volume, sum tt_actual, sum tt_actual^2, platoons on link}.
// Stage 1: thread per (link, replica) walks that replica's travel-time event table; stage 2: thread per link adds the
// replicas in index order (deterministic sums).
__global__ void k_ensemble_partial(const DevWorld *worlds, double *part) {
    const DevWorld &W = worlds[blockIdx.y];
    int l = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    int T = W.T, L = W.L;
    if (l >= L) return;
    double cur = W.l_length[l] / W.l_vmax[l], s1 = 0, s2 = 0;
    int best = 0;
#pragma unroll 8
    for (int i = 0; i < T; i++) {
        int ord = W.ev_ord[TLI(W, i, l)];
        double v = W.ev_tt[TLI(W, i, l)];
        if (ord > best) { best = ord; cur = v; }
        s1 += cur; s2 += cur * cur;
    }
    double *o = part + ((size_t)blockIdx.y * L + l) * 4;
    o[0] = T > 0 ? W.cum_arr[TLI(W, (T - 1), l)] : 0.0; o[1] = s1; o[2] = s2;
    o[3] = (double)(W.tail[l] - W.head[(*W.t_base & 1) * L + l]);
}
__global__ void k_ensemble_stats(const DevWorld *worlds, int R, const double *part, double *out) {
    int L = worlds[0].L;
    int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= L * 4) return;
    double a = 0;
    for (int r = 0; r < R; r++) a += part[(size_t)r * L * 4 + i];  // fixed replica order
    out[i] = a;
}

