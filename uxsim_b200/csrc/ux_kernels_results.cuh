// ux_kernels_results.cuh -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// Kernels off the hot path: clock advance, transposes, traveltime_actual, log scatter, vehicle snapshot, ensemble statistics.

// ---------------------------------------------------------------------------------------------------
// result kernels (off the hot path)
// ---------------------------------------------------------------------------------------------------
// [T][L][RS] (one replica's column of a replica-minor series) -> [L][T] contiguous
template <typename T>
__global__ void k_transpose(const T *src, T *dst, int rows, int cols, int RS) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * cols) return;
    int r = (int)(idx / cols), c = (int)(idx % cols);
    dst[(size_t)c * rows + r] = src[(size_t)idx * RS];
}
// one member of a replica-minor record array -> contiguous: dst[i] = the T at src + i * stride bytes
template <typename T>
__global__ void k_gather(const char *src, size_t stride, long long n, T *dst) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = *reinterpret_cast<const T *>(src + (size_t)i * stride);
}
// ... the same member of ALL replicas, replica-major: dst[r * n + i] = the T at src + (i * RS + r) * rec bytes
template <typename T>
__global__ void k_gather_all(const char *src, size_t rec, int RS, int R, long long n, T *dst) {
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n * R) return;
    long long i = q / R; int r = (int)(q % R);
    dst[(size_t)r * n + i] = *reinterpret_cast<const T *>(src + ((size_t)i * RS + r) * rec);
}

// initial capacity tokens and signal state of every (link, replica) / (node, replica): thread index = entity * R + replica
__global__ void k_init_dyn(const DevWorld *worlds, int R, const double *co, const double *ci, const int *ph, const double *st, const double *fc) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)(worlds[0].L + worlds[0].N) * R) return;
    const DevWorld &W = worlds[(int)(i % R)];
    long long e = i / R;
    if (e < W.L) { LR(W, e).cap_out_rem = co[e]; LR(W, e).cap_in_rem = ci[e]; }
    else if (e < (long long)W.L + W.N) { long long n = e - W.L; NS(W, n).sig_phase = ph[n]; NR(W, n).sig_t = st[n]; NR(W, n).fc_rem = fc[n]; }
}

// the non-zero initial values of the replica-major state of every replica (blockIdx.y): "none" markers and travel times of -1
__global__ void k_init_rep(const DevWorld *worlds) {
    const DevWorld &W = worlds[blockIdx.y];
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < W.P) {
        W.v_link[i] = -1; W.v_arr_ts[i] = -1; W.v_tt[i] = -1.0;
        if (W.log_mode) { W.v_gen_ts[i] = -1; W.v_end_ts[i] = -1; }
    }
    if (i < (long long)W.D * W.N) W.rnext[i] = -1;
}

// Link::ensure_traveltime_real (traffi.cpp:631-648): position i takes the value of the latest event whose entry step <= i.
UX_DEV void tt_actual_link(const DevWorld &W, int l, double *out) {
    double cur = W.l_tt0[l];  // the fill value of Link::Link (traffi.cpp:519): free-flow time at the speed the link was frozen with
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
    int hd = LQ(W, l).head[cur], n = LQ(W, l).tail - hd;
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
// Lanes = the 32 replicas of ONE link (grid y = lane groups): the event tables and curves are replica-minor, so every load of a warp is one
// contiguous row (a lane per link of one replica read a 4-byte word per 32-byte sector: 40 ms at 2 048 replicas of the Chicago sketch; this form: 4 ms).
__global__ void k_ensemble_partial(const DevWorld *worlds, double *part, int R) {
    const int idx = (int)(blockIdx.x * blockDim.x + threadIdx.x);
#ifdef UX_EMU
    const int l = idx, ln = 0, r = (int)blockIdx.y;  // (the emulation runs one replica per grid row)
    const DevWorld &W = worlds[r];
#else
    const int l = idx >> 5, ln = idx & 31, r = (int)blockIdx.y * 32 + ln;
    const DevWorld &W = worlds[(int)blockIdx.y * 32];  // the group's first replica; the lanes add their offsets
#endif
    int T = W.T, L = W.L;
    if (l >= L || r >= R) return;
    double cur = W.l_tt0[l], s1 = 0, s2 = 0;
    int best = 0;
#pragma unroll 8
    for (int i = 0; i < T; i++) {
        int ord = W.ev_ord[TLI(W, i, l) + ln];
        double v = W.ev_tt[TLI(W, i, l) + ln];
        if (ord > best) { best = ord; cur = v; }
        s1 += cur; s2 += cur * cur;
    }
    double *o = part + ((size_t)r * L + l) * 4;
    const LinkQ *lq = &LQ(W, l) + ln;
    o[0] = T > 0 ? W.cum_arr[TLI(W, (T - 1), l) + ln] : 0.0; o[1] = s1; o[2] = s2;
    o[3] = (double)(lq->tail - lq->head[*W.t_base & 1]);
}
__global__ void k_ensemble_stats(const DevWorld *worlds, int R, const double *part, double *out) {
    int L = worlds[0].L;
    int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= L * 4) return;
    double a = 0;
    for (int r = 0; r < R; r++) a += part[(size_t)r * L * 4 + i];  // fixed replica order
    out[i] = a;
}


// ---------------------------------------------------------------------------------------------------
// Edie states: Analyzer.compute_accurate_traj + compute_edie_state (analyzer.py:205-330) from the RUN
// records.  Contributions are accumulated in 2^-24 fixed point with integer atomics: the sums are independent of the order.
// ---------------------------------------------------------------------------------------------------
#define EDIE_FX 16777216.0
UX_DEV unsigned long long edie_fx(double v) { return (unsigned long long)(long long)floor(v * EDIE_FX + 0.5); }
// Python's float % and // for a positive divisor (floatobject.c float_divmod)
UX_DEV double py_mod(double x, double y) { double m = fmod(x, y); if (m != 0.0 && ((y < 0) != (m < 0))) m += y; return m; }
UX_DEV double py_floordiv(double x, double y) {
    double m = fmod(x, y), div = (x - m) / y;
    if (m != 0.0 && ((y < 0) != (m < 0))) div -= 1.0;
    if (div == 0.0) return 0.0;
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return fl;
}
UX_DEV void edie_add(unsigned long long *cell, unsigned long long v) {
#ifdef UX_EMU
    *cell += v;
#else
    atomicAdd(cell, v);
#endif
}
// one trajectory segment (x0,t0) -> (x1,t1) on a link with an nt x nx grid of dt x dx cells (analyzer.py:272-312)
UX_DEV void edie_segment(unsigned long long *tn, unsigned long long *dn, int nt, int nx, double dt, double dx, double x0, double t0, double x1, double t1) {
    double v0 = (t1 - t0 != 0) ? (x1 - x0) / (t1 - t0) : 0.0;
    int tt = (int)py_floordiv(t0, dt), xx = (int)py_floordiv(x0, dx);
    if (tt < 0) tt += nt;  // Python's negative index
    if (tt < 0 || !(tt < nt && xx < nx)) return;
    if (v0 > 0) {
        double xl0 = dx - py_mod(x0, dx), xl1 = py_mod(x1, dx), tl0 = xl0 / v0, tl1 = xl1 / v0;
        double x1c = py_floordiv(x1, dx);
        if ((double)xx == x1c) { edie_add(dn + (size_t)tt * nx + xx, edie_fx(x1 - x0)); edie_add(tn + (size_t)tt * nx + xx, edie_fx(t1 - t0)); }
        else {
            int jj = (int)(x1c - xx + 1);
            for (int j = 0; j < jj; j++) {
                if (xx + j >= nx) continue;
                size_t c = (size_t)tt * nx + xx + j;
                if (j == 0) { edie_add(dn + c, edie_fx(xl0)); edie_add(tn + c, edie_fx(tl0)); }
                else if (j == jj - 1) { edie_add(dn + c, edie_fx(xl1)); edie_add(tn + c, edie_fx(tl1)); }
                else { edie_add(dn + c, edie_fx(dx)); edie_add(tn + c, edie_fx(dx / v0)); }
            }
        }
    } else edie_add(tn + (size_t)tt * nx + xx, edie_fx(t1 - t0));
}

// RUN records -> per-platoon time order (platoon id, link, position)
__global__ void k_edie_order(const DevWorld *worlds, int replica, long long n, const long long *off, int *o_vid, int *o_link, double *o_x) {
    const DevWorld &W = worlds[replica];
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int vid = W.lg_vid[i];
    long long d = off[vid] + (W.lg_t[i] - W.v_gen_ts[vid]);
    o_vid[d] = vid; o_link[d] = W.lg_link[i]; o_x[d] = W.lg_x[i];
}

// thread per ordered record j: the segment to the platoon's next record on the same link, plus the free-flow extrapolation
// of the first / last record of a link traversal (compute_accurate_traj, analyzer.py:233-246)
__global__ void k_edie_accumulate(const DevWorld *worlds, int replica, long long n, const long long *off, const int *o_vid, const int *o_link, const double *o_x,
                                  const long long *cell_off, const int *cell_nx, const double *edie_dt, const double *edie_dx,
                                  unsigned long long *tn, unsigned long long *dn) {
    const DevWorld &W = worlds[replica];
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int vid = o_vid[j], l = o_link[j];
    long long a = off[vid], b = off[vid + 1];
    int nx = cell_nx[l];
    if (nx <= 0) return;
    int nt = (int)((cell_off[l + 1] - cell_off[l]) / nx);
    if (nt <= 0) return;
    unsigned long long *ltn = tn + cell_off[l], *ldn = dn + cell_off[l];
    double dt = edie_dt[l], dx = edie_dx[l], u = W.l_vmax[l], len = W.l_length[l];
    double x = o_x[j], t = (double)(W.v_gen_ts[vid] + (int)(j - a)) * W.dt;  // log_t_at: (first step + k) * delta_t
    bool first = j == a || o_link[j - 1] != l, last = j + 1 == b || o_link[j + 1] != l;
    if (first && x != 0 && x / u > W.dt * 0.01) edie_segment(ltn, ldn, nt, nx, dt, dx, 0.0, t - x / u, x, t);
    if (!last) edie_segment(ltn, ldn, nt, nx, dt, dx, x, t, o_x[j + 1], (double)(W.v_gen_ts[vid] + (int)(j + 1 - a)) * W.dt);
    else if (len - u * W.dt <= x && x < len) {
        double rem = len - x;
        if (rem / u > W.dt * 0.01) edie_segment(ltn, ldn, nt, nx, dt, dx, x, t, len, t + rem / u);
    }
}

__global__ void k_edie_finish(long long cells, const unsigned long long *tn, const unsigned long long *dn, double dn_scale, double *out_tn, double *out_dn) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    out_tn[c] = (double)(long long)tn[c] / EDIE_FX * dn_scale; out_dn[c] = (double)(long long)dn[c] / EDIE_FX * dn_scale;
}

// Analyzer.od_analysis (analyzer.py:121-181) as a reduction: one thread per OD pair walks the pair's platoons in id order
// (pair_off / pair_veh: the static CSR of the demand) -- counts, then mean and population standard deviation of the travel
// time of the completed trips and of the distance travelled by all of them (two passes, as numpy's std).  `vx` = the platoons'
// current positions (k_veh_snapshot): a RUN platoon's distance includes the part of its current link (oracle R4).
__global__ void k_od_stats(const DevWorld *worlds, int replica, int K, const long long *pair_off, const int *pair_veh, const double *vx, double *out) {
    const DevWorld &W = worlds[replica];
    int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (k >= K) return;
    long long a = pair_off[k], b = pair_off[k + 1];
    double n = 0.0, nc = 0.0, stt = 0.0, sd = 0.0;
    for (long long q = a; q < b; q++) {
        int v = pair_veh[q];
        double d = W.v_dist[v] + (W.v_state[v] == UX_VS_RUN ? vx[v] - W.v_entry_x[v] : 0.0), tt = W.v_tt[v];
        n += 1.0; sd += d;
        if (tt != -1.0) { nc += 1.0; stt += tt; }
    }
    double md = n > 0.0 ? sd / n : 0.0, mt = nc > 0.0 ? stt / nc : 0.0, vd = 0.0, vt = 0.0;
    for (long long q = a; q < b; q++) {
        int v = pair_veh[q];
        double d = W.v_dist[v] + (W.v_state[v] == UX_VS_RUN ? vx[v] - W.v_entry_x[v] : 0.0), tt = W.v_tt[v];
        vd += (d - md) * (d - md);
        if (tt != -1.0) vt += (tt - mt) * (tt - mt);
    }
    double *o = out + (size_t)k * 6;
    o[0] = n; o[1] = nc; o[2] = nc > 0.0 ? mt : -1.0; o[3] = nc > 0.0 ? sqrt(vt / nc) : -1.0; o[4] = md; o[5] = n > 0.0 ? sqrt(vd / n) : 0.0;
}

// Analyzer.link_analysis_coarse (analyzer.py:183-203): one thread per link -- volume and remaining vehicles from the last row
// of the cumulative curves, mean and population std of the positive entries of traveltime_actual, which is materialised on the
// fly from the event tables (running arg-max as in tt_actual_link), twice (mean, then squared deviations).
__global__ void k_linkc_stats(const DevWorld *worlds, int replica, double *out) {
    const DevWorld &W = worlds[replica];
    int l = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (l >= W.L) return;
    double vol = W.cum_dep[TLI(W, W.T - 1, l)], arr = W.cum_arr[TLI(W, W.T - 1, l)];
    double mean = -1.0, sdev = -1.0;
    if (vol != 0.0) {
        double cur = W.l_tt0[l], s = 0.0, n = 0.0;
        int best = 0;
        for (int i = 0; i < W.T; i++) {
            int ord = W.ev_ord[TLI(W, i, l)];
            double v = W.ev_tt[TLI(W, i, l)];
            if (ord > best) { best = ord; cur = v; }
            if (cur > 0.0) { s += cur; n += 1.0; }
        }
        if (n > 0.0) {
            mean = s / n;
            double q = 0.0;
            cur = W.l_tt0[l]; best = 0;
            for (int i = 0; i < W.T; i++) {
                int ord = W.ev_ord[TLI(W, i, l)];
                double v = W.ev_tt[TLI(W, i, l)];
                if (ord > best) { best = ord; cur = v; }
                if (cur > 0.0) q += (cur - mean) * (cur - mean);
            }
            sdev = sqrt(q / n);
        }
    }
    double *o = out + (size_t)l * 4;
    o[0] = vol; o[1] = arr - vol; o[2] = mean; o[3] = sdev;
}
