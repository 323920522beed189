// ux_kernels_dta.cuh -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// DTA route swap on the device.

// =====================================================================================================
// DTA solver batch operations: route swap of dta_solver.cpp:130-373 on the device.
// Every platoon evaluates the cost of each route of its OD pair on the finished run's traveltime_actual table --
// independent work per (platoon, route), serial only along a route (the clock advances link by link).
// =====================================================================================================
struct DtaDev {
    const double *tt;   // [L][T] traveltime_actual (k_tt_actual)
    const double *ext;  // [L][T] congestion externality (DSO) or nullptr
    const int *od_o, *od_d, *route_links; const long long *od_off, *route_off; long long n_od;  // route-set store, pairs sorted
    const long long *tr_off; int *tr_link, *tr_t;  // traversed routes (CSR per platoon): link, entry step
    int mode, has_external, replica;
    const unsigned char *flag_swap;
    int *pair; unsigned char *draws;       // k_dta_prepare -> host RNG -> k_dta_swap
    double *cost_actual, *t_gap; int *choice; unsigned char *pot;
};

// traveltime_actual tables of several replicas in one launch (blockIdx.y = entry of dd)
__global__ void k_dta_tt_actual(const DevWorld *worlds, const DtaDev *dd) {
    const DtaDev &D = dd[blockIdx.y];
    const DevWorld &W = worlds[D.replica];
    int l = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (l < W.L) tt_actual_link(W, l, const_cast<double *>(D.tt));
}

// link-entry events -> per-platoon CSR (position = offset + ordinal of the link on the platoon's route)
__global__ void k_dta_scatter(const DevWorld *worlds, const DtaDev *dd) {
    const DtaDev &D = dd[blockIdx.y];
    const DevWorld &W = worlds[D.replica];
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)*W.tv_count) return;
    int4 e = W.tv_ev[i];
    long long q = D.tr_off[e.x] + e.y;
    D.tr_link[q] = e.z; D.tr_t[q] = e.w;
}

// Link::estimate_congestion_externality (traffi.cpp:650-689) for every (link, step): thread per cell, [L][T]
__global__ void k_dta_externality(const DevWorld *worlds, const DtaDev *dd) {
    const DtaDev &D = dd[blockIdx.y];
    const DevWorld &W = worlds[D.replica];
    const double *tt = D.tt; double *ext = const_cast<double *>(D.ext);
    int n = W.T, L = W.L;
    long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (long long)L * n) return;
    int l = (int)(cell / n), ts = (int)(cell % n);
    const double *ttl = tt + (size_t)l * n;
    double dt = W.dt, fftt = W.l_length[l] / W.l_vmax[l] * 1.01;
    int ts_end = ts;
    for (int i = ts; i < n; i++) {  // congestion tail (forward scan)
        double tt_i = ttl[i];
        int fut = i + (int)(tt_i / dt);
        if (fut >= n) break;
        if (tt_i > fftt && W.cum_arr[TLI(W, fut, l)] - W.cum_dep[TLI(W, fut, l)] > 0) ts_end = i; else break;
    }
    double veh_count = W.cum_arr[TLI(W, ts_end, l)] - W.cum_arr[TLI(W, ts, l)];
    int ts_out = (int)(ts + (int)(ttl[ts] / dt)) + 1;
    if (ts_out >= n) ts_out = n - 1;
    int ts_out_leader = ts_out;
    double dep_out = W.cum_dep[TLI(W, ts_out, l)];
    for (int i = ts_out; i > 0; i--) if (W.cum_dep[TLI(W, i, l)] < dep_out) { ts_out_leader = i; break; }  // leader headway (backward scan)
    ext[cell] = (double)(ts_out - ts_out_leader) * dt * veh_count;
}

UX_DEV double dta_actual_tt(const DevWorld &W, const DtaDev &D, int l, double t) {  // dta_get_actual_travel_time dta_solver.h:56-65
    int n = W.T;
    if (n == 0) return W.l_length[l] / W.l_vmax[l];
    int idx = (int)(t / W.dt);
    if (idx < 0) idx = 0;
    if (idx >= n) idx = n - 1;
    return D.tt[(size_t)l * n + idx];
}
UX_DEV double dta_toll(const DevWorld &W, int l, double t) {  // dta_get_toll dta_solver.h:44-53 (series cover all T steps)
    if (W.l_toll) { int ts = (int)(t / W.dt); if (ts < 0) ts = 0; if (ts >= W.T) ts = W.T - 1; return W.l_toll[(size_t)ts * W.L + l]; }
    return W.l_penalty[l];
}
// dta_compute_route_cost_due / _dso (dta_solver.cpp:72-104)
UX_DEV double dta_route_cost(const DevWorld &W, const DtaDev &D, const int *links, int n, double departure_time) {
    double tt_total = 0.0, extra = 0.0, t = departure_time;
    for (int j = 0; j < n; j++) {
        int l = links[j];
        double ltt = dta_actual_tt(W, D, l, t);
        if (D.mode == 0) extra += dta_toll(W, l, t);
        else { int ts = (int)(t / W.dt); extra += (ts >= 0 && ts < W.T) ? D.ext[(size_t)l * W.T + ts] : 0.0; }
        tt_total += ltt; t += ltt;
    }
    return tt_total + extra;
}

// per platoon: travel cost, and whether it takes part in the swap (finished trip, OD pair in the store, traversed route in
// an external set): the ones that do consume one uniform draw each, in platoon order (dta_solver.cpp:149-203)
__global__ void k_dta_prepare(const DevWorld *worlds, const DtaDev *dd) {
    const DtaDev &D = dd[blockIdx.y];
    const DevWorld &W = worlds[D.replica];
    long long vi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vi >= W.P) return;
    long long a0 = D.tr_off[vi]; int ntr = (int)(D.tr_off[vi + 1] - a0);
    int st = W.v_state[vi], arr = W.v_arr_ts[vi];
    double arrival = (st == UX_VS_END && arr >= 0) ? (double)arr * W.dt : -1.0;
    D.cost_actual[vi] = ntr > 0 ? arrival - (double)D.tr_t[a0] * W.dt : 0.0;
    int choice = -2, pair = -1; unsigned char draws = 0;
    if (st == UX_VS_END) {  // dta_solver.cpp:168 / 308: every other state -- aborted trips included -- keeps no route and draws nothing
        choice = -1;
        int o = W.v_orig[vi], d = W.v_dest[vi];
        long long lo = 0, hi = D.n_od;
        while (lo < hi) { long long mid = (lo + hi) >> 1; int mo = D.od_o[mid]; if (mo < o || (mo == o && D.od_d[mid] < d)) lo = mid + 1; else hi = mid; }
        if (lo < D.n_od && D.od_o[lo] == o && D.od_d[lo] == d) {
            long long r0 = D.od_off[lo], r1 = D.od_off[lo + 1];
            bool forced = false;
            if (D.has_external) {
                bool found = false;
                for (long long r = r0; r < r1 && !found; r++) {
                    long long b = D.route_off[r];
                    if (D.route_off[r + 1] - b != ntr) continue;
                    bool same = true;
                    for (int j = 0; j < ntr && same; j++) same = D.route_links[b + j] == D.tr_link[a0 + j];
                    found = same;
                }
                if (!found && r1 > r0) { choice = (int)r0; forced = true; }
            }
            if (!forced) { pair = (int)lo; draws = 1; }
        }
    }
    D.choice[vi] = choice; D.pair[vi] = pair; D.draws[vi] = draws; D.t_gap[vi] = 0.0; D.pot[vi] = 0;
}

__global__ void k_fill_f64(double *p, long long n, double v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// routes_specified of one replica as a device CSR: length per platoon, then (after the host's scan) the link lists
__global__ void k_dta_rs_len(const DevWorld *worlds, const DtaDev *dd, int *len_out) {
    const DtaDev &D = dd[blockIdx.y];
    const DevWorld &W = worlds[D.replica];
    long long vi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vi >= W.P) return;
    int c = D.choice[vi];
    long long len = c == -2 ? 0 : c == -1 ? D.tr_off[vi + 1] - D.tr_off[vi] : D.route_off[c + 1] - D.route_off[c];
    len_out[(size_t)blockIdx.y * W.P + vi] = (int)len;
}
struct DtaGather { const int *off; int *links; };
__global__ void k_dta_rs_gather(const DevWorld *worlds, const DtaDev *dd, const DtaGather *gg) {
    const DtaDev &D = dd[blockIdx.y];
    const DtaGather &G = gg[blockIdx.y];
    const DevWorld &W = worlds[D.replica];
    long long vi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;  // 8 threads per platoon
    int sub = (int)(threadIdx.x & 7);
    if (vi >= W.P) return;
    int c = D.choice[vi];
    if (c == -2) return;
    const int *src = c == -1 ? D.tr_link + D.tr_off[vi] : D.route_links + D.route_off[c];
    int o = G.off[vi], len = G.off[vi + 1] - o;
    for (int j = sub; j < len; j += 8) G.links[o + j] = src[j];
}

// exclusive scan of n ints (block per job): out[0] = 0 ... out[n] = total; used for the traversal counts and the route lengths
struct DtaScanJob { const int *in; void *out; long long n; int out_is_i64; long long *total; };
__global__ void __launch_bounds__(1024) k_dta_scan(const DtaScanJob *jobs) {
    const DtaScanJob J = jobs[blockIdx.x];
    long long *o64 = (long long *)J.out; int *o32 = (int *)J.out;
#ifdef UX_EMU
    if (threadIdx.x) return;
    long long acc = 0;
    for (long long i = 0; i < J.n; i++) { if (J.out_is_i64) o64[i] = acc; else o32[i] = (int)acc; acc += J.in[i]; }
    if (J.out_is_i64) o64[J.n] = acc; else o32[J.n] = (int)acc;
    if (J.total) *J.total = acc;
#else
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    int tid = (int)threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < J.n; base += 1024) {
        long long i = base + tid;
        long long v = i < J.n ? (long long)J.in[i] : 0, x = v;
        for (int d = 1; d < 32; d <<= 1) { long long y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            long long t = warp_tot[lane], u = t;
            for (int d = 1; d < 32; d <<= 1) { long long y = __shfl_up_sync(0xffffffffu, u, d); if (lane >= d) u += y; }
            warp_tot[lane] = u - t;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        long long excl = carry_s + warp_tot[wid] + x - v;
        if (i < J.n) { if (J.out_is_i64) o64[i] = excl; else o32[i] = (int)excl; }
        __syncthreads();
        if (tid == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) { if (J.out_is_i64) o64[J.n] = carry_s; else o32[J.n] = (int)carry_s; if (J.total) *J.total = carry_s; }
#endif
}

// number of recorded link-entry events of replicas first .. first+n-1 (one small transfer instead of n)
__global__ void k_dta_counts(const DevWorld *worlds, int first, int n, unsigned long long *out) {
    int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i < n) out[i] = *worlds[first + i].tv_count;
}

// RouteSwapResult totals of one replica, accumulated in platoon order exactly as the reference's loop does
// (dta_solver.cpp:243-252: total_t_gap is a sequential FP64 sum): one thread per replica.
__global__ void k_dta_totals(const DevWorld *worlds, const DtaDev *dd, double *out) {
    if (threadIdx.x) return;
    const DtaDev &D = dd[blockIdx.x];
    const DevWorld &W = worlds[D.replica];
    double n_swap = 0, pot_n = 0, gap = 0;
    for (long long i = 0; i < W.P; i++) {
        bool pt = D.pot[i] != 0;
        if (pt) pot_n += W.dn;
        gap += D.t_gap[i];
        if (D.choice[i] >= 0 && pt) n_swap += W.dn;  // a forced external route has pot == 0
    }
    out[blockIdx.x * 3 + 0] = n_swap; out[blockIdx.x * 3 + 1] = pot_n; out[blockIdx.x * 3 + 2] = gap;
}

#define DTA_WARPS 4
// warp per platoon: lanes evaluate the alternative routes, then the scan over routes in store order (dta_solver.cpp:228-241)
__global__ void __launch_bounds__(32 * DTA_WARPS) k_dta_swap(const DevWorld *worlds, const DtaDev *dd) {
    const DtaDev &D = dd[blockIdx.y];
    LOAD_WORLD_N(W, worlds, 64, D.replica)
    WARP_BEGIN
    __shared__ double cost_s[DTA_WARPS][32];
    long long vi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int wib = (int)((threadIdx.x >> 5) % DTA_WARPS);
    if (vi >= W.P) return;
    int pair = D.pair[vi];
    if (pair < 0) return;
    long long a0 = D.tr_off[vi]; int ntr = (int)(D.tr_off[vi + 1] - a0);
    const int *tr = D.tr_link + a0;
    double departure_t = ntr > 0 ? (double)D.tr_t[a0] * W.dt : (double)W.v_depart_ts[vi] * W.dt;
    long long r0 = D.od_off[pair], r1 = D.od_off[pair + 1];
    double cost_current = 0.0, t_gap = 0.0; int changed_idx = -1; bool pot = false;
    bool flag_swap = D.flag_swap[vi] != 0;
    for (long long base = r0 - 1; base < r1; base += 32) {  // slot r0-1 is the traversed route itself
        WARP_FOR(lane) {
            long long r = base + lane;
            double c = 0.0;
            if (r == r0 - 1) {
                if (D.mode == 0) {  // DUE: tolls at the actual entry times (dta_solver.cpp:209-221)
                    double t = departure_t;
                    for (int j = 0; j < ntr; j++) { double ltt = dta_actual_tt(W, D, tr[j], t); c += dta_toll(W, tr[j], (double)D.tr_t[a0 + j] * W.dt); c += ltt; t += ltt; }
                } else c = dta_route_cost(W, D, tr, ntr, departure_t);
            } else if (r < r1) c = dta_route_cost(W, D, D.route_links + D.route_off[r], (int)(D.route_off[r + 1] - D.route_off[r]), departure_t);
            cost_s[wib][lane] = c;
        }
        WARP_SYNC();
        WARP_FOR(lane) if (lane == 0) {
            for (int q = 0; q < 32 && base + q < r1; q++) {
                double c = cost_s[wib][q];
                if (base + q == r0 - 1) { cost_current = c; continue; }
                if (c < cost_current) { t_gap = cost_current - c; pot = true; if (flag_swap) { changed_idx = (int)(base + q); cost_current = c; } }
            }
        }
        WARP_SYNC();
    }
    WARP_FOR(lane) if (lane == 0) { D.t_gap[vi] = t_gap; D.pot[vi] = pot ? 1 : 0; if (changed_idx >= 0) D.choice[vi] = changed_idx; }
}

