// engine.cu -- B200-native engine for UXsim's per-timestep simulation loop (World.exec_simulation).
//
// One translation unit: sm_100a CUDA kernels + the C++ host driver + the C-ABI of include/uxsim_b200.h.
// Reference behaviour: toruseo/UXsim uxsim/trafficpp/traffi.cpp (C++ engine) == uxsim/uxsim.py (Python
// engine); each kernel cites the reference functions it implements.  Design notes: DESIGN.md.
//
// Data layout (all FP64 where the reference is FP64; FMA contraction disabled with -fmad=false so every
// compare that decides an integer outcome sees the reference's values):
//   * every link owns a ring of `cap` slots inside one big SoA arena: X[2][C] (positions, double-buffered by
//     timestep parity: X[t&1] = position at the start of step t, X[(t+1)&1] = position before the last move,
//     i.e. the reference's x_old), V[C], VID[C], LANE[C].  A platoon's leader is the slot `lanes` positions
//     ahead in the same ring (traffi.cpp:175-181), so car-following is a coalesced read of X[p] and X[p-lanes].
//   * per-link queue = (head[2], tail) monotone counters; head is double-buffered like X because the update
//     kernel both reads it (all slots) and advances it (trip ends).
//   * time series are time-major [T][L] so one step writes one contiguous row.
//   * R replicas (seeds) = R DevWorld structs sharing the static network arrays; blockIdx.y = replica.
//
// Timestep = 2 kernels (+ route update every DELTAT_ROUTE steps), replayed from CUDA graphs:
//   k_node     warp per node : Link::update of the link sides the node owns, Node::generate, signal, node capacity,
//              then Node::transfer; all dependency levels in one launch (see DESIGN.md "transfer order")
//   k_update   warp per (link, part) : car_follow_newell + Vehicle::update for the occupied positions; the part-0
//              warp also handles the link head (trip end, abort, next-link choice) and the per-step log records
//   k_edge_cost / k_sssp / k_duo(+finish) : update_adj_time_matrix, route_search_all, route_choice_duo
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <mutex>
#include <chrono>
#include <algorithm>
#include <random>
#include <queue>
#include <thread>
#include <atomic>
#include <memory>

#include "../../include/uxsim_b200.h"
#include "../../include/ux_rng.h"

#ifdef UX_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define UX_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

// ---------------------------------------------------------------------------------------------------
// limits
// ---------------------------------------------------------------------------------------------------
#define UX_MAXDEG 16   // max out-/in-degree of a node
#define UX_MAXC 64     // max sum of lanes over the in-links (or the out-links) of one node
#define UX_SEG 32      // ring slots per update-kernel work item (one warp)

enum { LF_SEES_POPS = 1, LF_END_LT_START = 2 };            // link flags
enum { VF_WAIT_END = 1, VF_ABORT = 2, VF_CAND = 4 };       // vehicle flags
enum { DERR_NONE = 0, DERR_RING_FULL = 1, DERR_ROUTE_EXHAUSTED = 2, DERR_ROUTE_INCONSISTENT = 3, DERR_AVOID_ALL = 4, DERR_LOG_FULL = 5 };

// ---------------------------------------------------------------------------------------------------
// device world
// ---------------------------------------------------------------------------------------------------
struct DevWorld {
    int N, L, T, D, vis_words;
    long long P, C;
    int nseg;
    double dn, dt, duo_w, duo_time, noise;
    int itt, route_ts, hard_det, gradual, no_cyclic, log_mode;
    long long seed;
    // static network (shared by all replicas)
    const int *l_start, *l_end, *l_lanes, *l_rcap, *l_flags, *l_sg_off, *l_sg, *l_pair_first, *l_pair_off, *l_pair_list;
    const long long *l_roff;
    const double *l_length, *l_vmax, *l_dpl_dn, *l_udt, *l_cap_out, *l_cap_in, *l_mp, *l_penalty, *l_toll;
    const int *n_out_off, *n_out, *n_in_off, *n_in, *n_sig_off, *n_lanes, *n_out_lanes, *lvl_off, *lvl_nodes, *lvl_level, *dep_off, *dep_list;
    int n_levels;
    const double *n_sig, *n_fc;
    const int *v_orig, *v_dest, *v_depart_ts, *gen_off, *gen_list, *gen_ts, *dest_row, *row_dest;
    const int *vl_off[3], *vl_list[3];  // per-vehicle CSR: 0 specified_route, 1 links_preferred, 2 links_avoid
    const int *seg_link, *seg_start;
    const int *e_from, *e_to, *e_link;  // node-pair edges (leading link of each pair)
    const int *ein_off, *ein_from, *ein_eid;  // the same edges grouped by their END node (CSR), for the frontier search
    int n_edges;
    // dynamic (per replica)
    int *head, *tail, *tail_k1, *pop_k2, *hcount, *trips, *ev_cnt;
    long long *stat_count;
    double *cap_out_rem, *cap_in_rem;
    double *X, *V;
    int *VID, *LANE;
    int *sig_phase, *gen_head, *gen_avail;
    double *sig_t, *fc_rem;
    int *v_state, *v_next, *v_link, *v_arr_ts, *v_trav;
    double *v_mr, *v_atl, *v_tt, *v_dist, *v_entry_x;
    unsigned char *v_flags;
    // per-platoon per-step RUN records (Vehicle::log_data, traffi.cpp:1060-1110), appended in arrival order
    int *v_gen_ts, *v_end_ts; unsigned char *v_end_kind;  // step of generation / of trip end, 1 = ended in update, 2 = in transfer
    int *lg_vid, *lg_t, *lg_link, *lg_lane; double *lg_x, *lg_v, *lg_s;
    unsigned long long *lg_count; long long lg_cap;
    unsigned *v_visited;
    double *cum_arr, *cum_dep, *tt_inst, *ev_tt;
    int *ev_ord, *sig_log;
    double *pref, *rdist, *pair_cost;
    int *pref_active, *has_path, *rnext;
    // link-entry events (platoon, ordinal of the link on its route, link, step) in arrival order; only with UX_OPT_RECORD_TRAVERSALS
    int4 *tv_ev; unsigned long long *tv_count; long long tv_cap;
    int *node_done;  // per node: (timestep + 1) of the last finished transfer -- monotone, so it never needs a reset
    int *err;     // [0] code, [1] info
    int *t_base;  // timestep at the start of the running graph
};

#define UX_DEV __device__ __forceinline__

// The per-replica DevWorld lives in global memory; reading its ~90 pointers through a reference there would make
// every `W.x[...]` a dependent pointer load that must be repeated after each global store (possible aliasing).
// Each block therefore copies its replica's struct into shared memory once; shared memory cannot alias global
// stores, so the pointers stay in registers / LDS and the kernels' dependent-load chains halve.
#ifdef UX_EMU
#define LOAD_WORLD(W, worlds) static DevWorld W##_sh; W##_sh = (worlds)[blockIdx.y]; const DevWorld &W = W##_sh;
#define LOAD_WORLD_N(W, worlds, NT, REPLICA) static DevWorld W##_sh; W##_sh = (worlds)[REPLICA]; const DevWorld &W = W##_sh;
#else
#define LOAD_WORLD_N(W, worlds, NT, REPLICA)  /* the first NT threads of the block copy (every block has >= NT threads) */ \
    __shared__ DevWorld W##_sh;                                                                          \
    {                                                                                                    \
        const int *src_ = reinterpret_cast<const int *>(&(worlds)[REPLICA]);                          \
        int *dst_ = reinterpret_cast<int *>(&W##_sh);                                                    \
        _Pragma("unroll") for (int k_ = 0; k_ < (int)((sizeof(DevWorld) / 4 + (NT) - 1) / (NT)); k_++) {   \
            int i_ = (int)threadIdx.x + k_ * (NT);                                                       \
            if (threadIdx.x < (NT) && i_ < (int)(sizeof(DevWorld) / 4)) dst_[i_] = src_[i_];              \
        }                                                                                                \
    }                                                                                                    \
    __syncthreads();                                                                                     \
    const DevWorld &W = W##_sh;
#define LOAD_WORLD(W, worlds) LOAD_WORLD_N(W, worlds, 64, blockIdx.y)
#endif

// Warp-cooperative kernels are written as phases.  On the GPU WARP_FOR runs its body once with `lane` = the lane
// id and WARP_SYNC is __syncwarp (execution + shared-memory ordering); in the host emulation build one thread per
// warp runs every phase as a loop over the 32 lanes, which is equivalent because no phase depends on the order
// in which its lanes execute.
#ifdef UX_EMU
#define WARP_BEGIN if (threadIdx.x & 31) return;
#define WARP_FOR(lane) for (int lane = 0; lane < 32; lane++)
#define WARP_SYNC()
#else
#define WARP_BEGIN
#define WARP_FOR(lane) for (int lane = (int)(threadIdx.x & 31), once_ = 1; once_; once_ = 0)
#define WARP_SYNC() __syncwarp()
#endif

UX_DEV void dev_error(const DevWorld &W, int code, int info) {
    if (atomicCAS(&W.err[0], 0, code) == 0) W.err[1] = info;
}

UX_DEV int ring_slot(int head, int pos, int cap) {  // physical slot of queue position `pos` (cap is a power of two)
    return (int)(((unsigned)head + (unsigned)pos) & (unsigned)(cap - 1));
}

// random_choice (utils.h:62-87) with a supplied uniform u in [0,1)
UX_DEV int weighted_pick(const double *wts, int n, double u) {
    double wsum = 0.0;
    for (int i = 0; i < n; i++) wsum += wts[i];
    if (wsum <= 0.0) { int k = (int)(u * (double)n); return k >= n ? n - 1 : k; }
    double r = u * wsum, accum = 0.0;
    for (int i = 0; i < n; i++) { accum += wts[i]; if (r <= accum) return i; }
    return n - 1;
}

UX_DEV int ld_cg(const int *p) {
#ifdef UX_EMU
    return *p;
#else
    return __ldcg(p);
#endif
}

UX_DEV bool is_inf(double c) { return c > 1.0e300; }

UX_DEV bool list_has(const int *a, int n, int v) {
    for (int i = 0; i < n; i++) if (a[i] == v) return true;
    return false;
}

// Vehicle::route_next_link_choice (traffi.cpp:934-1033).  Returns the chosen link id or -1.
UX_DEV int route_choice(const DevWorld &W, int vid, int node, int t, unsigned entity, unsigned purpose, unsigned draw, int *deferred_err = nullptr) {
    int o0 = W.n_out_off[node], nset = W.n_out_off[node + 1] - o0;
    if (nset == 0) return -1;
    const int *linkset = W.n_out + o0;
    if (W.vl_off[0]) {  // specified_route
        int r0 = W.vl_off[0][vid], len = W.vl_off[0][vid + 1] - r0;
        if (len > 0) {
            int traveled = W.v_trav[vid];
            if (traveled >= len) { if (deferred_err) *deferred_err = DERR_ROUTE_EXHAUSTED; else dev_error(W, DERR_ROUTE_EXHAUSTED, vid); return -1; }
            int forced = W.vl_list[0][r0 + traveled];
            if (!list_has(linkset, nset, forced)) { if (deferred_err) *deferred_err = DERR_ROUTE_INCONSISTENT; else dev_error(W, DERR_ROUTE_INCONSISTENT, vid); return -1; }
            return forced;
        }
    }
    int bufa[UX_MAXDEG], bufb[UX_MAXDEG];
    int *out = bufa, *tmp = bufb, n = 0;
    for (int i = 0; i < nset; i++) out[n++] = linkset[i];
    if (W.vl_off[1]) {  // links_preferred: intersection if non-empty
        int r0 = W.vl_off[1][vid], len = W.vl_off[1][vid + 1] - r0;
        if (len > 0) {
            int m = 0;
            for (int i = 0; i < n; i++) if (list_has(W.vl_list[1] + r0, len, out[i])) tmp[m++] = out[i];
            if (m > 0) { int *s = out; out = tmp; tmp = s; n = m; }
        }
    }
    if (W.vl_off[2]) {  // links_avoid: subtraction, error if nothing is left
        int r0 = W.vl_off[2][vid], len = W.vl_off[2][vid + 1] - r0;
        if (len > 0) {
            int m = 0;
            for (int i = 0; i < n; i++) if (!list_has(W.vl_list[2] + r0, len, out[i])) tmp[m++] = out[i];
            if (m != n) { int *s = out; out = tmp; tmp = s; n = m; }
            if (n == 0) { if (deferred_err) *deferred_err = DERR_AVOID_ALL; else dev_error(W, DERR_AVOID_ALL, vid); return -1; }
        }
    }
    if (W.no_cyclic) {
        const unsigned *vis = W.v_visited + (size_t)vid * W.vis_words;
        int m = 0;
        for (int i = 0; i < n; i++) { int e = W.l_end[out[i]]; if (!((vis[e >> 5] >> (e & 31)) & 1u)) tmp[m++] = out[i]; }
        if (m > 0) { int *s = out; out = tmp; tmp = s; n = m; }
    }
    if (n == 0) return -1;
    double pw[UX_MAXDEG];
    int row = W.dest_row[W.v_dest[vid]];
    for (int i = 0; i < n; i++) pw[i] = row >= 0 ? W.pref[(size_t)row * W.L + out[i]] : 0.0;
    if (W.hard_det) {
        int best = 0;
        for (int i = 1; i < n; i++) if (pw[i] > pw[best]) best = i;
        return out[best];
    }
    double u = ux_rng_uniform(W.seed, (uint32_t)t, entity, purpose, draw);
    return out[weighted_pick(pw, n, u)];
}

// acceptance test of an out-link seen with effective head `hd` (traffi.cpp:130-142, 285-297)
UX_DEV bool can_accept(const DevWorld &W, int l, int hd, int cur) {
    int sz = W.tail[l] - hd, lanes = W.l_lanes[l];
    bool ok = false;
    if (sz < lanes) ok = true;
    else {
        int slot = ring_slot(hd, sz - lanes, W.l_rcap[l]);
        if (W.X[(size_t)cur * W.C + W.l_roff[l] + slot] > W.l_dpl_dn[l]) ok = true;
    }
    if (W.cap_in_rem[l] < W.dn) ok = false;
    return ok;
}

// push a platoon at the tail of link l (lane + leader rule traffi.cpp:168-183 / 385-398).  Returns the slot.
UX_DEV int ring_push(const DevWorld &W, int l, int hd, int cur, double x, double x_old, double v, int vid) {
    int cap = W.l_rcap[l], tl = W.tail[l], sz = tl - hd, lanes = W.l_lanes[l];
    if (sz >= cap) { dev_error(W, DERR_RING_FULL, l); return -1; }
    long long o = W.l_roff[l];
    int lane = 0;
    if (sz > 0) lane = (W.LANE[o + ring_slot(tl - 1, 0, cap)] + 1) % lanes;
    int slot = ring_slot(tl, 0, cap);
    W.X[(size_t)cur * W.C + o + slot] = x;
    W.X[(size_t)(cur ^ 1) * W.C + o + slot] = x_old;
    W.V[o + slot] = v;
    W.VID[o + slot] = vid;
    W.LANE[o + slot] = lane;
    W.tail[l] = tl + 1;
    return slot;
}

// One link-entry event for the DTA route swap (what dta_get_traveled_route reads back out of the per-step logs,
// dta_solver.cpp:30-57: a platoon entering link l in step t has its first log record on l at time t*dt).
UX_DEV void record_traversal(const DevWorld &W, int vid, int k, int l, int t) {
    if (!W.tv_count) return;
    unsigned long long idx = atomicAdd(W.tv_count, 1ull);
    if ((long long)idx >= W.tv_cap) { dev_error(W, DERR_LOG_FULL, l); return; }
    W.tv_ev[idx] = make_int4(vid, k, l, t);
}

UX_DEV void mark_entered(const DevWorld &W, int vid, int l, int t) {  // traffi.cpp:160-166, 372-378
    if (W.no_cyclic) { int s = W.l_start[l]; W.v_visited[(size_t)vid * W.vis_words + (s >> 5)] |= 1u << (s & 31); }
    int k = W.v_trav[vid];
    W.v_trav[vid] = k + 1;
    W.v_link[vid] = l;
    record_traversal(W, vid, k, l, t);
}

// record a link-exit travel time (traffi.cpp:357-362, 893-898) into the "last event per entry step" table
UX_DEV void record_tt_event(const DevWorld &W, int l, double atl, double tt) {
    int s = (int)(atl / W.dt);
    if (s < 0) s = 0;
    if (s >= W.T) s = W.T - 1;
    int ord = W.ev_cnt[l] + 1;
    W.ev_cnt[l] = ord;
    W.ev_ord[(size_t)s * W.L + l] = ord;
    W.ev_tt[(size_t)s * W.L + l] = tt;
}

// Vehicle::end_trip (traffi.cpp:886-927) for the front platoon `vid` of link l, currently at position x
UX_DEV void end_trip(const DevWorld &W, int vid, int l, int t, double x) {
    W.trips[l] += 1;
    W.cum_dep[(size_t)t * W.L + l] += W.dn;
    double atl = W.v_atl[vid];
    record_tt_event(W, l, atl, ((double)t + 1.0) * W.dt - atl);
    W.v_atl[vid] = (double)t * W.dt;
    W.v_dist[vid] += x - W.v_entry_x[vid];
    W.v_link[vid] = -1;
    if (W.v_flags[vid] & VF_ABORT) { W.v_state[vid] = UX_VS_ABORT; W.v_arr_ts[vid] = -1; W.v_tt[vid] = -1.0; }
    else {
        W.v_state[vid] = UX_VS_END;
        W.v_arr_ts[vid] = t;
        W.v_tt[vid] = (double)t * W.dt - (double)W.v_depart_ts[vid] * W.dt;
    }
    W.v_flags[vid] = 0;
}

// Sum of V over the queue of link l in the fixed order "32 interleaved partial sums + butterfly" (oracle R3).
UX_DEV double warp_queue_vsum(const DevWorld &W, int l, int hd, int n, int lane) {
    int cap = W.l_rcap[l];
    const double *V = W.V + W.l_roff[l];
#ifdef UX_EMU
    double part[32];
    for (int k = 0; k < 32; k++) { double s = 0.0; for (int i = k; i < n; i += 32) s += V[ring_slot(hd, i, cap)]; part[k] = s; }
    for (int off = 16; off >= 1; off >>= 1) { double nx[32]; for (int k = 0; k < 32; k++) nx[k] = part[k] + part[k ^ off]; for (int k = 0; k < 32; k++) part[k] = nx[k]; }
    (void)lane;
    return part[0];
#else
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += V[ring_slot(hd, i, cap)];
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    return s;
#endif
}

// ---------------------------------------------------------------------------------------------------
// K1: Link::update (traffi.cpp:542-565, 599-622) for the node's out-links, then Node::generate (105-189),
//     Node::signal_update (194-207), Node::flow_capacity_update (226-234).  One warp per node.
// ---------------------------------------------------------------------------------------------------
template <int MC, int MD>
struct PreSmem {
    int choice[MC], g_vid[MC], g_err[MC];
    int o_hd[MD], o_tail[MD], o_lanes[MD], o_cap[MD], o_lastlane[MD], o_win_off[MD + 1], o_win_n[MD];
    long long o_roff[MD];
    double o_ci[MD], o_dpl[MD], o_vmax[MD], o_cumarr[MD];
    double w_x[MC];
};

template <int MC, int MD>
UX_DEV void pre_node(const DevWorld &W, PreSmem<MC, MD> &G, double *tt_s, int *rel, int n, int t, int lane) {
    int cur = t & 1, L = W.L;
    int o0 = W.n_out_off[n], nout = W.n_out_off[n + 1] - o0;
    int i0 = W.n_in_off[n], nin = W.n_in_off[n + 1] - i0;
    bool fresh = (t % W.itt) == 0;
    // ---- instantaneous travel time of the out-links (whole warp per link, only every `itt` steps)
    if (fresh) {
        for (int k = 0; k < nout; k++) {
            int l = W.n_out[o0 + k];
            int hd = W.head[cur * L + l], nq = W.tail[l] - hd;
            double tt;
            if (nq > 0) {
                double vs = warp_queue_vsum(W, l, hd, nq, lane);
                double avg = vs / (double)nq;
                tt = avg > 0.0 ? W.l_length[l] / avg : W.l_length[l] / (W.l_vmax[l] / 100.0);
            } else tt = W.l_length[l] / W.l_vmax[l];
            if (lane == 0) tt_s[k] = tt;
        }
        WARP_SYNC();
    }
    // ---- scalar link update (Link::update traffi.cpp:542-565).  A link's inflow side (arrival curve, capacity_in
    //      tokens, instantaneous travel time) belongs to its START node, its outflow side (departure curve, capacity_out
    //      tokens, pops of this step) to its END node -- the same ownership the transfer uses, so that this node's whole
    //      timestep (update, generate, signal, transfer) needs nothing from other nodes except the dependency wait.
    WARP_FOR(ln) {
        for (int k = ln; k < nout; k += 32) {
            int l = W.n_out[o0 + k];
            size_t row = (size_t)t * L + l;
            if (fresh) W.tt_inst[row] = tt_s[k];
            else if (t > 0) W.tt_inst[row] = W.tt_inst[row - L];
            if (t != 0) W.cum_arr[row] = W.cum_arr[row - L];
            double ci = W.l_cap_in[l];
            if (ci < 10e9) { if (W.cap_in_rem[l] < W.dn * W.l_lanes[l]) W.cap_in_rem[l] += ci * W.dt; } else W.cap_in_rem[l] = 10e9;
        }
        for (int k = ln; k < nin; k += 32) {
            int l = W.n_in[i0 + k];
            size_t row = (size_t)t * L + l;
            if (t != 0) W.cum_dep[row] = W.cum_dep[row - L];
            double co = W.l_cap_out[l];
            if (co < 10e9) { if (W.cap_out_rem[l] < W.dn * W.l_lanes[l]) W.cap_out_rem[l] += co * W.dt; } else W.cap_out_rem[l] = 10e9;
            W.pop_k2[l] = 0;
        }
    }
    WARP_SYNC();
    // ---- release HOME -> WAIT: platoons with departure step <= t-1 are in the vertical queue (traffi.cpp:836-841).
    //      The per-origin list is sorted by departure step, so the released ones are a prefix: 32 entries per probe.
    int g0 = W.gen_off[n], glen = W.gen_off[n + 1] - g0, av = W.gen_avail[n];
    while (av < glen) {
        WARP_FOR(ln) if (ln == 0) *rel = 0;
        WARP_SYNC();
        WARP_FOR(ln) { int idx = av + ln; if (idx < glen && W.gen_ts[g0 + idx] <= t - 1) atomicAdd(rel, 1); }
        WARP_SYNC();
        int c = *rel;
        WARP_SYNC();
        av += c;
        if (c < 32) break;
    }
    // ---- generate (traffi.cpp:105-189).  The route draws of the first `natt` waiting platoons and the tail state of the
    //      out-links are fetched in parallel; lane 0 then walks the attempts on shared memory and stops at the first refusal.
    int gh = W.gen_head[n];
    int natt = av - gh;
    { int tl_ = W.n_out_lanes[n]; if (natt > tl_) natt = tl_; if (natt > MC) natt = MC; }
    if (nout == 0) natt = 0;
    if (natt > 0) {
        WARP_FOR(ln) {
            for (int i = ln; i < natt; i += 32) {
                int vid = W.gen_list[g0 + gh + i], e = 0;
                G.g_vid[i] = vid;
                G.choice[i] = route_choice(W, vid, n, t, (unsigned)n, UX_RNG_GENERATE, (unsigned)i, &e);
                G.g_err[i] = e;
            }
            for (int q = ln; q < nout; q += 32) {
                int ol = W.n_out[o0 + q];
                int hd = W.head[cur * L + ol], tl = W.tail[ol], lanes = W.l_lanes[ol], cap = W.l_rcap[ol];
                long long o = W.l_roff[ol];
                double ci_ = W.cap_in_rem[ol], dpl_ = W.l_dpl_dn[ol], vm_ = W.l_vmax[ol], ca_ = W.cum_arr[(size_t)t * L + ol];
                int sz = tl - hd;
                int ll_ = sz > 0 ? W.LANE[o + ring_slot(tl - 1, 0, cap)] : -1;
                G.o_hd[q] = hd; G.o_tail[q] = tl; G.o_lanes[q] = lanes; G.o_cap[q] = cap; G.o_roff[q] = o;
                G.o_ci[q] = ci_; G.o_dpl[q] = dpl_; G.o_vmax[q] = vm_; G.o_cumarr[q] = ca_; G.o_lastlane[q] = ll_;
                G.o_win_n[q] = sz < lanes ? sz : lanes;
            }
        }
        WARP_SYNC();
        WARP_FOR(ln) if (ln == 0) { int acc = 0; for (int q = 0; q < nout; q++) { G.o_win_off[q] = acc; acc += G.o_lanes[q]; } G.o_win_off[nout] = acc; }
        WARP_SYNC();
        WARP_FOR(ln) {
            int tot = G.o_win_off[nout];
            for (int e = ln; e < tot && e < MC; e += 32) {
                int q = 0;
                while (G.o_win_off[q + 1] <= e) q++;
                int k = e - G.o_win_off[q];
                if (k < G.o_win_n[q]) {
                    int sz = G.o_tail[q] - G.o_hd[q];
                    G.w_x[e] = W.X[(size_t)cur * W.C + G.o_roff[q] + ring_slot(G.o_hd[q], sz - G.o_win_n[q] + k, G.o_cap[q])];
                }
            }
        }
        WARP_SYNC();
        WARP_FOR(ln) if (ln == 0) {
            for (int i = 0; i < natt; i++) {
                int vid = G.g_vid[i], ol = G.choice[i];
                if (G.g_err[i]) { dev_error(W, G.g_err[i], vid); break; }
                W.v_next[vid] = ol;
                if (ol < 0) break;
                int q = 0;
                while (q < nout && W.n_out[o0 + q] != ol) q++;
                int lanes = G.o_lanes[q], sz = G.o_tail[q] - G.o_hd[q], w0 = G.o_win_off[q];
                bool ok = false;  // traffi.cpp:130-142
                if (sz < lanes) ok = true;
                else if (G.w_x[w0] > G.o_dpl[q]) ok = true;
                if (G.o_ci[q] < W.dn) ok = false;
                if (!ok) break;
                if (sz >= G.o_cap[q]) { dev_error(W, DERR_RING_FULL, ol); break; }
                gh++;
                int lane_new = sz > 0 ? (G.o_lastlane[q] + 1) % lanes : 0;
                int slot = ring_slot(G.o_tail[q], 0, G.o_cap[q]);
                long long o = G.o_roff[q];
                W.X[(size_t)cur * W.C + o + slot] = 0.0;
                W.X[(size_t)(cur ^ 1) * W.C + o + slot] = 0.0;
                W.V[o + slot] = G.o_vmax[q];
                W.VID[o + slot] = vid;
                W.LANE[o + slot] = lane_new;
                G.o_tail[q] += 1; G.o_lastlane[q] = lane_new;
                if (G.o_win_n[q] < lanes) { G.w_x[w0 + G.o_win_n[q]] = 0.0; G.o_win_n[q] += 1; }
                else { for (int e = w0; e < w0 + lanes - 1; e++) G.w_x[e] = G.w_x[e + 1]; G.w_x[w0 + lanes - 1] = 0.0; }
                W.v_state[vid] = UX_VS_RUN;
                if (W.log_mode) W.v_gen_ts[vid] = t;
                W.v_atl[vid] = (double)t * W.dt;
                W.v_entry_x[vid] = 0.0;
                mark_entered(W, vid, ol, t);
                G.o_cumarr[q] += W.dn;
                G.o_ci[q] -= W.dn;
            }
            W.gen_head[n] = gh;
        }
        WARP_SYNC();
        WARP_FOR(ln) {
            for (int q = ln; q < nout; q += 32) {
                int ol = W.n_out[o0 + q];
                W.cap_in_rem[ol] = G.o_ci[q]; W.cum_arr[(size_t)t * L + ol] = G.o_cumarr[q]; W.tail[ol] = G.o_tail[q];
            }
        }
        WARP_SYNC();
    }
    WARP_FOR(ln) { for (int k = ln; k < nout; k += 32) { int l = W.n_out[o0 + k]; W.tail_k1[l] = W.tail[l]; } }
    WARP_FOR(ln) if (ln == 0) {
        W.gen_avail[n] = av;
        // ---- signal (Node::signal_update traffi.cpp:194-207)
        int s0 = W.n_sig_off[n], nsig = W.n_sig_off[n + 1] - s0;
        int ph = W.sig_phase[n];
        if (nsig > 1) {
            double st = W.sig_t[n];
            if (st > W.n_sig[s0 + ph]) { ph++; st = 0; if (ph >= nsig) ph = 0; }
            st += W.dt;
            W.sig_t[n] = st; W.sig_phase[n] = ph;
        }
        W.sig_log[(size_t)t * W.N + n] = ph;
        // ---- node flow capacity (Node::flow_capacity_update traffi.cpp:226-234)
        double fc = W.n_fc[n];
        if (fc >= 0.0) { int nl = W.n_lanes[n]; if (W.fc_rem[n] < W.dn * (nl > 0 ? nl : 1)) W.fc_rem[n] += fc * W.dt; }
        else W.fc_rem[n] = 10e10;
    }
    WARP_SYNC();
}

// ---------------------------------------------------------------------------------------------------
// K2: Node::transfer (traffi.cpp:239-445).  One warp per node of dependency level `level`.
//
// The reference algorithm is inherently serial inside a node (slots are tried one after another and every move
// changes the queues the next slot sees), but everything it READS can be known up front.  So the warp first
// gathers the node's whole neighbourhood into shared memory in parallel (in-link heads, the candidate platoons at
// those heads, and for every requested out-link its tail state plus the window of its last `lanes` platoons),
// lane 0 then runs the serial slot loop purely on shared memory (global stores only, no dependent global loads),
// and the lanes write the per-link accumulators back in parallel.  A dependent chain of a few hundred global
// loads becomes five.
// ---------------------------------------------------------------------------------------------------
template <int MC, int MD>
struct XferSmem {
    // in-links
    int il[MD], hd[MD], hc[MD], cap[MD], popped[MD], pos_off[MD + 1], evc[MD], trips[MD];
    long long roff[MD];
    double co_rem[MD], mp[MD], vmax[MD], cumdep[MD];
    unsigned char sig_ok[MD];
    // head positions of the in-links
    int p_vid[MC], p_next[MC], p_dts[MC];
    short p_in[MC], p_j[MC], c_idx[MC], expd[MC];
    unsigned char p_flag[MC];
    double p_x[MC], p_xold[MC], p_v[MC], p_mr[MC], p_atl[MC], p_ex[MC], p_dist[MC];
    int nc, npos, nol, any_wait, nexp, first, cur_slot, n_picks;
    int omask[MD][2];          // per requested out-link: candidates currently eligible to move onto it
    unsigned char ogood[MD];   // per requested out-link: acceptance test result
    unsigned char c_first[MC], c_q[MC];
    double fc_rem;
    short shuf_j[MC];
    double u_merge[MC];
    // requested out-links
    int ol[MD], o_hd[MD], o_tail[MD], o_lanes[MD], o_cap[MD], o_lastlane[MD], o_win_off[MD + 1], o_win_n[MD];
    long long o_roff[MD];
    double o_ci[MD], o_dpl[MD], o_len[MD], o_vmax[MD], o_cumarr[MD];
    double w_x[MC], w_xold[MC];  // window of the last `lanes` platoons of each out-link (oldest first)
};

template <int MC, int MD>
UX_DEV void xfer_end_trip(const DevWorld &W, XferSmem<MC, MD> &S, int p, int t) {  // Vehicle::end_trip traffi.cpp:886-927
    int ii = S.p_in[p], il = S.il[ii], vid = S.p_vid[p];
    S.trips[ii] += 1;
    S.cumdep[ii] += W.dn;
    double atl = S.p_atl[p];
    {   // record_tt_event
        int s = (int)(atl / W.dt);
        if (s < 0) s = 0;
        if (s >= W.T) s = W.T - 1;
        int ord = ++S.evc[ii];
        W.ev_ord[(size_t)s * W.L + il] = ord;
        W.ev_tt[(size_t)s * W.L + il] = ((double)t + 1.0) * W.dt - atl;
    }
    W.v_atl[vid] = (double)t * W.dt;
    W.v_dist[vid] = S.p_dist[p] + (S.p_x[p] - S.p_ex[p]);
    W.v_link[vid] = -1;
    if (S.p_flag[p] & VF_ABORT) { W.v_state[vid] = UX_VS_ABORT; W.v_arr_ts[vid] = -1; W.v_tt[vid] = -1.0; }
    else { W.v_state[vid] = UX_VS_END; W.v_arr_ts[vid] = t; W.v_tt[vid] = (double)t * W.dt - (double)S.p_dts[p] * W.dt; }
    if (W.log_mode) { W.v_end_ts[vid] = t; W.v_end_kind[vid] = 2; }
    W.v_flags[vid] = 0;
    S.popped[ii] += 1;
}

#ifndef XFER_WARPS
#define XFER_WARPS 4
#endif
template <int MC, int MD>
UX_DEV void xfer_node(const DevWorld &W, XferSmem<MC, MD> &S, int n, int t, bool has_deps) {
    int cur = t & 1, L = W.L;
    size_t C = (size_t)W.C;
    int i0 = W.n_in_off[n], nin = W.n_in_off[n + 1] - i0;
    if (nin == 0) return;
    int nsig = W.n_sig_off[n + 1] - W.n_sig_off[n], phase = W.sig_phase[n];
    // ---- A1: in-link state
    WARP_FOR(lane) {
        for (int ii = lane; ii < nin; ii += 32) {
            // loads first, shared stores afterwards: the pointers are generic, so a shared store between two loads
            // would serialise them (the compiler must assume the load may target shared memory)
            int il = W.n_in[i0 + ii];
            int hd_ = W.head[cur * L + il], hc_ = W.hcount[il], cap_ = W.l_rcap[il], evc_ = W.ev_cnt[il], trips_ = W.trips[il];
            long long roff_ = W.l_roff[il];
            double co_ = W.cap_out_rem[il], mp_ = W.l_mp[il], vm_ = W.l_vmax[il], cd_ = W.cum_dep[(size_t)t * L + il];
            bool ok = true;
            if (nsig > 1) { int g0 = W.l_sg_off[il], gl = W.l_sg_off[il + 1] - g0; ok = list_has(W.l_sg + g0, gl, phase); }
            S.il[ii] = il; S.hd[ii] = hd_; S.hc[ii] = hc_; S.cap[ii] = cap_; S.roff[ii] = roff_;
            S.popped[ii] = 0; S.evc[ii] = evc_; S.trips[ii] = trips_;
            S.co_rem[ii] = co_; S.mp[ii] = mp_; S.vmax[ii] = vm_; S.cumdep[ii] = cd_;
            S.sig_ok[ii] = ok ? 1 : 0;
        }
    }
    WARP_SYNC();
    WARP_FOR(lane) if (lane == 0) {
        int acc = 0;
        for (int ii = 0; ii < nin; ii++) { S.pos_off[ii] = acc; acc += S.hc[ii]; }
        S.pos_off[nin] = acc; S.npos = acc < MC ? acc : MC;
    }
    WARP_SYNC();
    int npos = S.npos;
    if (npos == 0) return;
    // ---- A2: the platoons at the in-link heads
    WARP_FOR(lane) {
        for (int p = lane; p < npos; p += 32) {
            int ii = 0;
            while (S.pos_off[ii + 1] <= p) ii++;
            int j = p - S.pos_off[ii];
            long long o = S.roff[ii];
            int slot = ring_slot(S.hd[ii], j, S.cap[ii]);
            int vid = W.VID[o + slot];
            unsigned char f = W.v_flags[vid];
            int nx = W.v_next[vid];
            double x_ = 0, xo_ = 0, v_ = 0, mr_ = 0, atl_ = 0, ex_ = 0, di_ = 0; int dts_ = 0;
            if (f & (VF_CAND | VF_WAIT_END)) {
                x_ = W.X[(size_t)cur * C + o + slot]; xo_ = W.X[(size_t)(cur ^ 1) * C + o + slot]; v_ = W.V[o + slot];
                mr_ = W.v_mr[vid]; atl_ = W.v_atl[vid]; ex_ = W.v_entry_x[vid]; di_ = W.v_dist[vid]; dts_ = W.v_depart_ts[vid];
            }
            S.p_vid[p] = vid; S.p_in[p] = (short)ii; S.p_j[p] = (short)j; S.p_flag[p] = f; S.p_next[p] = nx;
            S.p_x[p] = x_; S.p_xold[p] = xo_; S.p_v[p] = v_; S.p_mr[p] = mr_; S.p_atl[p] = atl_; S.p_ex[p] = ex_; S.p_dist[p] = di_; S.p_dts[p] = dts_;
        }
    }
    WARP_SYNC();
    // ---- A3: candidates sorted by vehicle id (traffi.cpp:250-253; rank sort, one lane per head position) and the
    //          distinct requested out-links in order of first appearance (traffi.cpp:260-273)
    WARP_FOR(lane) if (lane == 0) { S.nc = 0; S.any_wait = 0; }
    WARP_SYNC();
    WARP_FOR(lane) {
        for (int p = lane; p < npos; p += 32) {
            unsigned char f = S.p_flag[p];
            if (f & VF_WAIT_END) atomicOr(&S.any_wait, 1);
            if (f & VF_CAND) {
                int vid = S.p_vid[p], rank = 0;
                for (int q = 0; q < npos; q++) if ((S.p_flag[q] & VF_CAND) && S.p_vid[q] < vid) rank++;
                S.c_idx[rank] = (short)p;
                atomicAdd(&S.nc, 1);
            }
        }
    }
    WARP_SYNC();
    WARP_FOR(lane) {
        int ncl = S.nc;
        for (int k = lane; k < ncl; k += 32) {
            int nl = S.p_next[S.c_idx[k]];
            bool first = nl >= 0;
            for (int q = 0; q < k && first; q++) if (S.p_next[S.c_idx[q]] == nl) first = false;
            S.c_first[k] = first ? 1 : 0;
        }
    }
    WARP_SYNC();
    WARP_FOR(lane) if (lane == 0) {
        int nol = 0;
        for (int k = 0; k < S.nc; k++) if (S.c_first[k] && nol < MD) S.ol[nol++] = S.p_next[S.c_idx[k]];
        S.nol = nol;
    }
    WARP_SYNC();
    int nc = S.nc, nol = S.nol;
    if (nc == 0 && !S.any_wait) return;
    WARP_FOR(lane) {
        for (int k = lane; k < nc; k += 32) {
            int nl = S.p_next[S.c_idx[k]], q = 255;
            for (int r = 0; r < nol; r++) if (S.ol[r] == nl) { q = r; break; }
            S.c_q[k] = (unsigned char)q;
        }
    }
    // ---- dependency wait (only nodes with SEES_POPS out-links): everything above -- in-link heads, candidate sort -- is
    //      independent of the lower-level nodes, so it overlaps their work; only the out-link tail state needs their pops.
    if (has_deps) {
        WARP_FOR(lane) {
            for (int q = W.dep_off[n] + lane; q < W.dep_off[n + 1]; q += 32) {
                volatile int *done = W.node_done + W.dep_list[q];
#ifdef UX_EMU
                if (*done < t + 1) dev_error(W, DERR_RING_FULL, -1);
#else
                while (*done < t + 1) __nanosleep(32);
#endif
            }
        }
#ifndef UX_EMU
        __threadfence();
#endif
        WARP_SYNC();
    }
    // ---- A4: out-link tail state
    WARP_FOR(lane) {
        for (int q = lane; q < nol; q += 32) {
            int ol = S.ol[q];
            int hd = W.head[cur * L + ol], fl_ = W.l_flags[ol], pk_ = ld_cg(W.pop_k2 + ol);  // written by a lower level in this launch
            int tl = W.tail[ol], lanes = W.l_lanes[ol], cap = W.l_rcap[ol];
            long long o = W.l_roff[ol];
            double ci_ = W.cap_in_rem[ol], dpl_ = W.l_dpl_dn[ol], len_ = W.l_length[ol], vm_ = W.l_vmax[ol], ca_ = W.cum_arr[(size_t)t * L + ol];
            if (fl_ & LF_SEES_POPS) hd += pk_;
            int sz = tl - hd;
            int ll_ = sz > 0 ? W.LANE[o + ring_slot(tl - 1, 0, cap)] : -1;
            S.o_hd[q] = hd; S.o_tail[q] = tl; S.o_lanes[q] = lanes; S.o_cap[q] = cap; S.o_roff[q] = o;
            S.o_ci[q] = ci_; S.o_dpl[q] = dpl_; S.o_len[q] = len_; S.o_vmax[q] = vm_; S.o_cumarr[q] = ca_;
            S.o_lastlane[q] = ll_;
            S.o_win_n[q] = sz < lanes ? sz : lanes;
        }
    }
    WARP_SYNC();
    WARP_FOR(lane) if (lane == 0) {
        int acc = 0;
        for (int q = 0; q < nol; q++) { S.o_win_off[q] = acc; acc += S.o_lanes[q]; }
        S.o_win_off[nol] = acc;
    }
    WARP_SYNC();
    WARP_FOR(lane) {
        int tot = S.o_win_off[nol];
        for (int e = lane; e < tot && e < MC; e += 32) {
            int q = 0;
            while (S.o_win_off[q + 1] <= e) q++;
            int k = e - S.o_win_off[q];
            if (k < S.o_win_n[q]) {
                int sz = S.o_tail[q] - S.o_hd[q];
                int slot = ring_slot(S.o_hd[q], sz - S.o_win_n[q] + k, S.o_cap[q]);
                S.w_x[e] = W.X[(size_t)cur * C + S.o_roff[q] + slot];
                S.w_xold[e] = W.X[(size_t)(cur ^ 1) * C + S.o_roff[q] + slot];
            }
        }
    }
    WARP_SYNC();
    // ---- B: the slot loop.  Slots are inherently sequential (each move changes what the next slot sees), but inside
    //          a slot the eligibility scan over the candidates runs one lane per candidate, and all random numbers of
    //          the node (shuffle indices, merge-lottery uniforms) are drawn up front in parallel: Philox is stateless.
    WARP_FOR(lane) if (lane == 0) {
        int nexp = 0;
        for (int q = 0; q < nol; q++) for (int r = 0; r < S.o_lanes[q] && nexp < MC; r++) S.expd[nexp++] = (short)q;
        S.nexp = nexp;
    }
    WARP_SYNC();
    int nexp = S.nexp;
    int nshuf = (!W.hard_det && nexp > 1) ? nexp - 1 : 0;
    if (!W.hard_det) {
        WARP_FOR(lane) {
            for (int i = 1 + lane; i < nexp; i += 32)  // Fisher-Yates step i uses draw number (nexp-1-i) (traffi.cpp:276-282)
                S.shuf_j[i] = (short)ux_rng_below(W.seed, (uint32_t)t, (uint32_t)n, UX_RNG_TRANSFER, (uint32_t)(nexp - 1 - i), i + 1);
            for (int d = lane; d < nexp; d += 32)
                S.u_merge[d] = ux_rng_uniform(W.seed, (uint32_t)t, (uint32_t)n, UX_RNG_TRANSFER, (uint32_t)(nshuf + d));
        }
        WARP_SYNC();
        WARP_FOR(lane) if (lane == 0) {
            for (int i = nexp - 1; i > 0; i--) { int j = S.shuf_j[i]; short tmp = S.expd[i]; S.expd[i] = S.expd[j]; S.expd[j] = tmp; }
        }
        WARP_SYNC();
    }
    // Rounds instead of slots: a slot whose out-link cannot accept or has no eligible candidate changes nothing (and
    // draws nothing), so each round evaluates every out-link / candidate in parallel against the current state, jumps
    // to the first slot (in shuffled order) that moves a platoon, applies that one move on lane 0, and repeats.
    double fc = W.n_fc[n];
    WARP_FOR(lane) if (lane == 0) { S.cur_slot = 0; S.n_picks = 0; S.fc_rem = W.fc_rem[n]; }
    WARP_SYNC();
    while (true) {
        WARP_FOR(lane) {
            if (lane == 0) S.first = 0x7fffffff;
            for (int q = lane; q < nol; q += 32) {
                S.omask[q][0] = 0; S.omask[q][1] = 0;
                int lanes_ol = S.o_lanes[q], sz = S.o_tail[q] - S.o_hd[q];
                bool ok = false;  // acceptance (traffi.cpp:285-300)
                if (sz < lanes_ol) ok = true;
                else if (S.w_x[S.o_win_off[q]] > S.o_dpl[q]) ok = true;
                if (S.o_ci[q] < W.dn) ok = false;
                if (S.fc_rem < W.dn) ok = false;
                S.ogood[q] = ok ? 1 : 0;
            }
        }
        WARP_SYNC();
        WARP_FOR(lane) {
            for (int k = lane; k < nc; k += 32) {
                int p = S.c_idx[k], q = S.c_q[k];
                if (p < 0 || q == 255) continue;
                int ii = S.p_in[p];
                if (S.p_j[p] != S.popped[ii]) continue;  // front of its in-link
                if (S.co_rem[ii] < W.dn) continue;
                if (!S.sig_ok[ii]) continue;
                atomicOr(&S.omask[q][k >> 5], (int)(1u << (k & 31)));
            }
        }
        WARP_SYNC();
        WARP_FOR(lane) {
            for (int s = S.cur_slot + lane; s < nexp; s += 32) {
                int q = S.expd[s];
                if (S.ogood[q] && (S.omask[q][0] | S.omask[q][1])) { atomicMin(&S.first, s); break; }
            }
        }
        WARP_SYNC();
        int s = S.first;
        if (s == 0x7fffffff) break;
        WARP_FOR(lane) if (lane == 0) {
            int q = S.expd[s], ol = S.ol[q];
            unsigned m0 = (unsigned)S.omask[q][0], m1 = (unsigned)S.omask[q][1];
            int lanes_ol = S.o_lanes[q];
            int sz = S.o_tail[q] - S.o_hd[q], w0 = S.o_win_off[q];
            double fc_rem = S.fc_rem;
            int n_picks = S.n_picks;
            short mv[MC]; double mp[MC]; int nm = 0;
            for (int wd = 0; wd < 2; wd++) {  // eligible candidates in ascending vehicle id
                unsigned m = wd ? m1 : m0;
                while (m) {
                    int bit = 0; while (!((m >> bit) & 1u)) bit++;
                    m &= m - 1;
                    int k = wd * 32 + bit;
                    mv[nm] = (short)k; mp[nm] = S.mp[S.p_in[S.c_idx[k]]]; nm++;
                }
            }
            int pick;
            if (W.hard_det) { pick = 0; for (int i = 1; i < nm; i++) if (mp[i] > mp[pick]) pick = i; }
            else pick = weighted_pick(mp, nm, S.u_merge[n_picks++]);
            int k = mv[pick], p = S.c_idx[k], vid = S.p_vid[p], ii = S.p_in[p], il = S.il[ii];
            S.co_rem[ii] -= W.dn; S.o_ci[q] -= W.dn;
            if (fc >= 0.0) fc_rem -= W.dn;
            S.cumdep[ii] += W.dn; S.o_cumarr[q] += W.dn;
            double atl = S.p_atl[p];
            {   // record_tt_event (traffi.cpp:357-362)
                int es = (int)(atl / W.dt);
                if (es < 0) es = 0;
                if (es >= W.T) es = W.T - 1;
                int ord = ++S.evc[ii];
                W.ev_ord[(size_t)es * L + il] = ord;
                W.ev_tt[(size_t)es * L + il] = (double)t * W.dt - atl;
            }
            W.v_atl[vid] = (double)t * W.dt;
            W.v_dist[vid] = S.p_dist[p] + (S.p_x[p] - S.p_ex[p]);
            S.popped[ii] += 1;
            if (W.no_cyclic) { int sn = W.l_start[ol]; W.v_visited[(size_t)vid * W.vis_words + (sn >> 5)] |= 1u << (sn & 31); }
            { int k = W.v_trav[vid]; W.v_trav[vid] = k + 1; record_traversal(W, vid, k, ol, t); }
            W.v_link[vid] = ol;
            W.v_flags[vid] = 0;
            // carry the residual move onto the new link (traffi.cpp:400-419)
            double x_next = S.p_mr[p] * S.o_vmax[q] / S.vmax[ii];
            if (sz >= lanes_ol) {
                double x_cong = S.w_xold[w0] - S.o_dpl[q];
                if (x_cong < 0.0) x_cong = 0.0;
                if (x_next > x_cong) x_next = x_cong;
            }
            if (x_next >= S.o_len[q]) x_next = S.o_len[q];
            if (x_next < 0.0) x_next = 0.0;
            if (sz >= S.o_cap[q]) dev_error(W, DERR_RING_FULL, ol);
            else {  // push (lane rule traffi.cpp:385-390)
                int lane_new = sz > 0 ? (S.o_lastlane[q] + 1) % lanes_ol : 0;
                int slot = ring_slot(S.o_tail[q], 0, S.o_cap[q]);
                long long o = S.o_roff[q];
                W.X[(size_t)cur * C + o + slot] = x_next;
                W.X[(size_t)(cur ^ 1) * C + o + slot] = S.p_xold[p];
                W.V[o + slot] = S.p_v[p] + x_next / W.dt;
                W.VID[o + slot] = vid;
                W.LANE[o + slot] = lane_new;
                S.o_tail[q] += 1; S.o_lastlane[q] = lane_new;
                if (S.o_win_n[q] < lanes_ol) { int e = w0 + S.o_win_n[q]; S.w_x[e] = x_next; S.w_xold[e] = S.p_xold[p]; S.o_win_n[q] += 1; }
                else {
                    for (int e = w0; e < w0 + lanes_ol - 1; e++) { S.w_x[e] = S.w_x[e + 1]; S.w_xold[e] = S.w_xold[e + 1]; }
                    S.w_x[w0 + lanes_ol - 1] = x_next; S.w_xold[w0 + lanes_ol - 1] = S.p_xold[p];
                }
            }
            W.v_entry_x[vid] = x_next;
            W.v_mr[vid] = 0.0;
            S.c_idx[k] = -1;  // remove from incoming
            // the new front of the in-link may be waiting for its trip end (traffi.cpp:421-424)
            if (S.popped[ii] < S.hc[ii]) {
                int fp = S.pos_off[ii] + S.popped[ii];
                if (fp < npos && (S.p_flag[fp] & VF_WAIT_END)) xfer_end_trip(W, S, fp, t);
            }
            S.fc_rem = fc_rem; S.n_picks = n_picks; S.cur_slot = s + 1;
        }
        WARP_SYNC();
    }
    WARP_FOR(lane) if (lane == 0) {
        // flush trip-end-waiting fronts (traffi.cpp:432-441)
        for (int ii = 0; ii < nin; ii++) {
            int lanes = S.hc[ii];  // hcount = min(lanes, queue length): positions beyond it cannot be waiting
            for (int r = 0; r < lanes; r++) {
                if (S.popped[ii] >= S.hc[ii]) break;
                int fp = S.pos_off[ii] + S.popped[ii];
                if (fp >= npos || !(S.p_flag[fp] & VF_WAIT_END)) break;
                xfer_end_trip(W, S, fp, t);
            }
        }
        W.fc_rem[n] = S.fc_rem;
    }
    WARP_SYNC();
    // ---- C: write the per-link accumulators back
    WARP_FOR(lane) {
        for (int ii = lane; ii < nin; ii += 32) {
            int il = S.il[ii];
            W.cap_out_rem[il] = S.co_rem[ii]; W.cum_dep[(size_t)t * L + il] = S.cumdep[ii];
            W.pop_k2[il] = S.popped[ii]; W.ev_cnt[il] = S.evc[ii]; W.trips[il] = S.trips[ii];
        }
        for (int q = lane; q < nol; q += 32) {
            int ol = S.ol[q];
            W.cap_in_rem[ol] = S.o_ci[q]; W.cum_arr[(size_t)t * L + ol] = S.o_cumarr[q]; W.tail[ol] = S.o_tail[q];
        }
    }
}

// k_node: one warp runs a node's whole timestep -- Link::update of the sides it owns, Node::generate, signal, node
// capacity (pre_node) and Node::transfer (xfer_node).
// All dependency levels run in ONE launch: warps are ordered by (level, node id), and a warp of level k > 0 waits
// until the specific lower-level nodes it depends on have finished this step (they are dispatched earlier -- the
// same forward-progress assumption as decoupled look-back scans).  Levels exist only for links short enough that the serial id-ordered
// visibility of Node.transfer matters (DESIGN.md "transfer order"); most networks have a single level and never wait.
template <int MC, int MD>
union NodeSmem {  // the two halves of a node's timestep use shared memory one after the other
    PreSmem<MC, MD> pre;
    XferSmem<MC, MD> xf;
};

// MC / MD bound the sum of lanes and the degree of a node; the launcher picks the smallest instantiation the network fits
// in, because the shared-memory footprint per warp decides how many node warps are resident.
template <int MC, int MD>
__global__ void __launch_bounds__(32 * XFER_WARPS, 32 / XFER_WARPS) k_node(const DevWorld *worlds, int step_off) {
    // replica = blockIdx.x (fastest in launch order): the blocks resident together are the same nodes of different replicas
    // -- same code path, same static tables -- and every replica's dependency chains start in the first wave
    LOAD_WORLD_N(W, worlds, (XFER_WARPS > 1 ? 64 : 32), blockIdx.x)
    WARP_BEGIN
    __shared__ NodeSmem<MC, MD> sm_all[XFER_WARPS];
    __shared__ double tt_all[XFER_WARPS][MD];
    __shared__ int rel_all[XFER_WARPS];
    int warp = (int)((blockIdx.y * blockDim.x + threadIdx.x) >> 5), wib = (int)((threadIdx.x >> 5) % XFER_WARPS);
    if (warp >= W.N) return;
    int n = W.lvl_nodes[warp], t = *W.t_base + step_off;
    int lvl_raw = W.n_levels > 1 ? W.lvl_level[warp] : 0, lvl = lvl_raw & 0xffff;  // bit 16: some node waits for this one
    pre_node(W, sm_all[wib].pre, tt_all[wib], &rel_all[wib], n, t, (int)(threadIdx.x & 31));
#ifndef UX_EMU
    __threadfence_block();
#endif
    XferSmem<MC, MD> &S = sm_all[wib].xf;
    xfer_node(W, S, n, t, lvl > 0);
    if (lvl_raw >> 16) {  // publish completion only where somebody is waiting: the fence is not free
        WARP_SYNC();
        WARP_FOR(lane) if (lane == 0) { __threadfence(); *((volatile int *)(W.node_done + n)) = t + 1; }
    }
}

// ---------------------------------------------------------------------------------------------------
// K3: Vehicle::car_follow_newell (traffi.cpp:798-823) + Vehicle::update RUN branch (844-878) for every
//     occupied ring slot.  The warp owning a link's first segment additionally handles the link head: the
//     platoons that can be at the link end are exactly the first `lanes` queue positions (anything further back
//     has a leader on the same link), so their trip end / abort / next-link choice is evaluated by one lane per
//     position in parallel and only the "is it the front" prefix (traffi.cpp:861, 869) is resolved serially.
// ---------------------------------------------------------------------------------------------------
// One RUN log record (Vehicle::log_data traffi.cpp:1060-1110) for the platoon at queue position `pos` of link l, taken
// BEFORE this step's move.  The spacing follows the reference's id-ordered emulation (1081-1095): a lower-id leader is
// seen after its move (or not at all if it just ended its trip), a higher-id leader before its move.
UX_DEV void log_run_record(const DevWorld &W, int l, int t, int cur, int hd, int n, int pos, int cap, long long o, int lanes,
                           double len, double udt, double dpl, int n_ended, int relabel_pushes) {
    size_t C = (size_t)W.C;
    const double *Xc = W.X + (size_t)cur * C + o;
    int p = ring_slot(hd, pos, cap);
    int vid = W.VID[o + p];
    double x = Xc[p], v = W.V[o + p];
    int lane = pos < relabel_pushes ? pos % lanes : W.LANE[o + p];
    double s = -1.0;
    if (pos >= lanes) {
        int pl = pos - lanes, lp = ring_slot(hd, pl, cap);
        int lvid = W.VID[o + lp];
        double xl = Xc[lp];
        if (lvid < vid) {
            if (pl >= n_ended) {  // leader still on the link after its own update: its post-move position
                double xn = xl + udt;
                if (pl >= lanes) { double gap = Xc[ring_slot(hd, pl - lanes, cap)] - dpl; if (gap < xl) gap = xl; if (xn > gap) xn = gap; }
                if (xn < xl) xn = xl;
                if (xn > len) xn = len;
                s = xn - x;
            }
        } else s = xl - x;
    }
    unsigned long long idx = atomicAdd(W.lg_count, 1ull);
    if ((long long)idx >= W.lg_cap) { dev_error(W, DERR_LOG_FULL, l); return; }
    W.lg_vid[idx] = vid; W.lg_t[idx] = t; W.lg_link[idx] = l; W.lg_lane[idx] = lane; W.lg_x[idx] = x; W.lg_v[idx] = v; W.lg_s[idx] = s;
}

#define UX_MAXLANE 64
#define UPD_WARPS 4
struct HeadSmem {
    int vid[UX_MAXLANE], next[UX_MAXLANE], dts[UX_MAXLANE];
    unsigned char kind[UX_MAXLANE];  // 0 not at end, 1 trip end, 2 abort, 3 transfer candidate
    double xn[UX_MAXLANE], atl[UX_MAXLANE], ex[UX_MAXLANE], dist[UX_MAXLANE];
    int k_ended;
};

// 4 warps per block and <= 48 registers (10 blocks per SM): the kernel is a swarm of tiny, latency-bound warps -- measured
// best among {8x4, 8x5, 4x10, 4x12} blocks-per-SM variants at 32 replicas.
__global__ void __launch_bounds__(32 * UPD_WARPS, 10) k_update(const DevWorld *worlds, int step_off) {
    LOAD_WORLD(W, worlds)
    WARP_BEGIN
    __shared__ HeadSmem hs_all[UPD_WARPS];
    int gw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (gw >= W.nseg) return;
    // work item = (link, part): the warp sweeps the 32-position windows part, part+nparts, ... of the link's queue, so the
    // work is proportional to the platoons present, and long queues are shared by several warps
    int l = W.seg_link[gw], part = W.seg_start[gw] & 0xffff, nparts = W.seg_start[gw] >> 16, L = W.L;
    int t = *W.t_base + step_off, cur = t & 1;
    size_t C = (size_t)W.C;
    int cap = W.l_rcap[l];
    long long o = W.l_roff[l];
    int hd0 = W.head[cur * L + l], pk2 = W.pop_k2[l], tl = W.tail[l];
    int lanes = W.l_lanes[l];
    double len = W.l_length[l], udt = W.l_udt[l], dpl = W.l_dpl_dn[l];
    int hd = hd0 + pk2, n = tl - hd;
    int s0 = part;
    const double *Xc = W.X + (size_t)cur * C + o;
    if (s0 == 0) {
        // ---- link head
        HeadSmem &H = hs_all[(threadIdx.x >> 5) % UPD_WARPS];
        int hc = n < lanes ? n : lanes;
        if (hc > UX_MAXLANE) hc = UX_MAXLANE;
        int end_node = W.l_end[l];
        bool dead_end = W.n_out_off[end_node + 1] == W.n_out_off[end_node];
        WARP_FOR(lane) {
            for (int j = lane; j < hc; j += 32) {
                int slot = ring_slot(hd, j, cap);
                int vid = W.VID[o + slot];
                double x = Xc[slot];
                int dest_ = W.v_dest[vid];
                double atl_ = W.v_atl[vid], ex_ = W.v_entry_x[vid], di_ = W.v_dist[vid]; int dts_ = W.v_depart_ts[vid];
                double xn = x + udt;  // positions < lanes have no leader
                if (xn < x) xn = x;
                bool over = xn > len;
                double mr_ = xn - len;
                if (over) xn = len;
                unsigned char kind = 0;
                int nx = -1;
                if (fabs(xn - len) < 1e-9) {
                    if (end_node == dest_) kind = 1;
                    else if (dead_end) kind = 2;  // traffi.cpp:864-871 (trip_abort is always 1 in this engine)
                    else { kind = 3; nx = route_choice(W, vid, end_node, t, (unsigned)vid, UX_RNG_VEHROUTE, 0u); }
                }
                if (over) W.v_mr[vid] = mr_;
                H.vid[j] = vid; H.kind[j] = kind; H.xn[j] = xn; H.next[j] = nx;
                H.atl[j] = atl_; H.ex[j] = ex_; H.dist[j] = di_; H.dts[j] = dts_;
            }
        }
        WARP_SYNC();
        WARP_FOR(lane) if (lane == 0) {
            // lane relabel when the end node (smaller id) emptied the link before this step's pushes
            if ((W.l_flags[l] & LF_END_LT_START) && !(W.l_flags[l] & LF_SEES_POPS) && pk2 > 0) {
                int n0 = W.tail_k1[l] - hd0, pushes = tl - W.tail_k1[l];
                if (pk2 == n0 && pushes > 0) for (int j = 0; j < pushes; j++) W.LANE[o + ring_slot(hd, j, cap)] = j % lanes;
            }
            W.stat_count[l] += n;
            int k = 0, evc = 0, trips = 0;
            double cumdep = 0.0;
            bool touched = false;
            for (int j = 0; j < hc; j++) {
                int vid = H.vid[j];
                unsigned char kind = H.kind[j], f = 0;
                if (kind == 3) { W.v_next[vid] = H.next[j]; f = VF_CAND; }
                else if (kind == 1 || kind == 2) {
                    f = kind == 1 ? VF_WAIT_END : (VF_WAIT_END | VF_ABORT);
                    if (kind == 2) W.v_next[vid] = -1;
                    if (k == j) {  // it is the front (everything ahead ended in this pass): end_trip, traffi.cpp:886-927
                        if (!touched) { evc = W.ev_cnt[l]; trips = W.trips[l]; cumdep = W.cum_dep[(size_t)t * L + l]; touched = true; }
                        trips += 1; cumdep += W.dn;
                        double atl = H.atl[j];
                        int es = (int)(atl / W.dt);
                        if (es < 0) es = 0;
                        if (es >= W.T) es = W.T - 1;
                        evc += 1;
                        W.ev_ord[(size_t)es * L + l] = evc;
                        W.ev_tt[(size_t)es * L + l] = ((double)t + 1.0) * W.dt - atl;
                        W.v_atl[vid] = (double)t * W.dt;
                        W.v_dist[vid] = H.dist[j] + (H.xn[j] - H.ex[j]);
                        W.v_link[vid] = -1;
                        if (kind == 2) { W.v_state[vid] = UX_VS_ABORT; W.v_arr_ts[vid] = -1; W.v_tt[vid] = -1.0; }
                        else { W.v_state[vid] = UX_VS_END; W.v_arr_ts[vid] = t; W.v_tt[vid] = (double)t * W.dt - (double)H.dts[j] * W.dt; }
                        if (W.log_mode) { W.v_end_ts[vid] = t; W.v_end_kind[vid] = 1; }
                        f = 0;
                        k++;
                    }
                }
                W.v_flags[vid] = f;
            }
            if (touched) { W.ev_cnt[l] = evc; W.trips[l] = trips; W.cum_dep[(size_t)t * L + l] = cumdep; }
            W.head[(cur ^ 1) * L + l] = hd + k;
            int rem = n - k;
            W.hcount[l] = rem < lanes ? rem : lanes;
            H.k_ended = k;
        }

        if (W.log_mode) {  // positions that may have a just-ended leader are logged here, once the ended count is known
            WARP_SYNC();
            int nlog = n < 2 * lanes ? n : 2 * lanes, k_ended = H.k_ended;
            WARP_FOR(lane) { for (int pos = lane; pos < nlog; pos += 32) log_run_record(W, l, t, cur, hd, n, pos, cap, o, lanes, len, udt, dpl, k_ended, 0); }
            WARP_SYNC();
        }
    }
    WARP_FOR(lane) {
        for (int pos = part * 32 + lane; pos < n; pos += nparts * 32) {
            int p = ring_slot(hd, pos, cap);
            if (W.log_mode && pos >= 2 * lanes) log_run_record(W, l, t, cur, hd, n, pos, cap, o, lanes, len, udt, dpl, 0, 0);
            double x = Xc[p];
            double xn = x + udt;
            if (pos >= lanes) {
                double gap = Xc[ring_slot(hd, pos - lanes, cap)] - dpl;
                if (gap < x) gap = x;
                if (xn > gap) xn = gap;
            }
            if (xn < x) xn = x;
            if (xn > len) xn = len;
            W.X[(size_t)(cur ^ 1) * C + o + p] = xn;
            W.V[o + p] = (xn - x) / W.dt;
        }
    }
}

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
        double tt = W.tt_inst[(size_t)t * W.L + l];
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
        int ord = W.ev_ord[(size_t)i * W.L + l];
        double v = W.ev_tt[(size_t)i * W.L + l];
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
        int ord = W.ev_ord[(size_t)i * L + l];
        double v = W.ev_tt[(size_t)i * L + l];
        if (ord > best) { best = ord; cur = v; }
        s1 += cur; s2 += cur * cur;
    }
    double *o = part + ((size_t)blockIdx.y * L + l) * 4;
    o[0] = T > 0 ? W.cum_arr[(size_t)(T - 1) * L + l] : 0.0; o[1] = s1; o[2] = s2;
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
        if (tt_i > fftt && W.cum_arr[(size_t)fut * L + l] - W.cum_dep[(size_t)fut * L + l] > 0) ts_end = i; else break;
    }
    double veh_count = W.cum_arr[(size_t)ts_end * L + l] - W.cum_arr[(size_t)ts * L + l];
    int ts_out = (int)(ts + (int)(ttl[ts] / dt)) + 1;
    if (ts_out >= n) ts_out = n - 1;
    int ts_out_leader = ts_out;
    double dep_out = W.cum_dep[(size_t)ts_out * L + l];
    for (int i = ts_out; i > 0; i--) if (W.cum_dep[(size_t)i * L + l] < dep_out) { ts_out_leader = i; break; }  // leader headway (backward scan)
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
    if (st == UX_VS_END || st == UX_VS_ABORT) {
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

// =====================================================================================================
// host side
// =====================================================================================================
struct HNode {
    double x, y, sig_offset, fc; std::vector<double> sig; int lanes;
    int sig_phase0; double sig_t0, fc_rem0;
    std::vector<int> in, out;
};
struct HLink {
    int start, end, lanes; double length, vmax, kappa, delta, dpl, tau, w, capacity, mp, cap_out, cap_in, cap_out_rem0, cap_in_rem0, penalty;
    std::vector<int> sg; std::vector<double> toll;
};
struct HVeh { int orig, dest, ts; };
struct Slab { char *base; size_t size, used; };

// OD route-set store of the DTA solvers (pairs sorted by (orig, dest)) with its device copy.  One store serves every day of
// a solve, so worlds share it by reference; the device copy lives as long as any world does.
struct DtaStore {
    std::vector<int> o, d, links; std::vector<long long> od_off{0}, route_off{0};
    int device = -1; void *d_block = nullptr;
    int *d_o = nullptr, *d_d = nullptr, *d_links = nullptr; long long *d_od_off = nullptr, *d_route_off = nullptr;
    ~DtaStore() { if (d_block) { cudaSetDevice(device); cudaFree(d_block); } }
};

struct ux_world {
    ux_world_params p;
    double dt; int total_ts, route_ts, timestep; int R;
    std::vector<HNode> nodes; std::vector<HLink> links; std::vector<HVeh> vehs;
    std::map<long long, std::vector<int>> vlinks[3];
    bool initialized = false, uploaded = false, net_dirty = false;
    std::string err;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<void *> allocs;  // cudaMalloc'ed scratch (freed with the world)
    std::vector<Slab> slabs;     // state memory (returned to the slab cache)
    std::vector<DevWorld> hw;  // host copies of the per-replica device worlds
    DevWorld *dworlds = nullptr;
    int n_levels = 1, n_active = 0;
    long long launches = 0, C = 0; int nseg = 0, n_edges = 0;
    std::vector<int> dest_row, row_dest;
    long long P_uploaded = 0;
    bool all_dests = false;
    int max_node_lanes = 0, max_node_deg = 0;
    std::vector<int> h_gen_off, h_gen_list, eff_ts;  // host copies of the per-origin FIFO lists; effective departure step per platoon
#ifndef UX_EMU
    std::map<long long, std::pair<cudaGraphExec_t, long long>> graphs;  // key -> (executable graph, kernels per launch)
#endif
    // device pointers of updatable static arrays
    double *d_vmax = nullptr, *d_dpl_dn = nullptr, *d_udt = nullptr, *d_cap_out = nullptr, *d_cap_in = nullptr, *d_mp = nullptr, *d_penalty = nullptr, *d_toll = nullptr, *d_n_sig = nullptr;
    int *d_l_flags = nullptr, *d_lvl_off = nullptr, *d_lvl_nodes = nullptr, *d_lvl_level = nullptr, *d_dep_off = nullptr, *d_dep_list = nullptr, *d_sg_off = nullptr, *d_sg = nullptr, *d_n_sig_off = nullptr;
    double *d_scratch = nullptr; size_t scratch_bytes = 0;
    double *d_ens = nullptr;
    // assembled per-platoon logs of one replica (World::build_all_vehicle_logs_flat_compact layout), built on demand
    struct LogCache {
        int replica = -1, timestep = -1;
        std::vector<long long> offsets, ltl_offsets;
        std::vector<double> t, x, v, s, ltl_t, enter_t;
        std::vector<int> state, lane, link, n_missing, ltl_id, enter_link, enter_veh;
    } logs;
    // DTA solver: OD route-set store (pairs sorted by (orig, dest)), its device copy, and the last route-swap result
    bool record_traversals = false;
    std::map<int, std::pair<std::vector<int>, std::vector<int>>> rep_routes;  // replica -> (offsets [P+1], link ids): per-replica enforced routes
    std::shared_ptr<DtaStore> dta_p = std::make_shared<DtaStore>();  // shared between the worlds of one solve (ux_dta_share_route_sets)
    struct DtaResult {
        bool valid = false; int replica = 0;
        long long P = 0, rs_total = 0, ra_total = 0;
        // host copies, filled from the device on first request (a replica-batched solve never asks)
        std::vector<long long> rs_off, ra_off; std::vector<int> rs_links, ra_links, choice; std::vector<double> ra_entry_t, cost_actual, t_gap;
        bool rs_fetched = false, ra_fetched = false, meta_fetched = false;
        const int *d_tr_link = nullptr, *d_tr_t = nullptr, *d_choice = nullptr; const long long *d_tr_off = nullptr; const double *d_cost = nullptr, *d_gap = nullptr;
        double n_swap = 0, potential_n_swap = 0, total_t_gap = 0;
    };
    std::vector<DtaResult> swaps;  // per replica
    char *swap_buf = nullptr;      // device scratch of the route swaps, kept between days (traversed routes stay there for lazy hand-back)
    size_t swap_buf_bytes = 0;
    struct NextRoutes { int *off = nullptr, *links = nullptr; size_t cap = 0; bool ready = false; };
    std::vector<NextRoutes> next_routes;  // per replica: routes_specified of the last swap as a device CSR (next day's enforced routes)
    // re-initialisation recipe of the per-replica dynamic state (ux_dta_next_day rewinds the world instead of rebuilding it)
    struct ReInit { void *dst; size_t bytes; int kind; std::shared_ptr<std::vector<char>> src; };  // 0: bytes of 0xff, 1: doubles of -1.0, 2: copy of src
    std::vector<ReInit> reinit;
    size_t dyn_slab0 = 0, dyn_used0 = 0, dyn_slab1 = 0, dyn_used1 = 0;
    bool rewindable = false;
    int swap_last = 0;             // replica of the most recent result (scalar queries without a replica index)
    double swap_ms = 0;
    // measurement
    long long h2d_bytes = 0, d2h_bytes = 0;
    double last_loop_ms = 0.0, upload_ms = 0.0, loop_wall_ms = 0.0, graph_build_ms = 0.0;
    int *d_err_all = nullptr;
    bool ktiming = false;
    double k_ms[UX_K_COUNT] = {0};
    long long k_launches[UX_K_COUNT] = {0};
#ifndef UX_EMU
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::vector<cudaEvent_t> kt_events;  // pairs (start, stop)
    std::vector<int> kt_ids;
#endif
};

static std::string g_err;

static int fail(ux_world *w, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    (w ? w->err : g_err) = buf;
    return code;
}

#define CUDA_OK(call)                                                                                        \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return fail(w, UX_ERR_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---- device memory.  A world bump-allocates its arrays from slabs; slabs come from (and return to) a per-device
// cache that outlives the world, so building the ~45 state arrays of each replica costs a handful of driver calls and a
// solver that builds a world per iteration (DTA days, ensemble batches) reuses the previous world's memory.  A slab is
// zeroed once when a world acquires it (on the world's stream), which is the initial value of almost all state.
static std::mutex g_slab_mu;
static std::map<int, std::vector<Slab>> g_slab_cache;  // device -> free slabs
static size_t g_slab_cached_bytes = 0;

static size_t slab_cache_limit() {  // bytes kept across worlds (env UXSIM_B200_SLAB_CACHE_MB; default 32 GiB)
    static size_t lim = [] { const char *e = getenv("UXSIM_B200_SLAB_CACHE_MB"); return e ? (size_t)atoll(e) << 20 : (size_t)32 << 30; }();
    return lim;
}
static void slab_cache_drop(int device) {
    std::lock_guard<std::mutex> g(g_slab_mu);
    for (Slab &s : g_slab_cache[device]) { cudaFree(s.base); g_slab_cached_bytes -= s.size; }
    g_slab_cache[device].clear();
}
static bool slab_acquire(ux_world *w, size_t need) {
    const size_t MB = (size_t)1 << 20;
    size_t next = w->slabs.empty() ? 4 * MB : std::min(w->slabs.back().size * 2, 512 * MB);
    size_t want = std::max((need + 2 * MB - 1) / (2 * MB) * (2 * MB), next);
    Slab got{nullptr, 0, 0};
    {
        std::lock_guard<std::mutex> g(g_slab_mu);
        auto &fl = g_slab_cache[w->device];
        int best = -1;
        for (int i = 0; i < (int)fl.size(); i++)
            if (fl[i].size >= need && fl[i].size <= 2 * want && (best < 0 || fl[i].size < fl[best].size)) best = i;
        if (best >= 0) { got = fl[best]; fl.erase(fl.begin() + best); g_slab_cached_bytes -= got.size; }
    }
    if (!got.base) {
        void *p = nullptr;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            slab_cache_drop(w->device);  // memory pressure: give the cache back and retry once
            if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return false; }
        }
        got.base = (char *)p; got.size = want;
    }
    got.used = 0;
    cudaMemsetAsync(got.base, 0, got.size, w->stream);
    w->slabs.push_back(got);
    return true;
}
static void slabs_release(ux_world *w) {
    std::lock_guard<std::mutex> g(g_slab_mu);
    for (Slab &s : w->slabs) {
        if (g_slab_cached_bytes + s.size <= slab_cache_limit()) { g_slab_cache[w->device].push_back(s); g_slab_cached_bytes += s.size; }
        else cudaFree(s.base);
    }
    w->slabs.clear();
}

template <typename T>
static T *dalloc(ux_world *w, size_t n, bool zero = true) {  // slab memory is already zero; `zero` documents who relies on it
    (void)zero;
    size_t bytes = ((n ? n : 1) * sizeof(T) + 255) & ~(size_t)255;
    if (w->slabs.empty() || w->slabs.back().size - w->slabs.back().used < bytes)
        if (!slab_acquire(w, bytes)) return nullptr;
    Slab &s = w->slabs.back();
    T *p = (T *)(s.base + s.used);
    s.used += bytes;
    return p;
}
template <typename T>
static T *dupload(ux_world *w, const std::vector<T> &v) {
    T *p = dalloc<T>(w, v.size(), false);
    // pageable source: the call returns once the data is staged, so `v` may die right after
    if (p && !v.empty()) { cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, w->stream); w->h2d_bytes += (long long)(v.size() * sizeof(T)); }
    return p;
}

extern "C" {

// create_world (bindings.cpp:150-181; World::World traffi.cpp:1165-1216)
UX_API ux_world *ux_create_world(const ux_world_params *params) {
    if (!params || params->delta_n <= 0 || params->tau <= 0) { fail(nullptr, UX_ERR_INVALID, "bad world params"); return nullptr; }
#ifndef UX_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { fail(nullptr, UX_ERR_CUDA, "no CUDA device: the uxsim_b200 engine has no CPU fallback"); return nullptr; }
#endif
    ux_world *w = new ux_world();
    w->p = *params;
    w->dt = params->tau * params->delta_n;
    w->total_ts = (int)(params->t_max / (params->tau * params->delta_n));
    w->route_ts = (int)(params->duo_update_time / (params->tau * params->delta_n));
    if (w->route_ts == 0) w->route_ts = 1;
    if (w->p.instantaneous_TT_timestep_interval <= 0) w->p.instantaneous_TT_timestep_interval = 5;
    w->R = params->num_replicas > 0 ? params->num_replicas : 1;
    w->timestep = 0;
    if (params->device >= 0) { w->device = params->device; cudaSetDevice(w->device); } else cudaGetDevice(&w->device);
    return w;
}

UX_API void ux_destroy_world(ux_world *w) {
    if (!w) return;
#ifndef UX_EMU
    cudaSetDevice(w->device);
    for (auto &g : w->graphs) cudaGraphExecDestroy(g.second.first);
#endif
    if (w->stream) cudaStreamSynchronize(w->stream);
    for (void *p : w->allocs) cudaFree(p);
    if (w->swap_buf) cudaFree(w->swap_buf);
    for (auto &nr_ : w->next_routes) { if (nr_.off) cudaFree(nr_.off); if (nr_.links) cudaFree(nr_.links); }
    slabs_release(w);
    if (w->stream) cudaStreamDestroy(w->stream);
    delete w;
}

// add_node (bindings.cpp:186-190; Node::Node traffi.cpp:47-100)
UX_API int ux_add_node(ux_world *w, double x, double y, const double *sig, int nsig, double sig_offset, double flow_capacity, int number_of_lanes) {
    if (w->initialized) return fail(w, UX_ERR_RUNTIME, "add_node after initialize_adj_matrix");
    HNode n;
    n.x = x; n.y = y; n.sig_offset = sig_offset;
    if (nsig <= 0) n.sig = {0.0}; else n.sig.assign(sig, sig + nsig);
    n.sig_phase0 = 0; n.sig_t0 = 0;
    double cycle = 0; for (double s : n.sig) cycle += s;
    if (n.sig.size() > 1 || (n.sig.size() == 1 && n.sig[0] != 0)) {
        double offset = cycle - sig_offset; int i = 0, guard = 0;
        while (true) {
            if (offset < n.sig[i]) { n.sig_phase0 = i; n.sig_t0 = offset; break; }
            offset -= n.sig[i]; i++; if (i >= (int)n.sig.size()) i = 0;
            if (++guard > 1000000) return fail(w, UX_ERR_INVALID, "signal offset walk does not terminate");
        }
    }
    if (flow_capacity >= 0.0) { n.fc = flow_capacity; n.fc_rem0 = flow_capacity * w->p.tau * w->p.delta_n; n.lanes = number_of_lanes > 0 ? number_of_lanes : (int)std::ceil(flow_capacity / 0.8); }
    else { n.fc = -1.0; n.fc_rem0 = 10e10; n.lanes = number_of_lanes; }
    w->nodes.push_back(n);
    return (int)w->nodes.size() - 1;
}

static void link_derive(ux_world *w, HLink &l) {  // Link::Link traffi.cpp:491-511
    l.delta = 1.0 / l.kappa;
    l.dpl = l.delta * l.lanes;
    l.tau = w->p.tau / l.lanes;
    l.w = 1 / l.tau / l.kappa;
    l.capacity = l.vmax * l.w * l.kappa / (l.vmax + l.w);
}

// add_link (bindings.cpp:192-207; Link::Link traffi.cpp:465-537)
UX_API int ux_add_link(ux_world *w, int start, int end, double vmax, double kappa, double length, int lanes, double merge_priority,
                       double cap_out, double cap_in, const int *sg, int nsg) {
    if (w->initialized) return fail(w, UX_ERR_RUNTIME, "add_link after initialize_adj_matrix");
    int N = (int)w->nodes.size();
    if (start < 0 || start >= N || end < 0 || end >= N) return fail(w, UX_ERR_OUT_OF_RANGE, "add_link: node id out of range");
    if (lanes < 1) return fail(w, UX_ERR_INVALID, "add_link: number_of_lanes must be >= 1");
    HLink l;
    l.start = start; l.end = end; l.lanes = lanes; l.length = length; l.vmax = vmax; l.kappa = kappa <= 0.0 ? 0.2 : kappa; l.mp = merge_priority; l.penalty = 0.0;
    link_derive(w, l);
    l.cap_out = cap_out < 0.0 ? l.capacity * 2 : cap_out;
    l.cap_out_rem0 = l.cap_out * w->dt;
    if (cap_in < 0.0) { l.cap_in = l.capacity * 2; l.cap_in_rem0 = l.capacity * 2; } else { l.cap_in = cap_in; l.cap_in_rem0 = cap_in * w->dt; }
    if (nsg <= 0) l.sg = {0}; else l.sg.assign(sg, sg + nsg);
    int id = (int)w->links.size();
    w->nodes[start].out.push_back(id); w->nodes[end].in.push_back(id);
    w->links.push_back(l);
    return id;
}

// add_demand (bindings.cpp:209-216; traffi.cpp:1910-1941)
UX_API int64_t ux_add_demand(ux_world *w, int orig, int dest, double start_t, double end_t, double flow, const int *links_pref, int npref) {
    int N = (int)w->nodes.size();
    if (orig < 0 || orig >= N || dest < 0 || dest >= N) return fail(w, UX_ERR_OUT_OF_RANGE, "add_demand: node id out of range");
    int ts_start = (int)(start_t / w->dt), ts_end = (int)(end_t / w->dt);
    double demand = 0.0; int64_t made = 0;
    for (int ts = ts_start; ts < ts_end; ts++) {
        demand += flow * w->dt;
        while (demand >= (double)w->p.delta_n) {
            if (npref > 0) w->vlinks[1][(long long)w->vehs.size()] = std::vector<int>(links_pref, links_pref + npref);
            w->vehs.push_back({orig, dest, ts});
            demand -= (double)w->p.delta_n; made++;
        }
    }
    return made;
}

// vectorised ingest (include/uxsim_b200.h): add_node / add_link / add_demand applied in index order
UX_API int64_t ux_add_nodes(ux_world *w, int64_t n, const double *x, const double *y, const int *so, const double *sig, const double *sig_offset,
                            const double *fc, const int *lanes) {
    w->nodes.reserve(w->nodes.size() + (size_t)n);
    for (int64_t i = 0; i < n; i++) { int rc = ux_add_node(w, x[i], y[i], sig + so[i], so[i + 1] - so[i], sig_offset[i], fc[i], lanes[i]); if (rc < 0) return rc; }
    return n;
}
UX_API int64_t ux_add_links(ux_world *w, int64_t n, const int *a, const int *b, const double *vmax, const double *kappa, const double *len, const int *lanes,
                            const double *mp, const double *co, const double *ci, const int *go, const int *sg) {
    w->links.reserve(w->links.size() + (size_t)n);
    for (int64_t i = 0; i < n; i++) { int rc = ux_add_link(w, a[i], b[i], vmax[i], kappa[i], len[i], lanes[i], mp[i], co[i], ci[i], sg + go[i], go[i + 1] - go[i]); if (rc < 0) return rc; }
    return n;
}
UX_API int64_t ux_add_demands(ux_world *w, int64_t n, const int *o, const int *d, const double *t0, const double *t1, const double *flow) {
    int64_t made = 0;
    for (int64_t i = 0; i < n; i++) { int64_t m = ux_add_demand(w, o[i], d[i], t0[i], t1[i], flow[i], nullptr, 0); if (m < 0) return m; made += m; }
    return made;
}
UX_API int ux_set_option(ux_world *w, int option, double value) {
    if (option == UX_OPT_KERNEL_TIMING) { w->ktiming = value != 0.0; return 0; }
    if (option == UX_OPT_ALL_DESTINATIONS) {
        if (w->uploaded) return fail(w, UX_ERR_RUNTIME, "UX_OPT_ALL_DESTINATIONS must be set before the first main_loop");
        w->all_dests = value != 0.0; return 0;
    }
    if (option == UX_OPT_RECORD_TRAVERSALS) {
        if (w->uploaded) return fail(w, UX_ERR_RUNTIME, "UX_OPT_RECORD_TRAVERSALS must be set before the first main_loop");
        w->record_traversals = value != 0.0; return 0;
    }
    return fail(w, UX_ERR_INVALID, "unknown option %d", option);
}

UX_API int64_t ux_add_vehicles(ux_world *w, int64_t n, const int *orig, const int *dest, const int *ts) {
    int N = (int)w->nodes.size();
    for (int64_t i = 0; i < n; i++) if (orig[i] < 0 || orig[i] >= N || dest[i] < 0 || dest[i] >= N) return fail(w, UX_ERR_OUT_OF_RANGE, "add_vehicles: node id out of range");
    w->vehs.reserve(w->vehs.size() + (size_t)n);
    for (int64_t i = 0; i < n; i++) w->vehs.push_back({orig[i], dest[i], ts[i]});
    return n;
}

UX_API int ux_set_vehicle_links(ux_world *w, int64_t vid, int kind, const int *links, int n) {
    if (vid < 0 || vid >= (int64_t)w->vehs.size()) return fail(w, UX_ERR_OUT_OF_RANGE, "vehicle index out of range");
    if (kind < 0 || kind > 2) return fail(w, UX_ERR_INVALID, "bad kind");
    if (w->uploaded) return fail(w, UX_ERR_RUNTIME, "set_vehicle_links after the simulation started is not supported");
    for (int i = 0; i < n; i++) if (links[i] < 0 || links[i] >= (int)w->links.size()) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    std::vector<int> v(links, links + (n > 0 ? n : 0));
    if (n > 0) w->vlinks[kind][vid] = v; else w->vlinks[kind].erase(vid);
    if (kind == 0) { if (n > 0) w->vlinks[1][vid] = v; else w->vlinks[1].erase(vid); }  // traffi.cpp:1035-1039
    return 0;
}

UX_API int ux_batch_enforce_routes_csr(ux_world *w, const int *ids, const int64_t *off, int64_t nveh) {
    if (nveh > (int64_t)w->vehs.size()) nveh = (int64_t)w->vehs.size();
    for (int64_t i = 0; i < nveh; i++) if (off[i + 1] > off[i]) { int rc = ux_set_vehicle_links(w, i, 0, ids + off[i], (int)(off[i + 1] - off[i])); if (rc) return rc; }
    return 0;
}

// Enforced routes of ONE replica (the day-to-day solvers run one independent chain per replica): same CSR as
// batch_enforce_routes_csr; for that replica it replaces the world-wide specified_route / links_preferred lists.
UX_API int ux_dta_enforce_routes(ux_world *w, int replica, const int *ids, const int64_t *off, int64_t nveh) {
    if (replica < 0 || replica >= w->R) return fail(w, UX_ERR_OUT_OF_RANGE, "replica %d out of range", replica);
    if (w->uploaded) return fail(w, UX_ERR_RUNTIME, "dta_enforce_routes after the simulation started is not supported");
    int64_t P = (int64_t)w->vehs.size();
    if (nveh > P) nveh = P;
    if (nveh > 0 && off[nveh] > 0x7fffffff) return fail(w, UX_ERR_CAPACITY, "dta_enforce_routes: more than 2^31 route entries");
    for (int64_t e = 0; e < (nveh > 0 ? off[nveh] : 0); e++) if (ids[e] < 0 || ids[e] >= (int)w->links.size()) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    auto &rr = w->rep_routes[replica];
    rr.first.assign(P + 1, 0);
    for (int64_t i = 0; i < nveh; i++) rr.first[i + 1] = (int)off[i + 1];
    for (int64_t i = nveh; i < P; i++) rr.first[i + 1] = rr.first[nveh];
    rr.second.assign(ids, ids + (nveh > 0 ? off[nveh] : 0));
    return 0;
}

UX_API int ux_set_t_max(ux_world *w, double t_max) {
    if (w->uploaded) return fail(w, UX_ERR_RUNTIME, "set_t_max after the simulation started is not supported");
    w->p.t_max = t_max;
    w->total_ts = (int)(t_max / w->dt);
    return 0;
}

UX_API int ux_set_toll_timeseries(ux_world *w, int link, const double *toll, int n) {
    if (link < 0 || link >= (int)w->links.size()) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    if (w->uploaded) return fail(w, UX_ERR_RUNTIME, "set_toll_timeseries after the simulation started is not supported");
    w->links[link].toll.assign(toll, toll + n);
    return 0;
}

// World::initialize_adj_matrix (traffi.cpp:1245-1265): freeze the network.
UX_API int ux_initialize_adj_matrix(ux_world *w) {
    if (w->initialized) return 0;
    if (w->p.adjust_node_capacity) {  // Node::adjust_node_capacity traffi.cpp:212-224
        for (auto &n : w->nodes) {
            if (n.fc < 0.0 && !n.in.empty() && !n.out.empty()) {
                double ci = w->links[n.in[0]].capacity, co = w->links[n.out[0]].capacity;
                for (int l : n.in) ci = std::max(ci, w->links[l].capacity);
                for (int l : n.out) co = std::max(co, w->links[l].capacity);
                n.fc = (ci + co) / 2.0; n.fc_rem0 = n.fc * w->dt;
                if (n.lanes <= 0) n.lanes = (int)std::ceil(n.fc / 0.8);
            }
        }
    }
    for (size_t i = 0; i < w->nodes.size(); i++) {
        auto &n = w->nodes[i];
        if (n.in.size() > UX_MAXDEG || n.out.size() > UX_MAXDEG) return fail(w, UX_ERR_INVALID, "node %d: degree > %d not supported", (int)i, UX_MAXDEG);
        int li = 0, lo = 0;
        for (int l : n.in) li += w->links[l].lanes;
        for (int l : n.out) lo += w->links[l].lanes;
        w->max_node_lanes = std::max(w->max_node_lanes, std::max(li, lo));
        w->max_node_deg = std::max(w->max_node_deg, (int)std::max(n.in.size(), n.out.size()));
        if (li > UX_MAXC || lo > UX_MAXC) return fail(w, UX_ERR_INVALID, "node %d: more than %d lanes in or out not supported", (int)i, UX_MAXC);
        for (int l : n.out) {
            if (w->links[l].lanes > UX_MAXLANE) return fail(w, UX_ERR_INVALID, "link %d: more than %d lanes not supported", l, UX_MAXLANE);
            if (w->links[l].end == (int)i) return fail(w, UX_ERR_INVALID, "link %d: self-loops are not supported", l);
        }
    }
    w->initialized = true;
    return 0;
}

// ---- transfer dependency levels (DESIGN.md "transfer order") --------------------------------------------
static void compute_levels(ux_world *w, std::vector<int> &lflags, std::vector<int> &lvl_off, std::vector<int> &lvl_nodes, std::vector<int> &lvl_level,
                           std::vector<int> &dep_off, std::vector<int> &dep_list) {
    int N = (int)w->nodes.size(), L = (int)w->links.size();
    lflags.assign(L, 0);
    std::vector<int> level(N, 0);
    for (int a = 0; a < N; a++) {  // ascending id: dependencies (smaller ids) are final
        int lv = 0;
        for (int l : w->nodes[a].out) {
            const HLink &k = w->links[l];
            if (k.end < a) {
                lflags[l] |= LF_END_LT_START;
                double bound = 2.0 * k.vmax * w->dt + k.dpl * w->p.delta_n;  // SURVEY 7.3 #1 effects (i) and (ii)
                if (k.length < bound) { lflags[l] |= LF_SEES_POPS; lv = std::max(lv, level[k.end] + 1); }
            }
        }
        level[a] = lv;
    }
    int nl = 0; for (int a = 0; a < N; a++) nl = std::max(nl, level[a] + 1);
    if (N == 0) nl = 1;
    lvl_off.assign(nl + 1, 0);
    for (int a = 0; a < N; a++) lvl_off[level[a] + 1]++;
    for (int i = 0; i < nl; i++) lvl_off[i + 1] += lvl_off[i];
    lvl_nodes.assign(N, 0);
    std::vector<int> fill(nl, 0);
    for (int a = 0; a < N; a++) lvl_nodes[lvl_off[level[a]] + fill[level[a]]++] = a;
    {   // Launch order of the node warps.  A node may only wait for warps with a smaller linear index, so dependencies
        // must come first; beyond that the order is free.  Critical-path order: a node's height = length of the longest
        // chain of dependents hanging off it; sort by (height desc, weight desc), so the chains of waiting nodes start
        // at the very beginning of the launch and the independent bulk fills the machine behind them, heaviest first.
        std::vector<int> weight(N, 0), height(N, 0);
        for (int a = 0; a < N; a++) {
            int wt = 0;
            for (int l : w->nodes[a].in) wt += 2 * w->links[l].lanes + 1;
            for (int l : w->nodes[a].out) wt += w->links[l].lanes + 1;
            weight[a] = wt * (1 + (int)w->nodes[a].out.size());
        }
        for (int a = N - 1; a >= 0; a--)  // a dependency always has the smaller id, so height[a] is final here
            for (int l : w->nodes[a].out) if (lflags[l] & LF_SEES_POPS) { int b = w->links[l].end; height[b] = std::max(height[b], height[a] + 1); }
        std::stable_sort(lvl_nodes.begin(), lvl_nodes.end(), [&](int a, int b) {
            if (height[a] != height[b]) return height[a] > height[b];
            return weight[a] > weight[b]; });
    }
    lvl_level.assign(N, 0);
    for (int q = 0; q < N; q++) lvl_level[q] = level[lvl_nodes[q]];
    {   // bit 16 of a node's entry: it has dependents (it must publish node_done)
        std::vector<char> waited(N, 0);
        for (int a = 0; a < N; a++) for (int l : w->nodes[a].out) if (lflags[l] & LF_SEES_POPS) waited[w->links[l].end] = 1;
        for (int q = 0; q < N; q++) if (waited[lvl_nodes[q]]) lvl_level[q] |= 0x10000;
    }
    dep_off.assign(N + 1, 0); dep_list.clear();
    for (int a = 0; a < N; a++) {
        for (int l : w->nodes[a].out) if (lflags[l] & LF_SEES_POPS) { int b = w->links[l].end; if (std::find(dep_list.begin() + dep_off[a], dep_list.end(), b) == dep_list.end()) dep_list.push_back(b); }
        dep_off[a + 1] = (int)dep_list.size();
    }
    w->n_levels = nl;
}

static int upload_link_params(ux_world *w) {
    int L = (int)w->links.size();
    std::vector<double> vmax(L), dpl_dn(L), udt(L), co(L), ci(L), mp(L), pen(L);
    for (int i = 0; i < L; i++) {
        const HLink &l = w->links[i];
        vmax[i] = l.vmax; dpl_dn[i] = l.dpl * w->p.delta_n; udt[i] = l.vmax * w->dt; co[i] = l.cap_out; ci[i] = l.cap_in; mp[i] = l.mp; pen[i] = l.penalty;
    }
    size_t b = sizeof(double) * L;
    CUDA_OK(cudaMemcpyAsync(w->d_vmax, vmax.data(), b, cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_dpl_dn, dpl_dn.data(), b, cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_udt, udt.data(), b, cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_cap_out, co.data(), b, cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_cap_in, ci.data(), b, cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_mp, mp.data(), b, cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_penalty, pen.data(), b, cudaMemcpyHostToDevice, w->stream));
    std::vector<int> lflags, lvl_off, lvl_nodes, lvl_level, dep_off, dep_list;
    compute_levels(w, lflags, lvl_off, lvl_nodes, lvl_level, dep_off, dep_list);
    dep_list.resize(w->links.size() + 1, 0);
    CUDA_OK(cudaMemcpyAsync(w->d_dep_off, dep_off.data(), sizeof(int) * dep_off.size(), cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_dep_list, dep_list.data(), sizeof(int) * dep_list.size(), cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_lvl_level, lvl_level.data(), sizeof(int) * lvl_level.size(), cudaMemcpyHostToDevice, w->stream));
    for (auto &d : w->hw) d.n_levels = w->n_levels;
    if (w->dworlds) CUDA_OK(cudaMemcpyAsync(w->dworlds, w->hw.data(), sizeof(DevWorld) * w->hw.size(), cudaMemcpyHostToDevice, w->stream));
    lvl_off.resize(w->nodes.size() + 2, lvl_off.back());
    CUDA_OK(cudaMemcpyAsync(w->d_l_flags, lflags.data(), sizeof(int) * L, cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_lvl_off, lvl_off.data(), sizeof(int) * lvl_off.size(), cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_lvl_nodes, lvl_nodes.data(), sizeof(int) * lvl_nodes.size(), cudaMemcpyHostToDevice, w->stream));
    return 0;
}

// Build all device state.  Called lazily by the first main_loop (so that vehicles added after
// initialize_adj_matrix but before the run are included, like the reference).
static int upload(ux_world *w) {
    auto t_begin = std::chrono::steady_clock::now();
    cudaSetDevice(w->device);
    CUDA_OK(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    int N = (int)w->nodes.size(), L = (int)w->links.size(), T = w->total_ts;
    long long P = (long long)w->vehs.size();
    DevWorld base;
    memset(&base, 0, sizeof(base));
    base.N = N; base.L = L; base.T = T; base.P = P;
    base.dn = w->p.delta_n; base.dt = w->dt; base.duo_w = w->p.duo_update_weight; base.duo_time = w->p.duo_update_time; base.noise = w->p.route_choice_uncertainty;
    base.itt = w->p.instantaneous_TT_timestep_interval; base.route_ts = w->route_ts; base.hard_det = w->p.hard_deterministic_mode;
    base.gradual = w->p.route_choice_update_gradual; base.no_cyclic = w->p.no_cyclic_routing; base.log_mode = w->p.vehicle_log_mode;
    base.vis_words = (N + 31) / 32;
    // ---- links
    std::vector<int> l_start(L), l_end(L), l_lanes(L), l_rcap(L), sg_off(L + 1, 0), sg;
    std::vector<long long> l_roff(L);
    std::vector<double> l_length(L);
    long long C = 0;
    std::vector<int> seg_link, seg_start;
    bool any_toll = false;
    for (int i = 0; i < L; i++) {
        const HLink &l = w->links[i];
        l_start[i] = l.start; l_end[i] = l.end; l_lanes[i] = l.lanes; l_length[i] = l.length;
        // jam storage of the link in platoons + entry slack (a platoon is accepted once the one `lanes` ahead has
        // moved one spacing, so at most length/(delta_per_lane*dn) per lane + the ones at x <= spacing)
        double per_lane = l.length / (l.dpl * w->p.delta_n);
        long long need = ((long long)std::floor(per_lane) + 3) * l.lanes + 2, cap = 4;
        while (cap < need) cap <<= 1;  // power of two: slot = position & (cap - 1)
        if (cap > 0x20000000) return fail(w, UX_ERR_INVALID, "link %d: ring too large", i);
        l_rcap[i] = (int)cap; l_roff[i] = C; C += cap;
        {   // work items of the update kernel: one warp per 512 ring slots, at least one per link
            int nparts = (int)((cap + 511) / 512);
            for (int part = 0; part < nparts; part++) { seg_link.push_back(i); seg_start.push_back(part | (nparts << 16)); }
        }
        sg_off[i + 1] = sg_off[i] + (int)l.sg.size();
        sg.insert(sg.end(), l.sg.begin(), l.sg.end());
        if (!l.toll.empty()) any_toll = true;
    }
    w->C = C; w->nseg = (int)seg_link.size();
    base.C = C; base.nseg = w->nseg;
    base.l_start = dupload(w, l_start); base.l_end = dupload(w, l_end); base.l_lanes = dupload(w, l_lanes); base.l_rcap = dupload(w, l_rcap);
    base.l_roff = dupload(w, l_roff); base.l_length = dupload(w, l_length);
    w->d_sg_off = dupload(w, sg_off); w->d_sg = dupload(w, sg);
    base.l_sg_off = w->d_sg_off; base.l_sg = w->d_sg;
    base.seg_link = dupload(w, seg_link); base.seg_start = dupload(w, seg_start);
    w->d_vmax = dalloc<double>(w, L); w->d_dpl_dn = dalloc<double>(w, L); w->d_udt = dalloc<double>(w, L); w->d_cap_out = dalloc<double>(w, L);
    w->d_cap_in = dalloc<double>(w, L); w->d_mp = dalloc<double>(w, L); w->d_penalty = dalloc<double>(w, L);
    w->d_l_flags = dalloc<int>(w, L); w->d_lvl_off = dalloc<int>(w, N + 2); w->d_lvl_nodes = dalloc<int>(w, N); w->d_lvl_level = dalloc<int>(w, N); w->d_dep_off = dalloc<int>(w, N + 1); w->d_dep_list = dalloc<int>(w, L + 1);
    base.l_vmax = w->d_vmax; base.l_dpl_dn = w->d_dpl_dn; base.l_udt = w->d_udt; base.l_cap_out = w->d_cap_out; base.l_cap_in = w->d_cap_in;
    base.l_mp = w->d_mp; base.l_penalty = w->d_penalty; base.l_flags = w->d_l_flags; base.lvl_off = w->d_lvl_off; base.lvl_nodes = w->d_lvl_nodes; base.lvl_level = w->d_lvl_level; base.dep_off = w->d_dep_off; base.dep_list = w->d_dep_list;
    { int rc = upload_link_params(w); if (rc) return rc; }
    if (any_toll) {  // [T][L]; links without a series use their static penalty (traffi.cpp:1285-1289)
        std::vector<double> toll((size_t)T * L);
        for (int t = 0; t < T; t++) for (int i = 0; i < L; i++) { const HLink &l = w->links[i]; toll[(size_t)t * L + i] = (!l.toll.empty() && t < (int)l.toll.size()) ? l.toll[t] : l.penalty; }
        w->d_toll = dupload(w, toll); base.l_toll = w->d_toll;
    }
    // node pairs: links sharing (start, end) are averaged (traffi.cpp:1297-1300)
    std::vector<int> pair_first(L), pair_off(L + 1, 0), pair_list, e_from, e_to, e_link;
    {
        std::map<std::pair<int, int>, int> first;
        std::vector<std::vector<int>> members(L);
        for (int i = 0; i < L; i++) {
            auto key = std::make_pair(w->links[i].start, w->links[i].end);
            auto it = first.find(key);
            if (it == first.end()) { first[key] = i; pair_first[i] = i; } else pair_first[i] = it->second;
            members[pair_first[i]].push_back(i);
        }
        // edges are indexed by e; pair_off/pair_list are indexed by the LEADING LINK id g = e_link[e]
        for (int i = 0; i < L; i++) { pair_off[i + 1] = pair_off[i] + (int)members[i].size(); pair_list.insert(pair_list.end(), members[i].begin(), members[i].end()); }
        for (int i = 0; i < L; i++) if (pair_first[i] == i) { e_from.push_back(w->links[i].start); e_to.push_back(w->links[i].end); e_link.push_back(i); }
    }
    w->n_edges = (int)e_link.size(); base.n_edges = w->n_edges;
    base.l_pair_first = dupload(w, pair_first); base.l_pair_off = dupload(w, pair_off); base.l_pair_list = dupload(w, pair_list);
    base.e_from = dupload(w, e_from); base.e_to = dupload(w, e_to); base.e_link = dupload(w, e_link);
    {
        int Ep = (int)e_link.size();
        std::vector<int> ein_off(N + 1, 0), ein_from(Ep), ein_eid(Ep), fill(N, 0);
        for (int e = 0; e < Ep; e++) ein_off[e_to[e] + 1]++;
        for (int i = 0; i < N; i++) ein_off[i + 1] += ein_off[i];
        for (int e = 0; e < Ep; e++) { int j = e_to[e], q = ein_off[j] + fill[j]++; ein_from[q] = e_from[e]; ein_eid[q] = e; }
        base.ein_off = dupload(w, ein_off); base.ein_from = dupload(w, ein_from); base.ein_eid = dupload(w, ein_eid);
    }
    // ---- nodes
    std::vector<int> out_off(N + 1, 0), in_off(N + 1, 0), out, in, sig_off(N + 1, 0), n_lanes(N), n_out_lanes(N);
    std::vector<double> n_sig, n_fc(N);
    for (int i = 0; i < N; i++) {
        const HNode &n = w->nodes[i];
        out_off[i + 1] = out_off[i] + (int)n.out.size(); in_off[i + 1] = in_off[i] + (int)n.in.size();
        out.insert(out.end(), n.out.begin(), n.out.end()); in.insert(in.end(), n.in.begin(), n.in.end());
        sig_off[i + 1] = sig_off[i] + (int)n.sig.size(); n_sig.insert(n_sig.end(), n.sig.begin(), n.sig.end());
        n_fc[i] = n.fc; n_lanes[i] = n.lanes;
        int tl = 0; for (int l : n.out) tl += w->links[l].lanes; n_out_lanes[i] = tl;
    }
    base.n_out_off = dupload(w, out_off); base.n_out = dupload(w, out); base.n_in_off = dupload(w, in_off); base.n_in = dupload(w, in);
    w->d_n_sig_off = dupload(w, sig_off); w->d_n_sig = dupload(w, n_sig);
    base.n_sig_off = w->d_n_sig_off; base.n_sig = w->d_n_sig; base.n_fc = dupload(w, n_fc); base.n_lanes = dupload(w, n_lanes); base.n_out_lanes = dupload(w, n_out_lanes);
    // ---- vehicles
    std::vector<int> v_orig(P), v_dest(P), v_ts(P), gen_off(N + 1, 0), gen_list(P);
    w->dest_row.assign(N, -1); w->row_dest.clear();
    for (long long i = 0; i < P; i++) { v_orig[i] = w->vehs[i].orig; v_dest[i] = w->vehs[i].dest; v_ts[i] = w->vehs[i].ts; gen_off[v_orig[i] + 1]++; }
    for (int i = 0; i < N; i++) gen_off[i + 1] += gen_off[i];
    {   // per-origin FIFO order = (departure step, creation order): traffi.cpp:1663-1676 + 1804-1808
        std::vector<long long> order(P);
        for (long long i = 0; i < P; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) { return v_ts[a] < v_ts[b]; });
        std::vector<int> fill(N, 0);
        for (long long q = 0; q < P; q++) { long long i = order[q]; gen_list[gen_off[v_orig[i]] + fill[v_orig[i]]++] = (int)i; }
    }
    for (long long i = 0; i < P; i++) if (w->dest_row[v_dest[i]] < 0) w->dest_row[v_dest[i]] = 0;
    if (w->all_dests) for (int i = 0; i < N; i++) w->dest_row[i] = 0;
    w->h_gen_off = gen_off; w->h_gen_list = gen_list; w->eff_ts = v_ts;
    for (int i = 0; i < N; i++) if (w->dest_row[i] == 0) { w->dest_row[i] = (int)w->row_dest.size(); w->row_dest.push_back(i); }
    int D = (int)w->row_dest.size(); w->n_active = D; base.D = D;
    base.v_orig = dupload(w, v_orig); base.v_dest = dupload(w, v_dest); base.v_depart_ts = dupload(w, v_ts);
    base.gen_off = dupload(w, gen_off); base.gen_list = dupload(w, gen_list);
    { std::vector<int> gen_ts(P); for (long long q = 0; q < P; q++) gen_ts[q] = v_ts[gen_list[q]]; base.gen_ts = dupload(w, gen_ts); }
    base.dest_row = dupload(w, w->dest_row); base.row_dest = dupload(w, w->row_dest);
    for (int kind = 0; kind < 3; kind++) {
        if (w->vlinks[kind].empty()) continue;
        std::vector<int> off(P + 1, 0), list;
        for (long long i = 0; i < P; i++) {
            auto it = w->vlinks[kind].find(i);
            if (it != w->vlinks[kind].end()) list.insert(list.end(), it->second.begin(), it->second.end());
            off[i + 1] = (int)list.size();
        }
        base.vl_off[kind] = dupload(w, off); base.vl_list[kind] = dupload(w, list);
    }
    w->P_uploaded = P;
    // ---- per-replica dynamic state
    std::vector<double> co_rem(L), ci_rem(L), sig_t(N), fc_rem(N);
    std::vector<int> sig_ph(N);
    for (int i = 0; i < L; i++) { co_rem[i] = w->links[i].cap_out_rem0; ci_rem[i] = w->links[i].cap_in_rem0; }
    for (int i = 0; i < N; i++) { sig_t[i] = w->nodes[i].sig_t0; sig_ph[i] = w->nodes[i].sig_phase0; fc_rem[i] = w->nodes[i].fc_rem0; }
    std::vector<int> tb(1, w->timestep);
    w->hw.assign(w->R, base);
    for (auto &kv : w->rep_routes) {  // per-replica enforced routes (ux_dta_enforce_routes): specified_route and links_preferred of that replica
        DevWorld &d = w->hw[kv.first];
        const int *o_ = dupload(w, kv.second.first), *l_ = dupload(w, kv.second.second);
        d.vl_off[0] = d.vl_off[1] = o_; d.vl_list[0] = d.vl_list[1] = l_;
    }
    // Everything allocated from here to the end of the replica loop is per-replica DYNAMIC state: zero, except what the
    // ReInit list records.  ux_dta_next_day rewinds a finished world by zeroing these slab ranges and replaying the list.
    dalloc<char>(w, 1);  // make sure the marks below sit inside a slab
    w->dyn_slab0 = w->slabs.size() - 1; w->dyn_used0 = w->slabs.back().used;
    w->reinit.clear();
    auto share = [](const void *p_, size_t b) { auto v = std::make_shared<std::vector<char>>(b); if (b) memcpy(v->data(), p_, b); return v; };
    auto h_co = share(co_rem.data(), 8 * co_rem.size()), h_ci = share(ci_rem.data(), 8 * ci_rem.size()), h_sp = share(sig_ph.data(), 4 * sig_ph.size()),
         h_st = share(sig_t.data(), 8 * sig_t.size()), h_fc = share(fc_rem.data(), 8 * fc_rem.size());
    auto reg_ff = [&](void *q, size_t b) { if (q) { cudaMemsetAsync(q, 0xff, b ? b : 1, w->stream); w->reinit.push_back({q, b ? b : 1, 0, nullptr}); } };
    auto reg_copy = [&](void *q, const std::shared_ptr<std::vector<char>> &src) { if (q && !src->empty()) w->reinit.push_back({q, src->size(), 2, src}); };
    w->d_err_all = dalloc<int>(w, 2 * (size_t)w->R);
    if (!w->d_err_all) return fail(w, UX_ERR_CUDA, "device allocation failed");
    size_t TL = (size_t)T * L;
    for (int r = 0; r < w->R; r++) {
        DevWorld &d = w->hw[r];
        d.seed = w->p.random_seed + r;
        d.head = dalloc<int>(w, 2 * (size_t)L); d.tail = dalloc<int>(w, L); d.tail_k1 = dalloc<int>(w, L); d.pop_k2 = dalloc<int>(w, L); d.hcount = dalloc<int>(w, L);
        d.trips = dalloc<int>(w, L); d.ev_cnt = dalloc<int>(w, L); d.stat_count = dalloc<long long>(w, L);
        d.cap_out_rem = dupload(w, co_rem); d.cap_in_rem = dupload(w, ci_rem);
        reg_copy(d.cap_out_rem, h_co); reg_copy(d.cap_in_rem, h_ci);
        d.X = dalloc<double>(w, 2 * (size_t)C); d.V = dalloc<double>(w, C); d.VID = dalloc<int>(w, C); d.LANE = dalloc<int>(w, C);
        d.sig_phase = dupload(w, sig_ph); d.sig_t = dupload(w, sig_t); d.fc_rem = dupload(w, fc_rem);
        reg_copy(d.sig_phase, h_sp); reg_copy(d.sig_t, h_st); reg_copy(d.fc_rem, h_fc);
        d.gen_head = dalloc<int>(w, N); d.gen_avail = dalloc<int>(w, N);
        d.v_state = dalloc<int>(w, P); d.v_next = dalloc<int>(w, P); d.v_link = dalloc<int>(w, P, false); d.v_arr_ts = dalloc<int>(w, P, false); d.v_trav = dalloc<int>(w, P);
        reg_ff(d.v_link, sizeof(int) * P); reg_ff(d.v_arr_ts, sizeof(int) * P);
        d.v_mr = dalloc<double>(w, P); d.v_atl = dalloc<double>(w, P); d.v_dist = dalloc<double>(w, P); d.v_entry_x = dalloc<double>(w, P);
        d.v_tt = dalloc<double>(w, P);
        if (d.v_tt && P) { UX_LAUNCH(k_fill_f64, dim3((unsigned)((P + 255) / 256)), dim3(256), w->stream, d.v_tt, (long long)P, -1.0); w->reinit.push_back({d.v_tt, (size_t)P * 8, 1, nullptr}); }
        d.v_flags = dalloc<unsigned char>(w, P);
        if (base.log_mode) {
            long long cap = 0;
            for (long long i = 0; i < P; i++) cap += std::max(0, T - w->vehs[i].ts - 1);  // a platoon can be RUN from ts+1 to T-1
            d.lg_cap = cap;
            d.v_gen_ts = dalloc<int>(w, P, false); d.v_end_ts = dalloc<int>(w, P, false); d.v_end_kind = dalloc<unsigned char>(w, P);
            reg_ff(d.v_gen_ts, sizeof(int) * P); reg_ff(d.v_end_ts, sizeof(int) * P);
            d.lg_vid = dalloc<int>(w, cap, false); d.lg_t = dalloc<int>(w, cap, false); d.lg_link = dalloc<int>(w, cap, false); d.lg_lane = dalloc<int>(w, cap, false);
            d.lg_x = dalloc<double>(w, cap, false); d.lg_v = dalloc<double>(w, cap, false); d.lg_s = dalloc<double>(w, cap, false);
            d.lg_count = dalloc<unsigned long long>(w, 1);
            if (!d.lg_vid || !d.lg_s || !d.lg_count || !d.v_gen_ts) return fail(w, UX_ERR_CUDA, "device allocation of the vehicle logs failed (%lld records)", cap);
        }
        if (w->record_traversals) {  // at most one link entry per platoon and step; in practice a few dozen per trip
            long long bound = 0;
            for (long long i = 0; i < P; i++) bound += std::max(0, T - w->vehs[i].ts - 1);
            d.tv_cap = std::min(bound, std::max(P * 48, (long long)1 << 20));
            d.tv_ev = dalloc<int4>(w, (size_t)d.tv_cap, false); d.tv_count = dalloc<unsigned long long>(w, 1);
            if (!d.tv_ev || !d.tv_count) return fail(w, UX_ERR_CUDA, "device allocation of the traversal log failed (%lld events)", d.tv_cap);
        }
        d.v_visited = base.no_cyclic ? dalloc<unsigned>(w, (size_t)P * base.vis_words) : nullptr;
        d.cum_arr = dalloc<double>(w, TL); d.cum_dep = dalloc<double>(w, TL); d.tt_inst = dalloc<double>(w, TL); d.ev_tt = dalloc<double>(w, TL);
        d.ev_ord = dalloc<int>(w, TL); d.sig_log = dalloc<int>(w, (size_t)T * N);
        d.pref = dalloc<double>(w, (size_t)D * L); d.rdist = dalloc<double>(w, (size_t)D * N); d.pair_cost = dalloc<double>(w, w->n_edges);
        d.pref_active = dalloc<int>(w, D); d.has_path = dalloc<int>(w, D); d.rnext = dalloc<int>(w, (size_t)D * N, false);
        reg_ff(d.rnext, sizeof(int) * (size_t)D * N);
        d.err = w->d_err_all + 2 * r; d.t_base = dupload(w, tb); d.node_done = dalloc<int>(w, N + 2); d.n_levels = w->n_levels;
        if (!d.head || !d.X || !d.cum_arr || !d.ev_ord || !d.pref || !d.t_base || !d.v_flags || !d.sig_log || !d.rnext)
            return fail(w, UX_ERR_CUDA, "device allocation failed (replica %d)", r);
    }
    w->dyn_slab1 = w->slabs.size() - 1; w->dyn_used1 = w->slabs.back().used;
    w->rewindable = w->timestep == 0;
    w->dworlds = dalloc<DevWorld>(w, w->R, false);
    if (!w->dworlds) return fail(w, UX_ERR_CUDA, "device allocation failed");
    CUDA_OK(cudaMemcpyAsync(w->dworlds, w->hw.data(), sizeof(DevWorld) * w->R, cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaStreamSynchronize(w->stream));
    w->uploaded = true;
    w->upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return 0;
}

static void launch_sssp(ux_world *w, cudaStream_t st) {
    int N = (int)w->nodes.size();
#ifdef UX_EMU
    UX_LAUNCH(k_sssp<false>, dim3(w->n_active, w->R), dim3(SSSP_THREADS), st, w->dworlds);
#else
    size_t bytes = (size_t)N * 12 + 16, fbytes = (size_t)((N + 3) & ~3) * 17 + 64;
    int use_smem = bytes <= 200 * 1024;
    // small graphs converge in a few dozen whole-edge sweeps (cheaper than one barrier per wavefront); large sparse graphs
    // are dominated by sweeps x E and switch to the worklist kernel
    if (N >= 2048 && N < 65536 && fbytes <= 220 * 1024) {
        if (fbytes > 48 * 1024) cudaFuncSetAttribute(k_sssp_frontier, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fbytes);
        k_sssp_frontier<<<dim3(w->n_active, w->R), dim3(SSSP_THREADS), fbytes, st>>>(w->dworlds);
        return;
    }
    if (use_smem && bytes > 48 * 1024) cudaFuncSetAttribute(k_sssp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (use_smem) k_sssp<true><<<dim3(w->n_active, w->R), dim3(SSSP_THREADS), bytes, st>>>(w->dworlds);
    else k_sssp<false><<<dim3(w->n_active, w->R), dim3(SSSP_THREADS), 0, st>>>(w->dworlds);
#endif
}

#ifdef UX_EMU
#define KT_BEGIN(kid)
#define KT_END(kid)
#else
#define KT_BEGIN(kid)                                                                    \
    if (w->ktiming) { cudaEvent_t a_, b_; cudaEventCreate(&a_); cudaEventCreate(&b_);     \
        w->kt_events.push_back(a_); w->kt_events.push_back(b_); w->kt_ids.push_back(kid); cudaEventRecord(a_, st); }
#define KT_END(kid) if (w->ktiming) cudaEventRecord(w->kt_events.back(), st);
#endif
// Vehicles added between main_loop calls (adddemand after the start; traffi.cpp:1663-1676 rebuilds the departure
// buckets on every call and clamps past-due departures to the first step of the chunk).  All per-platoon device arrays
// grow (old contents copied), the per-origin FIFO lists are rebuilt behind their already-released prefix, and the
// cached graphs are dropped.  Destinations must already have a route-preference row (see UX_OPT_ALL_DESTINATIONS).
}  // extern "C" (templates need C++ linkage)
template <typename T>
static T *dgrow(ux_world *w, T *old, size_t oldn, size_t newn, int fill_byte) {
    T *p = dalloc<T>(w, newn, false);
    if (!p) return nullptr;
    if (fill_byte) cudaMemsetAsync(p, fill_byte, (newn ? newn : 1) * sizeof(T), w->stream);
    if (old && oldn) cudaMemcpyAsync(p, old, oldn * sizeof(T), cudaMemcpyDeviceToDevice, w->stream);
    return p;
}
extern "C" {

static int extend_vehicles(ux_world *w) {
    cudaSetDevice(w->device);
    int N = (int)w->nodes.size(), T = w->total_ts, start_ts = w->timestep;
    long long P0 = w->P_uploaded, P = (long long)w->vehs.size();
    if (!w->rep_routes.empty()) return fail(w, UX_ERR_RUNTIME, "adding vehicles to a world with per-replica routes is not supported");
    w->rewindable = false;  // the per-platoon arrays move out of the recorded dynamic region
    for (long long i = P0; i < P; i++) if (w->dest_row[w->vehs[i].dest] < 0)
        return fail(w, UX_ERR_RUNTIME, "vehicle %lld added after the start has destination node %d, which had no demand before: set UX_OPT_ALL_DESTINATIONS "
                    "(World(..., all_destinations=True)) before the first exec_simulation", i, w->vehs[i].dest);
    w->eff_ts.resize(P);
    for (long long i = P0; i < P; i++) w->eff_ts[i] = std::max(w->vehs[i].ts, start_ts);  // traffi.cpp:1673
    // per-origin lists: keep the released prefix, merge the rest with the new platoons by (effective step, id)
    std::vector<std::vector<int>> lists(N);
    for (int o = 0; o < N; o++) lists[o].assign(w->h_gen_list.begin() + w->h_gen_off[o], w->h_gen_list.begin() + w->h_gen_off[o + 1]);
    std::vector<size_t> released(N, 0);
    for (int o = 0; o < N; o++) { size_t r = 0; while (r < lists[o].size() && w->eff_ts[lists[o][r]] <= start_ts - 1) r++; released[o] = r; }
    for (long long i = P0; i < P; i++) lists[w->vehs[i].orig].push_back((int)i);
    for (int o = 0; o < N; o++)
        std::stable_sort(lists[o].begin() + released[o], lists[o].end(), [&](int a, int b) { return w->eff_ts[a] < w->eff_ts[b]; });
    std::vector<int> gen_off(N + 1, 0), gen_list, gen_ts, v_orig(P), v_dest(P), v_ts(P);
    for (int o = 0; o < N; o++) { gen_off[o + 1] = gen_off[o] + (int)lists[o].size(); gen_list.insert(gen_list.end(), lists[o].begin(), lists[o].end()); }
    gen_ts.resize(gen_list.size());
    for (size_t q = 0; q < gen_list.size(); q++) gen_ts[q] = w->eff_ts[gen_list[q]];
    for (long long i = 0; i < P; i++) { v_orig[i] = w->vehs[i].orig; v_dest[i] = w->vehs[i].dest; v_ts[i] = w->vehs[i].ts; }
    w->h_gen_off = gen_off; w->h_gen_list = gen_list;
    const int *d_orig = dupload(w, v_orig), *d_dest = dupload(w, v_dest), *d_ts = dupload(w, v_ts);
    const int *d_goff = dupload(w, gen_off), *d_glist = dupload(w, gen_list), *d_gts = dupload(w, gen_ts);
    const int *vl_off[3] = {nullptr, nullptr, nullptr}, *vl_list[3] = {nullptr, nullptr, nullptr};
    for (int kind = 0; kind < 3; kind++) {
        if (w->vlinks[kind].empty()) continue;
        std::vector<int> off(P + 1, 0), list;
        for (long long i = 0; i < P; i++) {
            auto it = w->vlinks[kind].find(i);
            if (it != w->vlinks[kind].end()) list.insert(list.end(), it->second.begin(), it->second.end());
            off[i + 1] = (int)list.size();
        }
        vl_off[kind] = dupload(w, off); vl_list[kind] = dupload(w, list);
    }
    long long extra_log = 0;
    for (long long i = P0; i < P; i++) extra_log += std::max(0, T - w->eff_ts[i] - 1);
    for (int r = 0; r < w->R; r++) {
        DevWorld &d = w->hw[r];
        d.P = P; d.v_orig = d_orig; d.v_dest = d_dest; d.v_depart_ts = d_ts; d.gen_off = d_goff; d.gen_list = d_glist; d.gen_ts = d_gts;
        for (int kind = 0; kind < 3; kind++) { d.vl_off[kind] = vl_off[kind]; d.vl_list[kind] = vl_list[kind]; }
        d.v_state = dgrow(w, d.v_state, P0, P, 0); d.v_next = dgrow(w, d.v_next, P0, P, 0); d.v_link = dgrow(w, d.v_link, P0, P, 0xff);
        d.v_arr_ts = dgrow(w, d.v_arr_ts, P0, P, 0xff); d.v_trav = dgrow(w, d.v_trav, P0, P, 0);
        d.v_mr = dgrow(w, d.v_mr, P0, P, 0); d.v_atl = dgrow(w, d.v_atl, P0, P, 0); d.v_dist = dgrow(w, d.v_dist, P0, P, 0); d.v_entry_x = dgrow(w, d.v_entry_x, P0, P, 0);
        d.v_flags = dgrow(w, d.v_flags, P0, P, 0);
        { double *tt = dgrow(w, d.v_tt, P0, P, 0); std::vector<double> m1(P - P0, -1.0); if (tt) cudaMemcpyAsync(tt + P0, m1.data(), sizeof(double) * (P - P0), cudaMemcpyHostToDevice, w->stream); d.v_tt = tt; }
        if (d.no_cyclic) d.v_visited = dgrow(w, d.v_visited, (size_t)P0 * d.vis_words, (size_t)P * d.vis_words, 0);
        if (d.log_mode) {
            d.v_gen_ts = dgrow(w, d.v_gen_ts, P0, P, 0xff); d.v_end_ts = dgrow(w, d.v_end_ts, P0, P, 0xff); d.v_end_kind = dgrow(w, d.v_end_kind, P0, P, 0);
            unsigned long long cnt = 0; cudaMemcpy(&cnt, d.lg_count, sizeof(cnt), cudaMemcpyDeviceToHost);
            long long ncap = d.lg_cap + extra_log;
            d.lg_vid = dgrow(w, d.lg_vid, cnt, ncap, 0); d.lg_t = dgrow(w, d.lg_t, cnt, ncap, 0); d.lg_link = dgrow(w, d.lg_link, cnt, ncap, 0); d.lg_lane = dgrow(w, d.lg_lane, cnt, ncap, 0);
            d.lg_x = dgrow(w, d.lg_x, cnt, ncap, 0); d.lg_v = dgrow(w, d.lg_v, cnt, ncap, 0); d.lg_s = dgrow(w, d.lg_s, cnt, ncap, 0);
            d.lg_cap = ncap;
            if (!d.lg_vid || !d.lg_s) return fail(w, UX_ERR_CUDA, "device allocation failed while growing the vehicle logs");
        }
        if (!d.v_state || !d.v_tt || !d.v_flags || !d.v_entry_x) return fail(w, UX_ERR_CUDA, "device allocation failed while adding vehicles");
    }
    CUDA_OK(cudaMemcpyAsync(w->dworlds, w->hw.data(), sizeof(DevWorld) * w->R, cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaStreamSynchronize(w->stream));
    w->P_uploaded = P;
    w->logs.replica = -1;
    return 0;
}

// enqueue the kernels of `nsteps` timesteps starting at phase (t mod route_ts) = `phase` on the stream
static void enqueue_steps(ux_world *w, int nsteps, int phase, long long *launch_count) {
    int N = (int)w->nodes.size(), R = w->R;
    cudaStream_t st = w->stream;
    dim3 g_nodes(R, (N + XFER_WARPS - 1) / XFER_WARPS), g_update((w->nseg + UPD_WARPS - 1) / UPD_WARPS, R);
    long long lc = 0;
    for (int s = 0; s < nsteps; s++) {
        if (N > 0) {
            KT_BEGIN(UX_K_TRANSFER)
            if (w->max_node_lanes <= 8 && w->max_node_deg <= 8) UX_LAUNCH((k_node<8, 8>), g_nodes, dim3(32 * XFER_WARPS), st, w->dworlds, s);
            else if (w->max_node_lanes <= 32) UX_LAUNCH((k_node<32, 16>), g_nodes, dim3(32 * XFER_WARPS), st, w->dworlds, s);
            else UX_LAUNCH((k_node<UX_MAXC, UX_MAXDEG>), g_nodes, dim3(32 * XFER_WARPS), st, w->dworlds, s);
            KT_END(UX_K_TRANSFER) lc++;
        }
        if (w->nseg > 0) { KT_BEGIN(UX_K_UPDATE) UX_LAUNCH(k_update, g_update, dim3(32 * UPD_WARPS), st, w->dworlds, s); KT_END(UX_K_UPDATE) lc++; }
        bool route = ((phase + s) % w->route_ts) == 0;
        if (w->n_active > 0 && w->n_edges > 0) {
            if (route) {
                KT_BEGIN(UX_K_EDGE_COST) UX_LAUNCH(k_edge_cost, dim3((w->n_edges + 127) / 128, R), dim3(128), st, w->dworlds, s); KT_END(UX_K_EDGE_COST) lc++;
                KT_BEGIN(UX_K_SSSP) launch_sssp(w, st); KT_END(UX_K_SSSP) lc++;
            }
            if (route || w->p.route_choice_update_gradual) {
                long long tot = (long long)w->n_active * (long long)w->links.size();
                KT_BEGIN(UX_K_DUO) UX_LAUNCH(k_duo, dim3((unsigned)((tot + 255) / 256), R), dim3(256), st, w->dworlds, w->p.route_choice_update_gradual ? 1 : 0); KT_END(UX_K_DUO) lc++;
                KT_BEGIN(UX_K_MISC) UX_LAUNCH(k_duo_finish, dim3((w->n_active + 127) / 128, R), dim3(128), st, w->dworlds); KT_END(UX_K_MISC) lc++;
            }
        }
    }
    KT_BEGIN(UX_K_MISC) UX_LAUNCH(k_advance, dim3(1, R), dim3(32), st, w->dworlds, nsteps); KT_END(UX_K_MISC) lc++;
    *launch_count = lc;
}

static const char *derr_text(int code) {
    switch (code) {
    case DERR_RING_FULL: return "link ring buffer overflow (link %d)";
    case DERR_ROUTE_EXHAUSTED: return "Vehicle %d: specified route has no more links";
    case DERR_ROUTE_INCONSISTENT: return "Vehicle %d: specified route is inconsistent; the forced link is not an outgoing link of the current node";
    case DERR_AVOID_ALL: return "Vehicle %d: links_avoid excludes all outgoing links of the current node";
    case DERR_LOG_FULL: return "vehicle log buffer overflow (link %d)";
    default: return "device error (info %d)";
    }
}

// World::main_loop (traffi.cpp:1642-1866)
UX_API int ux_main_loop(ux_world *w, double duration_t, double until_t) {
    if (!w->initialized) { int rc = ux_initialize_adj_matrix(w); if (rc) return rc; }
    int start_ts = w->timestep, end_ts;
    double time_now = w->timestep * w->dt;
    if (duration_t < 0 && until_t < 0) end_ts = w->total_ts;
    else if (duration_t >= 0 && until_t < 0) end_ts = (int)std::floor((duration_t + time_now) / w->dt) + 1;
    else if (duration_t < 0 && until_t >= 0) end_ts = (int)std::floor(until_t / w->dt) + 1;
    else return fail(w, UX_ERR_RUNTIME, "Cannot specify both `duration_t` and `until_t` parameters for `World.main_loop`");
    if (end_ts > w->total_ts) end_ts = w->total_ts;
    if (end_ts <= start_ts) return 0;
    cudaSetDevice(w->device);
    if (!w->uploaded) { int rc = upload(w); if (rc) return rc; }
    if ((long long)w->vehs.size() != w->P_uploaded) { int rc = extend_vehicles(w); if (rc) return rc; }
    if (w->net_dirty) { int rc = upload_link_params(w); if (rc) return rc; w->net_dirty = false; }
    auto t_begin = std::chrono::steady_clock::now();
    // chunks between route updates: (steps, phase); the executable graph of a chunk shape is built once and cached
    std::vector<std::pair<int, int>> chunks;
    for (int t = start_ts; t < end_ts;) {
        int phase = t % w->route_ts;
        int nsteps = std::min(std::min(end_ts - t, w->route_ts - phase), 512);
        chunks.push_back({nsteps, phase});
        t += nsteps;
    }
#ifndef UX_EMU
    auto graph_key = [&](int nsteps, int phase) { return ((long long)nsteps << 32) | (unsigned)phase | ((long long)w->n_levels << 48); };
    if (!w->ktiming)
        for (auto &c : chunks) {
            long long key = graph_key(c.first, c.second), lc = 0;
            if (w->graphs.count(key)) continue;
            cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
            CUDA_OK(cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeThreadLocal));
            enqueue_steps(w, c.first, c.second, &lc);
            CUDA_OK(cudaStreamEndCapture(w->stream, &graph));
            CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
            cudaGraphDestroy(graph);
            w->graphs.emplace(key, std::make_pair(exec, lc));
        }
    w->graph_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    if (!w->ev_begin) { CUDA_OK(cudaEventCreate(&w->ev_begin)); CUDA_OK(cudaEventCreate(&w->ev_end)); }
    CUDA_OK(cudaEventRecord(w->ev_begin, w->stream));
#endif
    for (auto &c : chunks) {
        long long lc = 0;
#ifdef UX_EMU
        enqueue_steps(w, c.first, c.second, &lc);
#else
        if (w->ktiming) enqueue_steps(w, c.first, c.second, &lc);
        else {
            auto &g = w->graphs[graph_key(c.first, c.second)];
            lc = g.second;
            CUDA_OK(cudaGraphLaunch(g.first, w->stream));
        }
#endif
        w->launches += lc;
    }
#ifndef UX_EMU
    CUDA_OK(cudaEventRecord(w->ev_end, w->stream));
#endif
    CUDA_OK(cudaStreamSynchronize(w->stream));
    CUDA_OK(cudaGetLastError());
#ifndef UX_EMU
    { float ms = 0.f; cudaEventElapsedTime(&ms, w->ev_begin, w->ev_end); w->last_loop_ms = ms; }
    for (size_t i = 0; i < w->kt_ids.size(); i++) {
        float ms = 0.f; cudaEventElapsedTime(&ms, w->kt_events[2 * i], w->kt_events[2 * i + 1]);
        w->k_ms[w->kt_ids[i]] += ms; w->k_launches[w->kt_ids[i]] += 1;
        cudaEventDestroy(w->kt_events[2 * i]); cudaEventDestroy(w->kt_events[2 * i + 1]);
    }
    w->kt_events.clear(); w->kt_ids.clear();
#endif
    w->timestep = end_ts;
    {   // device-side error words of all replicas (one block, one copy)
        std::vector<int> e(2 * (size_t)w->R, 0);
        CUDA_OK(cudaMemcpy(e.data(), w->d_err_all, sizeof(int) * e.size(), cudaMemcpyDeviceToHost));
        for (int r = 0; r < w->R; r++) if (e[2 * r]) {
            int c0 = e[2 * r];
            int code = (c0 == DERR_RING_FULL || c0 == DERR_LOG_FULL) ? UX_ERR_CAPACITY : c0 == DERR_ROUTE_EXHAUSTED ? UX_ERR_OUT_OF_RANGE : UX_ERR_INVALID;
            return fail(w, code, derr_text(c0), e[2 * r + 1]);
        }
    }
    w->loop_wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    w->timestep = end_ts;
    return 0;
}

UX_API int ux_check_simulation_ongoing(ux_world *w) { return w->timestep < w->total_ts; }

// ---- interventions ---------------------------------------------------------------------------------------
UX_API int ux_set_link_param(ux_world *w, int link, int field, double value) {
    if (link < 0 || link >= (int)w->links.size()) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    HLink &l = w->links[link];
    double *rem = nullptr; double rv = 0;
    switch (field) {
    case UX_LINK_VMAX: if (value >= 0) { l.vmax = value; l.w = 1.0 / l.tau / l.kappa; l.capacity = l.vmax * l.w * l.kappa / (l.vmax + l.w); l.delta = 1.0 / l.kappa; } break;
    case UX_LINK_KAPPA: if (value >= 0) { l.kappa = value; l.w = 1.0 / l.tau / l.kappa; l.capacity = l.vmax * l.w * l.kappa / (l.vmax + l.w); l.delta = 1.0 / l.kappa; l.dpl = l.delta * l.lanes; } break;
    case UX_LINK_CAPACITY_OUT: l.cap_out = value; break;
    case UX_LINK_CAPACITY_IN: l.cap_in = value; break;
    case UX_LINK_CAPACITY_OUT_REMAIN: l.cap_out_rem0 = value; rem = &l.cap_out_rem0; rv = value; break;
    case UX_LINK_CAPACITY_IN_REMAIN: l.cap_in_rem0 = value; rem = &l.cap_in_rem0; rv = value; break;
    case UX_LINK_MERGE_PRIORITY: l.mp = value; break;
    case UX_LINK_ROUTE_CHOICE_PENALTY: l.penalty = value; break;
    default: return fail(w, UX_ERR_INVALID, "unknown link field %d", field);
    }
    if (w->uploaded) {
        cudaSetDevice(w->device);
        if (rem) for (int r = 0; r < w->R; r++) CUDA_OK(cudaMemcpy((field == UX_LINK_CAPACITY_OUT_REMAIN ? w->hw[r].cap_out_rem : w->hw[r].cap_in_rem) + link, &rv, sizeof(double), cudaMemcpyHostToDevice));
        else w->net_dirty = true;
    }
    return 0;
}

UX_API int ux_set_link_signal_group(ux_world *w, int link, const int *sg, int n) {
    if (link < 0 || link >= (int)w->links.size()) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    if (w->uploaded && n != (int)w->links[link].sg.size()) return fail(w, UX_ERR_RUNTIME, "changing the size of a signal group after the simulation started is not supported");
    w->links[link].sg.assign(sg, sg + n);
    if (w->uploaded) {
        cudaSetDevice(w->device);
        int off = 0; for (int i = 0; i < link; i++) off += (int)w->links[i].sg.size();
        CUDA_OK(cudaMemcpy(w->d_sg + off, sg, sizeof(int) * n, cudaMemcpyHostToDevice));
    }
    return 0;
}

UX_API int ux_set_node_signal(ux_world *w, int node, const double *sig, int nsig, int phase, double sig_t) {
    if (node < 0 || node >= (int)w->nodes.size()) return fail(w, UX_ERR_OUT_OF_RANGE, "node id out of range");
    HNode &n = w->nodes[node];
    bool reset = false;
    if (sig && nsig > 0) {
        if (w->uploaded && nsig != (int)n.sig.size()) return fail(w, UX_ERR_RUNTIME, "changing the number of signal phases after the simulation started is not supported");
        n.sig.assign(sig, sig + nsig);
    }
    if (!w->uploaded) {
        if (sig && nsig > 0 && n.sig_phase0 >= nsig) { n.sig_phase0 = 0; n.sig_t0 = 0; }
        if (phase >= 0) n.sig_phase0 = phase;
        if (sig_t >= 0) n.sig_t0 = sig_t;
        return 0;
    }
    (void)reset;
    cudaSetDevice(w->device);
    if (sig && nsig > 0) { int off = 0; for (int i = 0; i < node; i++) off += (int)w->nodes[i].sig.size(); CUDA_OK(cudaMemcpy(w->d_n_sig + off, sig, sizeof(double) * nsig, cudaMemcpyHostToDevice)); }
    for (int r = 0; r < w->R; r++) {
        if (phase >= 0) CUDA_OK(cudaMemcpy(w->hw[r].sig_phase + node, &phase, sizeof(int), cudaMemcpyHostToDevice));
        if (sig_t >= 0) CUDA_OK(cudaMemcpy(w->hw[r].sig_t + node, &sig_t, sizeof(double), cudaMemcpyHostToDevice));
    }
    return 0;
}

// ---- results ---------------------------------------------------------------------------------------------
static int64_t get_log_array(ux_world *w, int what, int replica, void *dst, int64_t capacity, int *dtype, bool size_only);
static long long dsum_ll(ux_world *w, const long long *d, int n) {
    std::vector<long long> h(n); if (n) cudaMemcpy(h.data(), d, sizeof(long long) * n, cudaMemcpyDeviceToHost);
    long long s = 0; for (auto v : h) s += v; return s;
}
static long long dsum_i(ux_world *w, const int *d, int n) {
    std::vector<int> h(n); if (n) cudaMemcpy(h.data(), d, sizeof(int) * n, cudaMemcpyDeviceToHost);
    long long s = 0; for (auto v : h) s += v; return s;
}

UX_API int64_t ux_get_int(ux_world *w, int what) {
    switch (what) {
    case UX_I_NUM_NODES: return (int64_t)w->nodes.size();
    case UX_I_NUM_LINKS: return (int64_t)w->links.size();
    case UX_I_NUM_VEHICLES: return (int64_t)w->vehs.size();
    case UX_I_TIMESTEP: return w->timestep;
    case UX_I_TOTAL_TIMESTEPS: return w->total_ts;
    case UX_I_NUM_REPLICAS: return w->R;
    case UX_I_ROUTE_UPDATE_STEPS: return w->route_ts;
    case UX_I_PLATOON_UPDATES: return w->uploaded ? dsum_ll(w, w->hw[0].stat_count, (int)w->links.size()) : 0;
    case UX_I_TRIPS_COMPLETED: return w->uploaded ? dsum_i(w, w->hw[0].trips, (int)w->links.size()) : 0;
    case UX_I_LOG_ENTRIES: { if (!w->uploaded || !w->p.vehicle_log_mode) return 0; int dt_; return get_log_array(w, UX_A_LOG_T, 0, nullptr, 0, &dt_, true); }
    case UX_I_KERNEL_LAUNCHES: return w->launches;
    case UX_I_NUM_ACTIVE_DESTS: return w->n_active;
    case UX_I_TRANSFER_LEVELS: return w->n_levels;
    case UX_I_LINK_EVENTS: return w->uploaded ? dsum_i(w, w->hw[0].ev_cnt, (int)w->links.size()) : 0;
    case UX_I_H2D_BYTES: return w->h2d_bytes;
    case UX_I_D2H_BYTES: return w->d2h_bytes;
    case UX_I_RING_SLOTS: return w->C;
    case UX_I_MAX_DEPART_TS: { int m = 0; for (auto &v : w->vehs) m = std::max(m, v.ts); return m; }
    default:
        if (what >= UX_I_KERNEL_LAUNCHES_OF && what < UX_I_KERNEL_LAUNCHES_OF + UX_K_COUNT) return w->k_launches[what - UX_I_KERNEL_LAUNCHES_OF];
        return fail(w, UX_ERR_INVALID, "unknown int query %d", what);
    }
}

UX_API double ux_get_double(ux_world *w, int what) {
    switch (what) {
    case UX_D_DELTA_T: return w->dt;
    case UX_D_TIME: return w->timestep * w->dt;
    case UX_D_T_MAX: return w->p.t_max;
    case UX_D_LAST_LOOP_MS: return w->last_loop_ms;
    case UX_D_DTA_N_SWAP: return w->swaps.empty() ? 0.0 : w->swaps[w->swap_last].n_swap;
    case UX_D_DTA_POTENTIAL_N_SWAP: return w->swaps.empty() ? 0.0 : w->swaps[w->swap_last].potential_n_swap;
    case UX_D_DTA_TOTAL_T_GAP: return w->swaps.empty() ? 0.0 : w->swaps[w->swap_last].total_t_gap;
    case UX_D_DTA_SWAP_MS: return w->swap_ms;
    case UX_D_UPLOAD_MS: return w->upload_ms;
    case UX_D_GRAPH_BUILD_MS: return w->graph_build_ms;
    case UX_D_LAST_LOOP_WALL_MS: return w->loop_wall_ms;
    default:
        if (what >= UX_D_DTA_N_SWAP_OF && what < UX_D_DTA_N_SWAP_OF + 3000) {  // per-replica scalars of a batched route swap
        int kind = (what - UX_D_DTA_N_SWAP_OF) / 1000, r = (what - UX_D_DTA_N_SWAP_OF) % 1000;
        if (r >= (int)w->swaps.size() || !w->swaps[r].valid) return NAN;
        return kind == 0 ? w->swaps[r].n_swap : kind == 1 ? w->swaps[r].potential_n_swap : w->swaps[r].total_t_gap;
    }
    if (what >= UX_D_KERNEL_MS_OF && what < UX_D_KERNEL_MS_OF + UX_K_COUNT) return w->k_ms[what - UX_D_KERNEL_MS_OF];
        return NAN;
    }
}

// ---- per-platoon logs: World::build_all_vehicle_logs_flat_compact (traffi.cpp:1990-2081) + build_enter_log_data (1965-1983) ----
// The device holds only the RUN records; HOME / WAIT / END entries are implied by the departure, generation and end
// steps (the reference reconstructs log_t / log_state from the same counters, traffi.cpp:1117-1133).
static int build_logs(ux_world *w, int replica) {
    auto &G = w->logs;
    if (G.replica == replica && G.timestep == w->timestep) return 0;
    const DevWorld &d = w->hw[replica];
    long long P = d.P; int Tnow = w->timestep; double dt = w->dt;
    std::vector<int> gen(P), endt(P), st(P), arr(P); std::vector<unsigned char> kind(P);
    unsigned long long nrec = 0;
    if (P) {
        CUDA_OK(cudaMemcpy(gen.data(), d.v_gen_ts, sizeof(int) * P, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(endt.data(), d.v_end_ts, sizeof(int) * P, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(kind.data(), d.v_end_kind, P, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(st.data(), d.v_state, sizeof(int) * P, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(arr.data(), d.v_arr_ts, sizeof(int) * P, cudaMemcpyDeviceToHost));
    }
    CUDA_OK(cudaMemcpy(&nrec, d.lg_count, sizeof(nrec), cudaMemcpyDeviceToHost));
    // RUN entries per platoon and their offsets in the time-ordered record array
    std::vector<long long> roff(P + 1, 0);
    for (long long i = 0; i < P; i++) {
        long long r = 0;
        if (gen[i] >= 0) { int last = kind[i] == 1 ? endt[i] : kind[i] == 2 ? endt[i] - 1 : Tnow - 1; r = last - gen[i] + 1; if (r < 0) r = 0; }
        roff[i + 1] = roff[i] + r;
    }
    if ((unsigned long long)roff[P] != nrec) return fail(w, UX_ERR_RUNTIME, "vehicle log bookkeeping mismatch: %lld expected, %llu recorded", roff[P], nrec);
    std::vector<int> r_t(nrec), r_link(nrec), r_lane(nrec); std::vector<double> r_x(nrec), r_v(nrec), r_s(nrec);
    if (nrec) {
        size_t need = sizeof(long long) * (P + 1) + nrec * (3 * sizeof(int) + 3 * sizeof(double)) + 256;
        int rc = 0;
        void *tmp = nullptr;
        CUDA_OK(cudaMalloc(&tmp, need));
        char *b = (char *)tmp;
        long long *d_off = (long long *)b; b += sizeof(long long) * (P + 1);
        double *ox = (double *)b; b += nrec * 8; double *ov = (double *)b; b += nrec * 8; double *os = (double *)b; b += nrec * 8;
        int *ot = (int *)b; b += nrec * 4; int *ol = (int *)b; b += nrec * 4; int *oa = (int *)b;
        cudaMemcpy(d_off, roff.data(), sizeof(long long) * (P + 1), cudaMemcpyHostToDevice);
        UX_LAUNCH(k_log_scatter, dim3((unsigned)((nrec + 255) / 256)), dim3(256), w->stream, w->dworlds, replica, (long long)nrec, d_off, ot, ol, oa, ox, ov, os);
        cudaStreamSynchronize(w->stream);
        cudaMemcpy(r_t.data(), ot, nrec * 4, cudaMemcpyDeviceToHost); cudaMemcpy(r_link.data(), ol, nrec * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(r_lane.data(), oa, nrec * 4, cudaMemcpyDeviceToHost); cudaMemcpy(r_x.data(), ox, nrec * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(r_v.data(), ov, nrec * 8, cudaMemcpyDeviceToHost); cudaMemcpy(r_s.data(), os, nrec * 8, cudaMemcpyDeviceToHost);
        w->d2h_bytes += (long long)(nrec * 36);
        cudaFree(tmp);
        (void)rc;
    }
    G.offsets.assign(P + 1, 0); G.ltl_offsets.assign(P + 1, 0); G.n_missing.assign(P, 0);
    G.t.clear(); G.x.clear(); G.v.clear(); G.s.clear(); G.state.clear(); G.lane.clear(); G.link.clear();
    G.ltl_t.clear(); G.ltl_id.clear(); G.enter_t.clear(); G.enter_link.clear(); G.enter_veh.clear();
    for (long long i = 0; i < P; i++) {
        G.offsets[i] = (long long)G.t.size(); G.ltl_offsets[i] = (long long)G.ltl_t.size();
        int ts0 = w->eff_ts[i];
        G.ltl_t.push_back((double)w->vehs[i].ts * dt); G.ltl_id.push_back(-1);  // home entry (traffi.cpp:2059-2061)
        if (ts0 < Tnow) {
            G.n_missing[i] = ((double)ts0 * dt > 0) ? (int)(((double)ts0 * dt) / dt) : 0;
            auto push = [&](int tstep, int state, double x, double v, double s_, int lane, int link) {
                G.t.push_back((double)tstep * dt); G.state.push_back(state); G.x.push_back(x); G.v.push_back(v); G.s.push_back(s_); G.lane.push_back(lane); G.link.push_back(link);
            };
            push(ts0, UX_VS_HOME, -1, -1, -1.0, -1, -1);
            int wait_end = gen[i] >= 0 ? gen[i] : Tnow;
            for (int tt = ts0 + 1; tt < wait_end; tt++) push(tt, UX_VS_WAIT, -1, -1, -1.0, -1, -1);
            int prev = -999;
            for (long long q = roff[i]; q < roff[i + 1]; q++) {
                push(r_t[q], UX_VS_RUN, r_x[q], r_v[q], r_s[q], r_lane[q], r_link[q]);
                if (r_link[q] >= 0 && r_link[q] != prev) {
                    G.ltl_t.push_back((double)r_t[q] * dt); G.ltl_id.push_back(r_link[q]);
                    G.enter_link.push_back(r_link[q]); G.enter_t.push_back((double)r_t[q] * dt); G.enter_veh.push_back((int)i);
                    prev = r_link[q];
                }
            }
            if (kind[i]) push(endt[i], st[i], -1, -1, -1.0, -1, -1);
        }
        if (st[i] == UX_VS_END) { G.ltl_t.push_back(arr[i] >= 0 ? (double)arr[i] * dt : -1.0); G.ltl_id.push_back(-2); }
    }
    G.offsets[P] = (long long)G.t.size(); G.ltl_offsets[P] = (long long)G.ltl_t.size();
    G.replica = replica; G.timestep = w->timestep;
    return 0;
}

static bool is_log_array(int what) { return (what >= UX_A_LOG_OFFSETS && what <= UX_A_LTL_ID) || (what >= UX_A_ENTER_LINK && what <= UX_A_ENTER_VEH); }

// copy one assembled log array out; returns elements or an error code
static int64_t get_log_array(ux_world *w, int what, int replica, void *dst, int64_t capacity, int *dtype, bool size_only) {
    if (!w->p.vehicle_log_mode) return fail(w, UX_ERR_RUNTIME, "vehicle logs are off (vehicle_log_mode = 0)");
    if (!w->uploaded) { if (!w->initialized) { int rc = ux_initialize_adj_matrix(w); if (rc) return rc; } cudaSetDevice(w->device); int rc = upload(w); if (rc) return rc; }
    if (replica < 0 || replica >= w->R) return fail(w, UX_ERR_OUT_OF_RANGE, "replica %d out of range", replica);
    cudaSetDevice(w->device);
    int rc = build_logs(w, replica); if (rc) return rc;
    auto &G = w->logs;
    const void *src = nullptr; int64_t n = 0; int dt = UX_DT_F64;
    switch (what) {
    case UX_A_LOG_OFFSETS: src = G.offsets.data(); n = (int64_t)G.offsets.size(); dt = UX_DT_I64; break;
    case UX_A_LTL_OFFSETS: src = G.ltl_offsets.data(); n = (int64_t)G.ltl_offsets.size(); dt = UX_DT_I64; break;
    case UX_A_LOG_T: src = G.t.data(); n = (int64_t)G.t.size(); break;
    case UX_A_LOG_X: src = G.x.data(); n = (int64_t)G.x.size(); break;
    case UX_A_LOG_V: src = G.v.data(); n = (int64_t)G.v.size(); break;
    case UX_A_LOG_S: src = G.s.data(); n = (int64_t)G.s.size(); break;
    case UX_A_LOG_STATE: src = G.state.data(); n = (int64_t)G.state.size(); dt = UX_DT_I32; break;
    case UX_A_LOG_LANE: src = G.lane.data(); n = (int64_t)G.lane.size(); dt = UX_DT_I32; break;
    case UX_A_LOG_LINK: src = G.link.data(); n = (int64_t)G.link.size(); dt = UX_DT_I32; break;
    case UX_A_LOG_N_MISSING: src = G.n_missing.data(); n = (int64_t)G.n_missing.size(); dt = UX_DT_I32; break;
    case UX_A_LTL_T: src = G.ltl_t.data(); n = (int64_t)G.ltl_t.size(); break;
    case UX_A_LTL_ID: src = G.ltl_id.data(); n = (int64_t)G.ltl_id.size(); dt = UX_DT_I32; break;
    case UX_A_ENTER_LINK: src = G.enter_link.data(); n = (int64_t)G.enter_link.size(); dt = UX_DT_I32; break;
    case UX_A_ENTER_TIME: src = G.enter_t.data(); n = (int64_t)G.enter_t.size(); break;
    case UX_A_ENTER_VEH: src = G.enter_veh.data(); n = (int64_t)G.enter_veh.size(); dt = UX_DT_I32; break;
    default: return fail(w, UX_ERR_INVALID, "not a log array");
    }
    if (dtype) *dtype = dt;
    if (size_only) return n;
    if (capacity < n) return fail(w, UX_ERR_INVALID, "get_array: capacity %lld < %lld", (long long)capacity, (long long)n);
    size_t es = dt == UX_DT_I32 ? 4 : 8;
    if (n) memcpy(dst, src, es * (size_t)n);
    return n;
}

UX_API int64_t ux_array_size(ux_world *w, int what, int *dtype) {
    int64_t N = (int64_t)w->nodes.size(), L = (int64_t)w->links.size(), T = w->total_ts, P = (int64_t)w->vehs.size();
    int dt = UX_DT_F64; int64_t n = -1;
    if (is_log_array(what)) return get_log_array(w, what, 0, nullptr, 0, dtype, true);
    switch (what) {
    case UX_A_CUM_ARRIVAL: case UX_A_CUM_DEPARTURE: case UX_A_TRAVELTIME_INSTANT: case UX_A_TRAVELTIME_ACTUAL: n = L * T; break;
    case UX_A_VEH_STATE: case UX_A_VEH_ARRIVAL_TS: case UX_A_VEH_DEPART_TS: case UX_A_VEH_ORIG: case UX_A_VEH_DEST: case UX_A_VEH_LINK: case UX_A_VEH_LANE: n = P; dt = UX_DT_I32; break;
    case UX_A_VEH_TRAVEL_TIME: case UX_A_VEH_DISTANCE: case UX_A_VEH_X: case UX_A_VEH_V: n = P; break;
    case UX_A_SIGNAL_LOG: n = N * T; dt = UX_DT_I32; break;
    case UX_A_NODE_SIGNAL_PHASE: n = N; dt = UX_DT_I32; break;
    case UX_A_NODE_SIGNAL_T: n = N; break;
    case UX_A_LINK_NUM_VEHICLES: n = L; dt = UX_DT_I32; break;
    case UX_A_LINK_CAPACITY_IN_REMAIN: case UX_A_LINK_CAPACITY_OUT_REMAIN: case UX_A_LINK_STAT_COUNT: case UX_A_LINK_TRIPS_COMPLETED: n = L; break;
    case UX_A_ROUTE_PREF: n = N * L; break;
    case UX_A_ROUTE_DIST: n = N * N; break;
    case UX_A_ROUTE_NEXT: n = N * N; dt = UX_DT_I32; break;
    case UX_A_DTA_OD_ORIG: case UX_A_DTA_OD_DEST: n = (int64_t)w->dta_p->o.size(); dt = UX_DT_I32; break;
    case UX_A_DTA_OD_OFFSETS: n = (int64_t)w->dta_p->od_off.size(); dt = UX_DT_I64; break;
    case UX_A_DTA_ROUTE_OFFSETS: n = (int64_t)w->dta_p->route_off.size(); dt = UX_DT_I64; break;
    case UX_A_DTA_ROUTE_LINKS: n = (int64_t)w->dta_p->links.size(); dt = UX_DT_I32; break;
    case UX_A_DTA_RS_OFFSETS: case UX_A_DTA_RA_OFFSETS: n = P + 1; dt = UX_DT_I64; break;
    case UX_A_DTA_RS_LINKS: n = 0; for (auto &r : w->swaps) if (r.valid) n = std::max(n, (int64_t)r.rs_total); dt = UX_DT_I32; break;  // upper bound over the replicas
    case UX_A_DTA_RA_LINKS: n = 0; for (auto &r : w->swaps) if (r.valid) n = std::max(n, (int64_t)r.ra_total); dt = UX_DT_I32; break;
    case UX_A_DTA_RA_ENTRY_T: n = 0; for (auto &r : w->swaps) if (r.valid) n = std::max(n, (int64_t)r.ra_total); break;
    case UX_A_DTA_COST_ACTUAL: case UX_A_DTA_T_GAP: n = P; break;
    case UX_A_DTA_CHOICE: n = P; dt = UX_DT_I32; break;
    default: return fail(w, UX_ERR_INVALID, "array %d is not available from the CUDA engine", what);
    }
    if (dtype) *dtype = dt;
    return n;
}

static int ensure_scratch(ux_world *w, size_t bytes) {
    if (bytes <= w->scratch_bytes) return 0;
    size_t want = std::max(bytes, std::max<size_t>(2 * w->scratch_bytes, (size_t)4 << 20));  // from the slabs: no driver allocation in steady state
    char *p = dalloc<char>(w, want, false);
    if (!p) return fail(w, UX_ERR_CUDA, "device allocation failed (%zu bytes of scratch)", want);
    w->d_scratch = (double *)p; w->scratch_bytes = want;
    return 0;
}

// Produce the array `what` of `replica` in DEVICE memory in its ABI layout; *dptr is either state memory or scratch.
static int64_t device_array(ux_world *w, int what, int replica, void **dptr, int *dtype) {
    int dt; int64_t n = ux_array_size(w, what, &dt);
    if (n < 0) return n;
    if (replica < 0 || replica >= w->R) return fail(w, UX_ERR_OUT_OF_RANGE, "replica %d out of range", replica);
    if (!w->initialized) { int rc = ux_initialize_adj_matrix(w); if (rc) return rc; }
    cudaSetDevice(w->device);
    if (!w->uploaded) { int rc = upload(w); if (rc) return rc; }
    const DevWorld &d = w->hw[replica];
    int N = d.N, L = d.L, T = d.T; long long P = d.P;
    cudaStream_t st = w->stream;
    if (dtype) *dtype = dt;
    auto transpose_d = [&](const double *src, int rows, int cols) -> int {
        int rc = ensure_scratch(w, sizeof(double) * (size_t)rows * cols + 16); if (rc) return rc;
        long long tot = (long long)rows * cols;
        if (tot > 0) UX_LAUNCH(k_transpose<double>, dim3((unsigned)((tot + 255) / 256)), dim3(256), st, src, w->d_scratch, rows, cols);
        *dptr = w->d_scratch; return 0; };
    auto transpose_i = [&](const int *src, int rows, int cols) -> int {
        int rc = ensure_scratch(w, sizeof(int) * (size_t)rows * cols + 16); if (rc) return rc;
        long long tot = (long long)rows * cols;
        if (tot > 0) UX_LAUNCH(k_transpose<int>, dim3((unsigned)((tot + 255) / 256)), dim3(256), st, src, (int *)w->d_scratch, rows, cols);
        *dptr = w->d_scratch; return 0; };
    int rc = 0;
    switch (what) {
    case UX_A_CUM_ARRIVAL: rc = transpose_d(d.cum_arr, T, L); break;
    case UX_A_CUM_DEPARTURE: rc = transpose_d(d.cum_dep, T, L); break;
    case UX_A_TRAVELTIME_INSTANT: rc = transpose_d(d.tt_inst, T, L); break;
    case UX_A_TRAVELTIME_ACTUAL:
        rc = ensure_scratch(w, sizeof(double) * (size_t)T * L + 16); if (rc) break;
        if (L > 0) UX_LAUNCH(k_tt_actual, dim3((L + 127) / 128), dim3(128), st, w->dworlds, replica, w->d_scratch);
        *dptr = w->d_scratch; break;
    case UX_A_SIGNAL_LOG: rc = transpose_i(d.sig_log, T, N); break;
    case UX_A_VEH_ARRIVAL_TS: *dptr = d.v_arr_ts; break;
    case UX_A_VEH_DEPART_TS: *dptr = (void *)d.v_depart_ts; break;
    case UX_A_VEH_TRAVEL_TIME: *dptr = d.v_tt; break;
    case UX_A_VEH_ORIG: *dptr = (void *)d.v_orig; break;
    case UX_A_VEH_DEST: *dptr = (void *)d.v_dest; break;
    case UX_A_VEH_LINK: *dptr = d.v_link; break;
    case UX_A_NODE_SIGNAL_PHASE: *dptr = d.sig_phase; break;
    case UX_A_NODE_SIGNAL_T: *dptr = d.sig_t; break;
    case UX_A_LINK_CAPACITY_IN_REMAIN: *dptr = d.cap_in_rem; break;
    case UX_A_LINK_CAPACITY_OUT_REMAIN: *dptr = d.cap_out_rem; break;
    case UX_A_VEH_X: case UX_A_VEH_V: case UX_A_VEH_LANE: {
        rc = ensure_scratch(w, (sizeof(double) * 2 + sizeof(int)) * (size_t)(P ? P : 1) + 64); if (rc) break;
        double *vx = w->d_scratch, *vv = vx + P; int *vl = (int *)(vv + P);
        CUDA_OK(cudaMemsetAsync(vx, 0, sizeof(double) * 2 * P + sizeof(int) * P, st));
        if (w->nseg > 0) UX_LAUNCH(k_veh_snapshot, dim3((w->nseg * 32 + 255) / 256), dim3(256), st, w->dworlds, replica, vx, vv, vl);
        *dptr = what == UX_A_VEH_X ? (void *)vx : what == UX_A_VEH_V ? (void *)vv : (void *)vl; break; }
    default: *dptr = nullptr; break;  // assembled on the host
    }
    if (rc) return rc;
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
    return n;
}

UX_API int64_t ux_get_device_array(ux_world *w, int what, int replica, void **dptr, int *dtype) {
    int64_t n = device_array(w, what, replica, dptr, dtype);
    if (n >= 0 && *dptr == nullptr) return fail(w, UX_ERR_INVALID, "array %d has no device form; use get_array", what);
    return n;
}


// ---------------------------------------------------------------------------------------------------
// DTA solver batch operations
// ---------------------------------------------------------------------------------------------------
}  // extern "C"
namespace {
// Forward shortest-path tree from `start` with the reference's conventions (World::route_search_all traffi.cpp:1317-1385):
// neighbours in ascending node order, strict improvement, ties between queue entries broken by the smaller node id, the
// first hop inherited from the parent.  Only the first-hop row is kept.
struct FirstHopSearch {
    std::vector<double> dist; std::vector<char> done;
    std::priority_queue<std::pair<double, int>, std::vector<std::pair<double, int>>, std::greater<std::pair<double, int>>> pq;
    void run(int N, int start, const std::vector<int> &off, const std::vector<int> &to, const std::vector<double> &wgt, int *hop) {
        dist.assign(N, INFINITY); done.assign(N, 0);
        for (int i = 0; i < N; i++) hop[i] = -1;
        dist[start] = 0.0; hop[start] = start;
        pq.push({0.0, start});
        while (!pq.empty()) {
            int cur = pq.top().second; pq.pop();
            if (done[cur]) continue;
            done[cur] = 1;
            double dcur = dist[cur];
            for (int q = off[cur]; q < off[cur + 1]; q++) {
                if (!(wgt[q] > 0.0)) continue;
                int nx = to[q];
                double nd = dcur + wgt[q];
                if (nd < dist[nx]) { dist[nx] = nd; hop[nx] = cur == start ? nx : hop[cur]; pq.push({nd, nx}); }
            }
        }
    }
};
template <typename F>
void parallel_for(int n, F body) {  // body(thread index, item)
    int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    if (n < 64) nt = 1;
    std::atomic<int> next{0};
    auto work = [&](int tid) { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) body(tid, i); };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
}
}  // namespace
extern "C" {

static DtaStore &dta_new_store(ux_world *w) { w->dta_p = std::make_shared<DtaStore>(); w->swaps.clear(); return *w->dta_p; }

// World.enumerate_od_route_sets_ids_cpp (bindings.cpp:461-465) = World::enumerate_k_random_routes_cpp (traffi.cpp:1387-1481).
// Host side (it runs once per solve): one shortest-path tree per source and round, sources spread over host threads.
UX_API int64_t ux_dta_enumerate_route_sets(ux_world *w, int k, unsigned int seed) {
    if (!w->initialized) { int rc = ux_initialize_adj_matrix(w); if (rc) return rc; }
    int N = (int)w->nodes.size(), L = (int)w->links.size();
    if ((long long)N * N > (1ll << 31)) return fail(w, UX_ERR_CAPACITY, "route enumeration keeps an N x N first-hop table; %d nodes is too many", N);
    // node pairs in (from, to) order; a pair's weight and link id are those of its LAST link (traffi.cpp:1392-1395, 1412-1421)
    std::map<std::pair<int, int>, int> pair_id;
    for (int l = 0; l < L; l++) pair_id[{w->links[l].start, w->links[l].end}] = 0;
    std::vector<int> off(N + 1, 0), to, pair_link;
    { int q = 0; for (auto &kv : pair_id) { kv.second = q++; off[kv.first.first + 1]++; to.push_back(kv.first.second); } }
    for (int i = 0; i < N; i++) off[i + 1] += off[i];
    pair_link.assign(to.size(), -1);
    std::vector<int> link_pair(L);
    for (int l = 0; l < L; l++) { link_pair[l] = pair_id[{w->links[l].start, w->links[l].end}]; pair_link[link_pair[l]] = l; }
    std::vector<double> wgt(to.size(), 0.0);
    std::vector<int> hop((size_t)N * N);
    std::vector<std::vector<std::vector<int>>> routes((size_t)N * N);  // [src*N+dst] -> routes
    std::vector<char> present((size_t)N * N, 0);
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> rdist(0.0, 100.0);
    int nt = 32;
    std::vector<FirstHopSearch> searchers(nt);
    std::vector<std::vector<int>> paths(nt);
    int counter = 0, iteration = 0; double avg_old = 0.0;
    while (true) {
        for (int l = 0; l < L; l++) {
            double base = w->links[l].length / w->links[l].vmax;
            wgt[link_pair[l]] = iteration == 0 ? base : base * rdist(rng);
        }
        parallel_for(N, [&](int tid, int src) { searchers[tid].run(N, src, off, to, wgt, hop.data() + (size_t)src * N); });
        std::vector<long long> tot_routes(N, 0), tot_pairs(N, 0);
        parallel_for(N, [&](int tid, int src) {
            std::vector<int> &path = paths[tid];
            for (int dst = 0; dst < N; dst++) {
                if (src == dst || hop[(size_t)src * N + dst] == -1) continue;
                auto &rs = routes[(size_t)src * N + dst];
                present[(size_t)src * N + dst] = 1;
                tot_pairs[src]++; tot_routes[src] += (long long)rs.size();
                if ((int)rs.size() >= k) continue;
                path.clear();
                int cur = src; bool valid = true;
                while (cur != dst) {  // chain the first hops of the trees rooted at the nodes passed
                    int nx = hop[(size_t)cur * N + dst];
                    if (nx == -1 || nx == cur || (int)path.size() > N) { valid = false; break; }
                    auto it = pair_id.find({cur, nx});
                    if (it == pair_id.end()) { valid = false; break; }
                    path.push_back(pair_link[it->second]);
                    cur = nx;
                }
                if (!valid || path.empty()) continue;
                if (std::find(rs.begin(), rs.end(), path) == rs.end()) rs.push_back(path);
            }
        });
        long long total_routes = 0, total_pairs = 0;
        for (int i = 0; i < N; i++) { total_routes += tot_routes[i]; total_pairs += tot_pairs[i]; }
        double avg = total_pairs > 0 ? (double)total_routes / (double)total_pairs : 0.0;
        if (avg == avg_old) counter++; else { counter = 0; avg_old = avg; }
        iteration++;
        if (counter > 10) break;
    }
    DtaStore &S = dta_new_store(w);
    for (int src = 0; src < N; src++) for (int dst = 0; dst < N; dst++) {
        if (!present[(size_t)src * N + dst]) continue;
        S.o.push_back(src); S.d.push_back(dst);
        for (auto &r : routes[(size_t)src * N + dst]) { S.links.insert(S.links.end(), r.begin(), r.end()); S.route_off.push_back((long long)S.links.size()); }
        S.od_off.push_back((long long)S.route_off.size() - 1);
    }
    return (int64_t)S.route_off.size() - 1;
}

// World.make_od_route_set_store (bindings.cpp:467-489)
UX_API int ux_dta_set_route_sets(ux_world *w, int64_t n_od, const int *orig, const int *dest, const int64_t *od_off, const int64_t *route_off, const int *links) {
    int N = (int)w->nodes.size(), L = (int)w->links.size();
    for (int64_t p = 0; p < n_od; p++) {
        if (orig[p] < 0 || orig[p] >= N || dest[p] < 0 || dest[p] >= N) return fail(w, UX_ERR_OUT_OF_RANGE, "route sets: node id out of range");
        if (p > 0 && (orig[p] < orig[p - 1] || (orig[p] == orig[p - 1] && dest[p] <= dest[p - 1]))) return fail(w, UX_ERR_INVALID, "route sets must be sorted by (orig, dest) without duplicates");
    }
    int64_t n_routes = od_off[n_od], n_links = route_off[n_routes];
    for (int64_t e = 0; e < n_links; e++) if (links[e] < 0 || links[e] >= L) return fail(w, UX_ERR_OUT_OF_RANGE, "route sets: link id out of range");
    DtaStore &S = dta_new_store(w);
    S.o.assign(orig, orig + n_od); S.d.assign(dest, dest + n_od); S.links.assign(links, links + n_links);
    S.od_off.assign(od_off, od_off + n_od + 1); S.route_off.assign(route_off, route_off + n_routes + 1);
    return 0;
}

// Share `from`'s store with `w` (no copy; the device image is uploaded once per solve, not once per day)
UX_API int ux_dta_share_route_sets(ux_world *w, ux_world *from) {
    if (!from) return fail(w, UX_ERR_INVALID, "dta_share_route_sets: no source world");
    if (from->nodes.size() != w->nodes.size() || from->links.size() != w->links.size()) return fail(w, UX_ERR_INVALID, "dta_share_route_sets: the worlds have different networks");
    w->dta_p = from->dta_p;
    w->swaps.clear();
    return 0;
}

// The per-platoon arrays and link lists of a swap result stay on the device until somebody asks for them (a replica-batched
// solve never does: ux_dta_next_day consumes the routes where they are).
static int dta_fetch(ux_world *w, int r, bool want_meta, bool want_rs, bool want_ra) {
    auto &R = w->swaps[r];
    const long long P = R.P;
    if (want_meta && !R.meta_fetched) {
        R.choice.resize(P); R.cost_actual.resize(P); R.t_gap.resize(P); R.ra_off.resize(P + 1); R.rs_off.resize(P + 1);
        std::vector<int> o32(P + 1);
        if (P) { CUDA_OK(cudaMemcpy(R.choice.data(), R.d_choice, 4 * P, cudaMemcpyDeviceToHost)); CUDA_OK(cudaMemcpy(R.cost_actual.data(), R.d_cost, 8 * P, cudaMemcpyDeviceToHost));
                 CUDA_OK(cudaMemcpy(R.t_gap.data(), R.d_gap, 8 * P, cudaMemcpyDeviceToHost)); }
        CUDA_OK(cudaMemcpy(R.ra_off.data(), R.d_tr_off, 8 * (P + 1), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(o32.data(), w->next_routes[r].off, 4 * (P + 1), cudaMemcpyDeviceToHost));
        for (long long i = 0; i <= P; i++) R.rs_off[i] = o32[i];
        R.meta_fetched = true; w->d2h_bytes += (long long)(P * 32 + 16);
    }
    if (want_rs && !R.rs_fetched) {
        size_t n = (size_t)R.rs_total;
        R.rs_links.resize(n);
        if (n) CUDA_OK(cudaMemcpy(R.rs_links.data(), w->next_routes[r].links, 4 * n, cudaMemcpyDeviceToHost));
        R.rs_fetched = true; w->d2h_bytes += (long long)(4 * n);
    }
    if (want_ra && !R.ra_fetched) {
        size_t n = (size_t)R.ra_total;
        std::vector<int> tt(n);
        R.ra_links.resize(n); R.ra_entry_t.resize(n);
        if (n) { CUDA_OK(cudaMemcpy(R.ra_links.data(), R.d_tr_link, 4 * n, cudaMemcpyDeviceToHost)); CUDA_OK(cudaMemcpy(tt.data(), R.d_tr_t, 4 * n, cudaMemcpyDeviceToHost)); }
        for (size_t i = 0; i < n; i++) R.ra_entry_t[i] = (double)tt[i] * w->dt;
        R.ra_fetched = true; w->d2h_bytes += (long long)(8 * n);
    }
    return 0;
}

// World.route_swap_due / route_swap_dso (bindings.cpp:661-706; dta_solver.cpp:130-373).  replica == -1: every replica of
// the world in the same launches (replica r draws its candidates from mt19937(rng_seed + r)).  Everything per-platoon stays
// on the device: offsets are device scans, the totals a device reduction in platoon order; the host sees a few scalars per
// replica plus, for DUE, one byte per platoon (who draws a random number) because the candidate stream is the reference's
// sequential std::mt19937.
UX_API int ux_dta_route_swap(ux_world *w, int replica, int mode, double swap_prob, int swap_num, int has_external, unsigned int rng_seed) {
    if (!w->uploaded || w->timestep == 0) return fail(w, UX_ERR_RUNTIME, "dta_route_swap needs a finished run");
    if (replica < -1 || replica >= w->R) return fail(w, UX_ERR_OUT_OF_RANGE, "replica %d out of range", replica);
    if (!w->record_traversals) return fail(w, UX_ERR_RUNTIME, "dta_route_swap needs UX_OPT_RECORD_TRAVERSALS set before the run");
    if (mode != 0 && mode != 1) return fail(w, UX_ERR_INVALID, "mode must be 0 (DUE) or 1 (DSO)");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    std::vector<int> reps;
    if (replica >= 0) reps.push_back(replica); else for (int r = 0; r < w->R; r++) reps.push_back(r);
    const int nr = (int)reps.size();
    const long long P = w->hw[0].P; const int L = w->hw[0].L, T = w->hw[0].T;
    DtaStore &S = *w->dta_p;
    w->swaps.resize(w->R); w->next_routes.resize(w->R);
    for (int r : reps) w->swaps[r].valid = false;
    for (int r = 0; r < w->R; r++) if (w->swaps[r].valid) { int rc = dta_fetch(w, r, true, false, true); if (rc) return rc; }  // other replicas' results lose their device backing
    if (!S.d_block || S.device != w->device) {  // the store goes to the device once per solve
        if (S.d_block) { cudaSetDevice(S.device); cudaFree(S.d_block); S.d_block = nullptr; cudaSetDevice(w->device); }
        size_t b0 = (4 * S.o.size() + 255) & ~(size_t)255, b1 = (4 * S.links.size() + 255) & ~(size_t)255, b2 = (8 * S.od_off.size() + 255) & ~(size_t)255,
               b3 = (8 * S.route_off.size() + 255) & ~(size_t)255;
        CUDA_OK(cudaMalloc(&S.d_block, 2 * b0 + b1 + b2 + b3 + 256));
        char *b = (char *)S.d_block;
        S.device = w->device;
        S.d_o = (int *)b; S.d_d = (int *)(b + b0); S.d_links = (int *)(b + 2 * b0); S.d_od_off = (long long *)(b + 2 * b0 + b1); S.d_route_off = (long long *)(b + 2 * b0 + b1 + b2);
        cudaMemcpyAsync(S.d_o, S.o.data(), 4 * S.o.size(), cudaMemcpyHostToDevice, st); cudaMemcpyAsync(S.d_d, S.d.data(), 4 * S.d.size(), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(S.d_links, S.links.data(), 4 * S.links.size(), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(S.d_od_off, S.od_off.data(), 8 * S.od_off.size(), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(S.d_route_off, S.route_off.data(), 8 * S.route_off.size(), cudaMemcpyHostToDevice, st);
        CUDA_OK(cudaStreamSynchronize(st));
        w->h2d_bytes += (long long)(8 * S.o.size() + 4 * S.links.size() + 8 * (S.od_off.size() + S.route_off.size()));
    }
    // number of link-entry events per replica (sizes the traversed-route arrays)
    std::vector<unsigned long long> n_ev(nr, 0);
    std::vector<long long> ev_base(nr + 1, 0);
    { int rc = ensure_scratch(w, 8 * (size_t)nr + 16); if (rc) return rc; }
    UX_LAUNCH(k_dta_counts, dim3((nr + 127) / 128), dim3(128), st, w->dworlds, reps[0], nr, (unsigned long long *)w->d_scratch);
    cudaMemcpyAsync(n_ev.data(), w->d_scratch, 8 * (size_t)nr, cudaMemcpyDeviceToHost, st);
    CUDA_OK(cudaStreamSynchronize(st));
    unsigned long long max_ev = 0;
    for (int q = 0; q < nr; q++) {
        if ((long long)n_ev[q] > w->hw[reps[q]].tv_cap) return fail(w, UX_ERR_CAPACITY, "traversal log overflow (replica %d)", reps[q]);
        ev_base[q + 1] = ev_base[q] + (long long)n_ev[q]; max_ev = std::max(max_ev, n_ev[q]);
    }
    const size_t ne = (size_t)ev_base[nr], TL = (size_t)T * L, nP = (size_t)nr * P;
    size_t bytes = 0;
    auto carve = [&](size_t b) { size_t o = bytes; bytes += (b + 255) & ~(size_t)255; return o; };
    const size_t o_tt = carve(8 * TL * nr), o_ext = carve(mode == 1 ? 8 * TL * nr : 0), o_troff = carve(8 * (size_t)nr * (P + 1)), o_trl = carve(4 * ne), o_trt = carve(4 * ne), o_fs = carve(nP),
                 o_pair = carve(4 * nP), o_draw = carve(nP), o_ca = carve(8 * nP), o_tg = carve(8 * nP), o_ch = carve(4 * nP), o_pot = carve(nP), o_dd = carve(sizeof(DtaDev) * nr),
                 o_len = carve(4 * nP), o_gg = carve(sizeof(DtaGather) * nr), o_jobs = carve(sizeof(DtaScanJob) * 2 * nr), o_tot = carve(8 * 4 * (size_t)nr);
    if (bytes + 256 > w->swap_buf_bytes) {  // grown with headroom and kept: a day-to-day solve asks for nearly the same size every day
        if (w->swap_buf) { cudaFree(w->swap_buf); w->swap_buf = nullptr; w->swap_buf_bytes = 0; }
        size_t want = bytes + bytes / 4 + 256;
        CUDA_OK(cudaMalloc((void **)&w->swap_buf, want));
        w->swap_buf_bytes = want;
    }
    char *buf = w->swap_buf;
    auto release = [&](int rc) { return rc; };
    std::vector<DtaDev> hd(nr);
    std::vector<DtaScanJob> jobs(2 * nr);
    for (int q = 0; q < nr; q++) {
        DtaDev &D = hd[q];
        memset(&D, 0, sizeof(D));
        D.tt = (double *)(buf + o_tt) + (size_t)q * TL; D.ext = mode == 1 ? (double *)(buf + o_ext) + (size_t)q * TL : nullptr;
        D.od_o = S.d_o; D.od_d = S.d_d; D.route_links = S.d_links; D.od_off = S.d_od_off; D.route_off = S.d_route_off; D.n_od = (long long)S.o.size();
        D.tr_off = (const long long *)(buf + o_troff) + (size_t)q * (P + 1); D.tr_link = (int *)(buf + o_trl) + ev_base[q]; D.tr_t = (int *)(buf + o_trt) + ev_base[q];
        D.mode = mode; D.has_external = has_external ? 1 : 0; D.replica = reps[q];
        D.flag_swap = (const unsigned char *)(buf + o_fs) + (size_t)q * P; D.pair = (int *)(buf + o_pair) + (size_t)q * P; D.draws = (unsigned char *)(buf + o_draw) + (size_t)q * P;
        D.cost_actual = (double *)(buf + o_ca) + (size_t)q * P; D.t_gap = (double *)(buf + o_tg) + (size_t)q * P; D.choice = (int *)(buf + o_ch) + (size_t)q * P; D.pot = (unsigned char *)(buf + o_pot) + (size_t)q * P;
        auto &NR = w->next_routes[reps[q]];
        NR.ready = false;
        if (!NR.off) { if (cudaMalloc((void **)&NR.off, 4 * (size_t)(P + 1)) != cudaSuccess) return release(fail(w, UX_ERR_CUDA, "device allocation failed (next-day routes)")); }
        jobs[q] = {w->hw[reps[q]].v_trav, (void *)D.tr_off, P, 1, nullptr};                               // traversal counts -> tr_off
        jobs[nr + q] = {(const int *)(buf + o_len) + (size_t)q * P, (void *)NR.off, P, 0, (long long *)(buf + o_tot) + 3 * (size_t)nr + q};  // route lengths -> next-day offsets
    }
    const DtaDev *d_dd = (const DtaDev *)(buf + o_dd);
    const DtaScanJob *d_jobs = (const DtaScanJob *)(buf + o_jobs);
#ifndef UX_EMU
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st);
#endif
    cudaMemcpyAsync(buf + o_dd, hd.data(), sizeof(DtaDev) * nr, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(buf + o_jobs, jobs.data(), sizeof(DtaScanJob) * 2 * nr, cudaMemcpyHostToDevice, st);
    UX_LAUNCH(k_dta_scan, dim3(nr), dim3(1024), st, d_jobs);
    if (max_ev) UX_LAUNCH(k_dta_scatter, dim3((unsigned)((max_ev + 255) / 256), nr), dim3(256), st, w->dworlds, d_dd);
    if (L > 0) UX_LAUNCH(k_dta_tt_actual, dim3((L + 63) / 64, nr), dim3(64), st, w->dworlds, d_dd);
    if (mode == 1 && TL) UX_LAUNCH(k_dta_externality, dim3((unsigned)((TL + 127) / 128), nr), dim3(128), st, w->dworlds, d_dd);
    if (P) UX_LAUNCH(k_dta_prepare, dim3((unsigned)((P + 127) / 128), nr), dim3(128), st, w->dworlds, d_dd);
    // swap candidates: the reference's own stream, std::mt19937(rng_seed) + uniform_real_distribution (dta_solver.cpp:148-149, 265-283)
    std::vector<unsigned char> draws(mode == 0 ? nP : 0), flag(nP, 0);
    if (mode == 0 && nP) {
        cudaMemcpyAsync(draws.data(), buf + o_draw, nP, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) return release(fail(w, UX_ERR_CUDA, "CUDA error in the route-swap preparation: %s", cudaGetErrorString(cudaGetLastError())));
    }
    parallel_for(nr, [&](int, int q) {  // replicas draw independently
        std::mt19937 rng(replica >= 0 ? rng_seed : rng_seed + (unsigned)reps[q]);
        std::uniform_real_distribution<double> udist(0.0, 1.0);
        unsigned char *fl = flag.data() + (size_t)q * P;
        if (mode == 1 && swap_num >= 0) {
            std::vector<int> idx(P);
            for (long long i = 0; i < P; i++) idx[i] = (int)i;
            std::shuffle(idx.begin(), idx.end(), rng);
            long long m = std::min<long long>(swap_num, P);
            for (long long i = 0; i < m; i++) fl[idx[i]] = 1;
        } else if (mode == 1) { for (long long i = 0; i < P; i++) fl[i] = udist(rng) < swap_prob; }
        else { const unsigned char *dr = draws.data() + (size_t)q * P; for (long long i = 0; i < P; i++) if (dr[i]) fl[i] = udist(rng) < swap_prob; }
    });
    std::vector<double> totals(4 * (size_t)nr, 0.0);  // [3*nr] doubles, then nr int64 route-entry totals
    std::vector<long long> rs_tot(nr, 0);
    if (nP) {
        cudaMemcpyAsync(buf + o_fs, flag.data(), nP, cudaMemcpyHostToDevice, st);
        UX_LAUNCH(k_dta_swap, dim3((unsigned)((P + DTA_WARPS - 1) / DTA_WARPS), nr), dim3(32 * DTA_WARPS), st, w->dworlds, d_dd);
        UX_LAUNCH(k_dta_rs_len, dim3((unsigned)((P + 127) / 128), nr), dim3(128), st, w->dworlds, d_dd, (int *)(buf + o_len));
    }
    UX_LAUNCH(k_dta_scan, dim3(nr), dim3(1024), st, d_jobs + nr);
    UX_LAUNCH(k_dta_totals, dim3(nr), dim3(32), st, w->dworlds, d_dd, (double *)(buf + o_tot));
    cudaMemcpyAsync(totals.data(), buf + o_tot, 8 * 4 * (size_t)nr, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return release(fail(w, UX_ERR_CUDA, "CUDA error in the route swap: %s", cudaGetErrorString(cudaGetLastError())));
    memcpy(rs_tot.data(), totals.data() + 3 * (size_t)nr, 8 * (size_t)nr);
    std::vector<DtaGather> hg(nr);
    for (int q = 0; q < nr; q++) {
        auto &R = w->swaps[reps[q]];
        auto &NR = w->next_routes[reps[q]];
        if (rs_tot[q] > 0x7fffffff) return release(fail(w, UX_ERR_CAPACITY, "more than 2^31 route entries (replica %d)", reps[q]));
        size_t need = (size_t)rs_tot[q] + 1;
        if (need > NR.cap) {
            if (NR.links) cudaFree(NR.links);
            NR.cap = need + need / 4; NR.links = nullptr;
            if (cudaMalloc((void **)&NR.links, 4 * NR.cap) != cudaSuccess) { NR.cap = 0; return release(fail(w, UX_ERR_CUDA, "device allocation failed (next-day routes)")); }
        }
        hg[q].off = NR.off; hg[q].links = NR.links;
        R = ux_world::DtaResult();
        R.P = P; R.rs_total = rs_tot[q]; R.ra_total = (long long)n_ev[q];
        R.d_tr_link = hd[q].tr_link; R.d_tr_t = hd[q].tr_t; R.d_tr_off = hd[q].tr_off; R.d_choice = hd[q].choice; R.d_cost = hd[q].cost_actual; R.d_gap = hd[q].t_gap;
        R.n_swap = totals[3 * q + 0]; R.potential_n_swap = totals[3 * q + 1]; R.total_t_gap = totals[3 * q + 2];
        R.replica = reps[q]; R.valid = true;
        w->swap_last = reps[q];
    }
    if (nP) {  // the next day's enforced routes, assembled on the device
        cudaMemcpyAsync(buf + o_gg, hg.data(), sizeof(DtaGather) * nr, cudaMemcpyHostToDevice, st);
        UX_LAUNCH(k_dta_rs_gather, dim3((unsigned)((P * 8 + 255) / 256), nr), dim3(256), st, w->dworlds, d_dd, (const DtaGather *)(buf + o_gg));
    }
#ifndef UX_EMU
    cudaEventRecord(e1, st);
#endif
    if (cudaStreamSynchronize(st) != cudaSuccess) return release(fail(w, UX_ERR_CUDA, "CUDA error assembling the next-day routes: %s", cudaGetErrorString(cudaGetLastError())));
#ifndef UX_EMU
    { float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); w->swap_ms = ms; cudaEventDestroy(e0); cudaEventDestroy(e1); }
#endif
    for (int q = 0; q < nr; q++) w->next_routes[reps[q]].ready = true;
    w->d2h_bytes += (long long)((mode == 0 ? nP : 0) + 32 * (size_t)nr);
    w->h2d_bytes += (long long)(nP + (sizeof(DtaDev) + 2 * sizeof(DtaScanJob) + sizeof(DtaGather)) * nr);
    return release(0);
}

// Rewind a finished world to t = 0 for the next day of a day-to-day solve: the per-replica dynamic state is re-initialised in
// place (zero + the recorded non-zero initial values), every replica that has a route-swap result gets its routes_specified
// as enforced routes straight from device memory, and the static network, the platoon tables and the captured graphs stay.
// Equivalent to building a new world, enforcing the routes and running it -- without the rebuild.
UX_API int ux_dta_next_day(ux_world *w) {
    if (!w->uploaded) return fail(w, UX_ERR_RUNTIME, "dta_next_day needs a finished run");
    if (!w->rewindable) return fail(w, UX_ERR_RUNTIME, "this world cannot be rewound (vehicles were added after the start)");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    CUDA_OK(cudaStreamSynchronize(st));
    for (size_t i = w->dyn_slab0; i <= w->dyn_slab1 && i < w->slabs.size(); i++) {
        size_t a0 = i == w->dyn_slab0 ? w->dyn_used0 : 0, a1 = i == w->dyn_slab1 ? w->dyn_used1 : w->slabs[i].used;
        if (a1 > a0) CUDA_OK(cudaMemsetAsync(w->slabs[i].base + a0, 0, a1 - a0, st));
    }
    for (auto &ri : w->reinit) {
        if (ri.kind == 0) cudaMemsetAsync(ri.dst, 0xff, ri.bytes, st);
        else if (ri.kind == 1) { long long n = (long long)(ri.bytes / 8); UX_LAUNCH(k_fill_f64, dim3((unsigned)((n + 255) / 256)), dim3(256), st, (double *)ri.dst, n, -1.0); }
        else cudaMemcpyAsync(ri.dst, ri.src->data(), ri.bytes, cudaMemcpyHostToDevice, st);
    }
    for (int r = 0; r < w->R && r < (int)w->next_routes.size(); r++) {
        auto &NR = w->next_routes[r];
        if (!NR.ready) continue;
        DevWorld &d = w->hw[r];
        d.vl_off[0] = d.vl_off[1] = NR.off; d.vl_list[0] = d.vl_list[1] = NR.links;
    }
    CUDA_OK(cudaMemcpyAsync(w->dworlds, w->hw.data(), sizeof(DevWorld) * w->R, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));
    w->timestep = 0;
    w->logs.replica = -1;
    return 0;
}

static int64_t get_dta_array(ux_world *w, int what, int replica, void *dst, int64_t capacity) {
    int dt; int64_t n = ux_array_size(w, what, &dt);
    if (n < 0) return n;
    const void *src = nullptr;
    if (what >= UX_A_DTA_RS_OFFSETS) {
        if (replica < 0 || replica >= (int)w->swaps.size() || !w->swaps[replica].valid) return fail(w, UX_ERR_RUNTIME, "no route swap result for replica %d", replica);
        auto &r = w->swaps[replica];
        { int rc = dta_fetch(w, replica, what != UX_A_DTA_RS_LINKS && what != UX_A_DTA_RA_LINKS && what != UX_A_DTA_RA_ENTRY_T, what == UX_A_DTA_RS_LINKS,
                             what == UX_A_DTA_RA_LINKS || what == UX_A_DTA_RA_ENTRY_T); if (rc) return rc; }
        switch (what) {
        case UX_A_DTA_RS_OFFSETS: src = r.rs_off.data(); break;
        case UX_A_DTA_RA_OFFSETS: src = r.ra_off.data(); break;
        case UX_A_DTA_RS_LINKS: src = r.rs_links.data(); n = (int64_t)r.rs_links.size(); break;
        case UX_A_DTA_RA_LINKS: src = r.ra_links.data(); n = (int64_t)r.ra_links.size(); break;
        case UX_A_DTA_RA_ENTRY_T: src = r.ra_entry_t.data(); n = (int64_t)r.ra_entry_t.size(); break;
        case UX_A_DTA_COST_ACTUAL: src = r.cost_actual.data(); break;
        case UX_A_DTA_T_GAP: src = r.t_gap.data(); break;
        default: src = r.choice.data(); break;
        }
    } else {
        DtaStore &d = *w->dta_p;
        src = what == UX_A_DTA_OD_ORIG ? (const void *)d.o.data() : what == UX_A_DTA_OD_DEST ? (const void *)d.d.data() : what == UX_A_DTA_OD_OFFSETS ? (const void *)d.od_off.data()
            : what == UX_A_DTA_ROUTE_OFFSETS ? (const void *)d.route_off.data() : (const void *)d.links.data();
    }
    if (capacity < n) return fail(w, UX_ERR_INVALID, "get_array: capacity %lld < %lld", (long long)capacity, (long long)n);
    if (n > 0) memcpy(dst, src, (dt == UX_DT_I32 ? 4 : 8) * (size_t)n);
    return n;
}

UX_API int64_t ux_get_array(ux_world *w, int what, int replica, void *dst, int64_t capacity) {
    if (is_log_array(what)) return get_log_array(w, what, replica, dst, capacity, nullptr, false);
    if (what >= UX_A_DTA_OD_ORIG && what <= UX_A_DTA_RA_ENTRY_T) return get_dta_array(w, what, replica, dst, capacity);
    if (what == UX_A_VEH_ORIG || what == UX_A_VEH_DEST || what == UX_A_VEH_DEPART_TS) {  // static: answered from the host copy (no upload forced)
        int64_t P = (int64_t)w->vehs.size();
        if (replica < 0 || replica >= w->R) return fail(w, UX_ERR_OUT_OF_RANGE, "replica %d out of range", replica);
        if (capacity < P) return fail(w, UX_ERR_INVALID, "get_array: capacity %lld < %lld", (long long)capacity, (long long)P);
        int *oi = (int *)dst;
        for (int64_t i = 0; i < P; i++) oi[i] = what == UX_A_VEH_ORIG ? w->vehs[i].orig : what == UX_A_VEH_DEST ? w->vehs[i].dest : w->vehs[i].ts;
        return P;
    }
    if (replica == -1) {  // every replica's copy of a per-replica state array, replica-major, in ONE device->host transfer
        const int R = w->R;
        int dt0; int64_t n0 = ux_array_size(w, what, &dt0);
        if (n0 < 0) return n0;
        if (capacity < n0 * R) return fail(w, UX_ERR_INVALID, "get_array: capacity %lld < %lld", (long long)capacity, (long long)(n0 * R));
        if (!w->initialized) { int rc = ux_initialize_adj_matrix(w); if (rc) return rc; }
        cudaSetDevice(w->device);
        if (!w->uploaded) { int rc = upload(w); if (rc) return rc; }
        size_t es0 = dt0 == UX_DT_I32 ? 4 : 8, row = es0 * (size_t)n0;
        std::vector<const void *> src(R, nullptr);
        for (int r = 0; r < R; r++) {
            const DevWorld &d = w->hw[r];
            switch (what) {
            case UX_A_VEH_ARRIVAL_TS: src[r] = d.v_arr_ts; break;
            case UX_A_VEH_TRAVEL_TIME: src[r] = d.v_tt; break;
            case UX_A_VEH_LINK: src[r] = d.v_link; break;
            case UX_A_NODE_SIGNAL_PHASE: src[r] = d.sig_phase; break;
            case UX_A_NODE_SIGNAL_T: src[r] = d.sig_t; break;
            case UX_A_LINK_CAPACITY_IN_REMAIN: src[r] = d.cap_in_rem; break;
            case UX_A_LINK_CAPACITY_OUT_REMAIN: src[r] = d.cap_out_rem; break;
            default: break;
            }
        }
        if (what == UX_A_LINK_STAT_COUNT) {  // i64 on the device, f64 in the ABI
            int rc = ensure_scratch(w, 8 * (size_t)n0 * R + 16); if (rc) return rc;
            for (int r = 0; r < R; r++) cudaMemcpyAsync((char *)w->d_scratch + row * r, w->hw[r].stat_count, row, cudaMemcpyDeviceToDevice, w->stream);
            std::vector<long long> h((size_t)n0 * R);
            if (n0 * R) CUDA_OK(cudaMemcpyAsync(h.data(), w->d_scratch, row * R, cudaMemcpyDeviceToHost, w->stream));
            CUDA_OK(cudaStreamSynchronize(w->stream));
            double *o = (double *)dst;
            for (size_t i = 0; i < h.size(); i++) o[i] = (double)h[i];
        } else {
            if (!src[0]) return fail(w, UX_ERR_INVALID, "array %d cannot be fetched for all replicas at once", what);
            int rc = ensure_scratch(w, row * R + 16); if (rc) return rc;
            for (int r = 0; r < R; r++) cudaMemcpyAsync((char *)w->d_scratch + row * r, src[r], row, cudaMemcpyDeviceToDevice, w->stream);
            if (n0 * R) CUDA_OK(cudaMemcpyAsync(dst, w->d_scratch, row * R, cudaMemcpyDeviceToHost, w->stream));
            CUDA_OK(cudaStreamSynchronize(w->stream));
        }
        w->d2h_bytes += (long long)(row * R);
        return n0 * R;
    }
    void *dptr = nullptr; int dt = 0;
    int64_t n = device_array(w, what, replica, &dptr, &dt);
    if (n < 0) return n;
    if (capacity < n) return fail(w, UX_ERR_INVALID, "get_array: capacity %lld < %lld", (long long)capacity, (long long)n);
    size_t es = dt == UX_DT_F64 ? 8 : dt == UX_DT_I32 ? 4 : 8;
    if (dptr) { if (n > 0) CUDA_OK(cudaMemcpy(dst, dptr, es * (size_t)n, cudaMemcpyDeviceToHost)); w->d2h_bytes += (long long)(es * (size_t)n); return n; }
    const DevWorld &d = w->hw[replica];
    int N = d.N, L = d.L; long long P = d.P;
    double *o = (double *)dst; int *oi = (int *)dst;
    switch (what) {
    case UX_A_VEH_STATE: {  // HOME/WAIT are derived from the departure step (traffi.cpp:836-841)
        std::vector<int> st(P); if (P) CUDA_OK(cudaMemcpy(st.data(), d.v_state, sizeof(int) * P, cudaMemcpyDeviceToHost));
        for (long long i = 0; i < P; i++) {
            int s = st[i];
            if (s == UX_VS_HOME && w->eff_ts[i] <= w->timestep - 1) s = UX_VS_WAIT;
            oi[i] = s;
        }
        break; }
    case UX_A_VEH_DISTANCE: {
        std::vector<double> dist(P), ex(P), vx(P); std::vector<int> st(P);
        void *px = nullptr; int64_t m = device_array(w, UX_A_VEH_X, replica, &px, nullptr); if (m < 0) return m;
        if (P) { CUDA_OK(cudaMemcpy(vx.data(), px, sizeof(double) * P, cudaMemcpyDeviceToHost)); CUDA_OK(cudaMemcpy(dist.data(), d.v_dist, sizeof(double) * P, cudaMemcpyDeviceToHost));
                 CUDA_OK(cudaMemcpy(ex.data(), d.v_entry_x, sizeof(double) * P, cudaMemcpyDeviceToHost)); CUDA_OK(cudaMemcpy(st.data(), d.v_state, sizeof(int) * P, cudaMemcpyDeviceToHost)); }
        for (long long i = 0; i < P; i++) o[i] = dist[i] + (st[i] == UX_VS_RUN ? vx[i] - ex[i] : 0.0);
        break; }
    case UX_A_LINK_NUM_VEHICLES: {
        std::vector<int> hd(2 * (size_t)L), tl(L);
        if (L) { CUDA_OK(cudaMemcpy(hd.data(), d.head, sizeof(int) * 2 * L, cudaMemcpyDeviceToHost)); CUDA_OK(cudaMemcpy(tl.data(), d.tail, sizeof(int) * L, cudaMemcpyDeviceToHost)); }
        int cur = w->timestep & 1;
        for (int l = 0; l < L; l++) oi[l] = tl[l] - hd[(size_t)cur * L + l];
        break; }
    case UX_A_LINK_STAT_COUNT: { std::vector<long long> h(L); if (L) CUDA_OK(cudaMemcpy(h.data(), d.stat_count, sizeof(long long) * L, cudaMemcpyDeviceToHost)); for (int l = 0; l < L; l++) o[l] = (double)h[l]; break; }
    case UX_A_LINK_TRIPS_COMPLETED: { std::vector<int> h(L); if (L) CUDA_OK(cudaMemcpy(h.data(), d.trips, sizeof(int) * L, cudaMemcpyDeviceToHost)); for (int l = 0; l < L; l++) o[l] = (double)h[l]; break; }
    case UX_A_ROUTE_PREF: {
        std::vector<double> h((size_t)d.D * L); if (!h.empty()) CUDA_OK(cudaMemcpy(h.data(), d.pref, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
        for (int k = 0; k < N; k++) { int r = w->dest_row[k]; for (int l = 0; l < L; l++) o[(size_t)k * L + l] = r >= 0 ? h[(size_t)r * L + l] : 0.0; }
        break; }
    case UX_A_ROUTE_DIST: {
        std::vector<double> h((size_t)d.D * N); if (!h.empty()) CUDA_OK(cudaMemcpy(h.data(), d.rdist, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
        bool searched = w->timestep > 0;
        for (int i = 0; i < N; i++) for (int k = 0; k < N; k++) { int r = w->dest_row[k]; o[(size_t)i * N + k] = (r >= 0 && searched) ? h[(size_t)r * N + i] : INFINITY; }
        break; }
    case UX_A_ROUTE_NEXT: {
        std::vector<int> h((size_t)d.D * N); if (!h.empty()) CUDA_OK(cudaMemcpy(h.data(), d.rnext, sizeof(int) * h.size(), cudaMemcpyDeviceToHost));
        for (int i = 0; i < N; i++) for (int k = 0; k < N; k++) { int r = w->dest_row[k]; oi[(size_t)i * N + k] = r >= 0 ? h[(size_t)r * N + i] : -1; }
        break; }
    default: return fail(w, UX_ERR_INVALID, "array %d is not available from the CUDA engine", what);
    }
    return n;
}

UX_API int64_t ux_ensemble_link_stats_device(ux_world *w, void **dptr) {
    if (!w->uploaded) return fail(w, UX_ERR_RUNTIME, "ensemble_link_stats before the simulation ran");
    cudaSetDevice(w->device);
    int L = (int)w->links.size();
    if (!w->d_ens) w->d_ens = dalloc<double>(w, (size_t)L * 4);
    if (L > 0) {
        int rc = ensure_scratch(w, 8 * (size_t)L * 4 * w->R + 16); if (rc) return rc;
        UX_LAUNCH(k_ensemble_partial, dim3((L + 63) / 64, w->R), dim3(64), w->stream, w->dworlds, w->d_scratch);
        UX_LAUNCH(k_ensemble_stats, dim3((L * 4 + 127) / 128), dim3(128), w->stream, w->dworlds, w->R, (const double *)w->d_scratch, w->d_ens);
    }
    CUDA_OK(cudaStreamSynchronize(w->stream));
    *dptr = w->d_ens;
    return (int64_t)L * 4;
}

UX_API int64_t ux_ensemble_link_stats(ux_world *w, double *dst, int64_t capacity) {
    void *d = nullptr;
    int64_t n = ux_ensemble_link_stats_device(w, &d);
    if (n < 0) return n;
    if (capacity < n) return fail(w, UX_ERR_INVALID, "capacity too small");
    if (n > 0) CUDA_OK(cudaMemcpy(dst, d, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return n;
}

UX_API const char *ux_last_error(ux_world *w) { return w ? w->err.c_str() : g_err.c_str(); }

}  // extern "C"
