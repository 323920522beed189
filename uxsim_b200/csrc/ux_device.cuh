// ux_device.cuh -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// Limits, the per-replica device world (DevWorld), warp-phase macros and the device helpers shared by the kernels.

// ---------------------------------------------------------------------------------------------------
// limits
// ---------------------------------------------------------------------------------------------------
#define UX_MAXDEG 16   // max out-/in-degree of a node
#define UX_MAXC 64     // max sum of lanes over the in-links (or the out-links) of one node
#define UX_SEG 32      // ring slots per update-kernel work item (one warp)

enum { LF_SEES_POPS = 1, LF_END_LT_START = 2 };            // link flags
enum { VF_WAIT_END = 1, VF_ABORT = 2, VF_CAND = 4 };       // vehicle flags
enum { DERR_NONE = 0, DERR_RING_FULL = 1, DERR_ROUTE_EXHAUSTED = 2, DERR_ROUTE_INCONSISTENT = 3, DERR_AVOID_ALL = 4, DERR_LOG_FULL = 5, DERR_DEP_TIMEOUT = 6 };
#define UX_DEP_SPIN_BUDGET (1ll << 24)  // x >= 64 ns: about 2 s before a dependency wait gives up with DERR_DEP_TIMEOUT

// ---------------------------------------------------------------------------------------------------
// device world
// ---------------------------------------------------------------------------------------------------
// 16-byte records of the replica-minor dynamic state.  LOADS may take a whole record (one 128-bit access); STORES always go
// to single members, because different members of one record belong to different writers (tail: the link's start node,
// head: the link update, pop_k2: the end node ...).
struct LinkQ { int head[2], tail, tail_k1; };  // monotone queue counters: head double-buffered by step parity, tail, tail after this step's generation
struct LinkC { int pop_k2, hcount, trips, ev_cnt; };  // pops by this step's transfer, head positions the next transfer must look at, finished trips, exit events
struct LinkR { double cap_out_rem, cap_in_rem; };
struct NodeS { int sig_phase, gen_head, gen_avail, node_act; };  // node_act: (step + 1) of the last link update that left a candidate / trip-end waiter at the head of an in-link
struct NodeR { double sig_t, fc_rem; };

struct DevWorld {
    int N, L, T, D, vis_words;
    long long P, C;
    int nseg;
    double dn, dt, duo_w, duo_time, noise;
    int itt, route_ts, hard_det, gradual, no_cyclic, log_mode;
    int node_rl;    // the node step runs in replica-lane form (k_node_rl): k_link_pre then also measures the mean-speed travel times
    int pre_split;  // the scalar link / node bookkeeping of a step runs in k_link_pre (large worlds) instead of inside k_node
    long long seed;
    // static network (shared by all replicas)
    const int *l_start, *l_end, *l_lanes, *l_rcap, *l_flags, *l_sg_off, *l_sg, *l_pair_first, *l_pair_off, *l_pair_list;
    const long long *l_roff;
    const double *l_tt0;  // length / vmax at initialize_adj_matrix: what traveltime_actual holds before the first exit event
    const double *l_length, *l_vmax, *l_dpl_dn, *l_udt, *l_cap_out, *l_cap_in, *l_mp, *l_penalty, *l_toll;
    const int *n_out_off, *n_out, *n_in_off, *n_in, *n_sig_off, *n_lanes, *n_out_lanes, *n_out_woff, *lvl_off, *lvl_nodes, *lvl_level, *dep_off, *dep_list;
    int n_levels;
    const double *n_sig, *n_fc;
    const int *v_orig, *v_dest, *v_depart_ts, *gen_off, *gen_list, *gen_ts, *dest_row, *row_dest;
    const int *vl_off[3], *vl_list[3];  // per-vehicle CSR: 0 specified_route, 1 links_preferred, 2 links_avoid
    const int *seg_link, *seg_start;
    const int *e_from, *e_to, *e_link;  // node-pair edges (leading link of each pair)
    const int *ein_off, *ein_from, *ein_eid;  // the same edges grouped by their END node (CSR), for the frontier search
    const int *eout_off, *eout_to;            // ... and grouped by their START node, end nodes ascending (next-hop pass)
    const int *e_inpos, *e_outpos;            // position of edge e in the two CSRs (k_edge_cost scatters the costs there)
    int n_edges;
    // dynamic, REPLICA-MINOR: the per-link / per-node scalars of all replicas interleaved -- element (entity e, replica r) lives at
    // base[e * RS + r]; each replica's DevWorld holds base + r, so `field[e * RS]` is its own element.  A kernel whose lanes are
    // the replicas of one entity (k_link_pre, k_node_rl, k_update_rl) reads them fully coalesced; a kernel whose lanes walk the
    // entities of one replica sees the old layout at RS == 1.  The scalars are packed into 16-byte records by who reads them together.
    int RS, rep;            // replica stride (elements); this replica's index
    long long rep_stride;   // bytes between the replica-major arrays (rings, per-platoon state, route tables, logs) of consecutive replicas; 0 = not uniform
    LinkQ *lq;              // queue counters of a link
    LinkC *lc;              // per-step / cumulative counts of a link
    LinkR *lr;              // capacity tokens of the two link sides
    NodeS *ns;              // signal phase, vertical-queue cursors, activity step of a node
    NodeR *nr;              // signal clock, node capacity tokens
    long long *stat_count;
    // dynamic, replica-major (contiguous per replica)
    double *X, *V;
    int *VID, *LANE;
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
    double *sssp_delta;          // bucket width of the near-far shortest-path search (k_cost_stats)
    double *cost_in, *cost_out;  // pair_cost in ein / eout order with unusable edges (cost <= 0, inf) already +inf
    double *cost_hist;           // [route updates][n_edges]: pair_cost of every update (World::route_dist_record is rebuilt from it on demand)
    int *pref_active, *has_path, *rnext;
    // link-entry events (platoon, ordinal of the link on its route, link, step) in arrival order; only with UX_OPT_RECORD_TRAVERSALS
    int4 *tv_ev; unsigned long long *tv_count; long long tv_cap;
    int *node_done;  // per node: (timestep + 1) of the last finished transfer -- monotone, so it never needs a reset
    int *err;     // [0] code, [1] info
    int *t_base;  // timestep at the start of the running graph
};

#define UX_DEV __device__ __forceinline__

// Index of (step t, link l) in the per-replica link series (cumulative curves, instantaneous travel time, travel-time
// event tables) and of (step t, node n) in the signal log.  Time-major rows [T][L]; the getters transpose to the ABI's [L][T].
// Link-major (-DUX_LINK_MAJOR: a link's consecutive steps share 32-byte sectors, no transpose at hand-back) was measured
// 2-3 % SLOWER on the B200 (Chicago x32: 79.1 vs 77.4 ms; at deltan=5: 430 vs 416 ms): the loop is bound by dependent-load
// latency, not by the DRAM sectors this layout saves.
// The series are replica-minor as well ([T][L][RS]): a step writes one contiguous row for all links and replicas.
#ifndef UX_LINK_MAJOR
#define UX_TIME_MAJOR 1
#define TLI(W, t, l) (((size_t)(t) * (size_t)(W).L + (size_t)(l)) * (size_t)(W).RS)
#define TNI(W, t, n) (((size_t)(t) * (size_t)(W).N + (size_t)(n)) * (size_t)(W).RS)
#else
#define TLI(W, t, l) (((size_t)(l) * (size_t)(W).T + (size_t)(t)) * (size_t)(W).RS)
#define TNI(W, t, n) (((size_t)(n) * (size_t)(W).T + (size_t)(t)) * (size_t)(W).RS)
#endif
// replica-minor per-link / per-node state of the replica W belongs to
#define LQ(W, l) ((W).lq[(size_t)(l) * (size_t)(W).RS])
#define LC(W, l) ((W).lc[(size_t)(l) * (size_t)(W).RS])
#define LR(W, l) ((W).lr[(size_t)(l) * (size_t)(W).RS])
#define LSTAT(W, l) ((W).stat_count[(size_t)(l) * (size_t)(W).RS])
#define NS(W, n) ((W).ns[(size_t)(n) * (size_t)(W).RS])
#define NR(W, n) ((W).nr[(size_t)(n) * (size_t)(W).RS])
#define NDONE(W, n) ((W).node_done[(size_t)(n) * (size_t)(W).RS])

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
// S supplies the static tables (network, platoon destinations), W the replica's own state (route lists, traversal counts,
// visited sets, preference rows, seed, error word).  They are the same struct except in the replica-lane kernels, where S is the
// lane group's world staged in shared memory and W the lane's own pointer table in global memory (read for a handful of pointers only).
UX_DEV int route_choice(const DevWorld &S, const DevWorld &W, int vid, int node, int t, unsigned entity, unsigned purpose, unsigned draw, int *deferred_err = nullptr) {
    int o0 = S.n_out_off[node], nset = S.n_out_off[node + 1] - o0;
    if (nset == 0) return -1;
    const int *linkset = S.n_out + o0;
    const int *vo0 = W.vl_off[0], *vo1 = W.vl_off[1], *vo2 = W.vl_off[2];
    if (vo0) {  // specified_route
        int r0 = vo0[vid], len = vo0[vid + 1] - r0;
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
    if (vo1) {  // links_preferred: intersection if non-empty
        int r0 = vo1[vid], len = vo1[vid + 1] - r0;
        if (len > 0) {
            int m = 0;
            for (int i = 0; i < n; i++) if (list_has(W.vl_list[1] + r0, len, out[i])) tmp[m++] = out[i];
            if (m > 0) { int *s = out; out = tmp; tmp = s; n = m; }
        }
    }
    if (vo2) {  // links_avoid: subtraction, error if nothing is left
        int r0 = vo2[vid], len = vo2[vid + 1] - r0;
        if (len > 0) {
            int m = 0;
            for (int i = 0; i < n; i++) if (!list_has(W.vl_list[2] + r0, len, out[i])) tmp[m++] = out[i];
            if (m != n) { int *s = out; out = tmp; tmp = s; n = m; }
            if (n == 0) { if (deferred_err) *deferred_err = DERR_AVOID_ALL; else dev_error(W, DERR_AVOID_ALL, vid); return -1; }
        }
    }
    if (S.no_cyclic) {
        const unsigned *vis = W.v_visited + (size_t)vid * S.vis_words;
        int m = 0;
        for (int i = 0; i < n; i++) { int e = S.l_end[out[i]]; if (!((vis[e >> 5] >> (e & 31)) & 1u)) tmp[m++] = out[i]; }
        if (m > 0) { int *s = out; out = tmp; tmp = s; n = m; }
    }
    if (n == 0) return -1;
    double pw[UX_MAXDEG];
    int row = S.dest_row[S.v_dest[vid]];
    const double *pref = W.pref + (size_t)(row >= 0 ? row : 0) * S.L;
    for (int i = 0; i < n; i++) pw[i] = row >= 0 ? pref[out[i]] : 0.0;
    if (S.hard_det) {
        int best = 0;
        for (int i = 1; i < n; i++) if (pw[i] > pw[best]) best = i;
        return out[best];
    }
    double u = ux_rng_uniform(W.seed, (uint32_t)t, entity, purpose, draw);
    return out[weighted_pick(pw, n, u)];
}
UX_DEV int route_choice(const DevWorld &W, int vid, int node, int t, unsigned entity, unsigned purpose, unsigned draw, int *deferred_err = nullptr) {
    return route_choice(W, W, vid, node, t, entity, purpose, draw, deferred_err);
}

// acceptance test of an out-link seen with effective head `hd` (traffi.cpp:130-142, 285-297)
UX_DEV bool can_accept(const DevWorld &W, int l, int hd, int cur) {
    int sz = LQ(W, l).tail - hd, lanes = W.l_lanes[l];
    bool ok = false;
    if (sz < lanes) ok = true;
    else {
        int slot = ring_slot(hd, sz - lanes, W.l_rcap[l]);
        if (W.X[(size_t)cur * W.C + W.l_roff[l] + slot] > W.l_dpl_dn[l]) ok = true;
    }
    if (LR(W, l).cap_in_rem < W.dn) ok = false;
    return ok;
}

// push a platoon at the tail of link l (lane + leader rule traffi.cpp:168-183 / 385-398).  Returns the slot.
UX_DEV int ring_push(const DevWorld &W, int l, int hd, int cur, double x, double x_old, double v, int vid) {
    int cap = W.l_rcap[l], tl = LQ(W, l).tail, sz = tl - hd, lanes = W.l_lanes[l];
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
    LQ(W, l).tail = tl + 1;
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
    int ord = LC(W, l).ev_cnt + 1;
    LC(W, l).ev_cnt = ord;
    W.ev_ord[TLI(W, s, l)] = ord;
    W.ev_tt[TLI(W, s, l)] = tt;
}

// Vehicle::end_trip (traffi.cpp:886-927) for the front platoon `vid` of link l, currently at position x
UX_DEV void end_trip(const DevWorld &W, int vid, int l, int t, double x) {
    LC(W, l).trips += 1;
    W.cum_dep[TLI(W, t, l)] += W.dn;
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

