/*
 * uxsim_oracle.c -- TEST INFRASTRUCTURE, NOT THE PRODUCT.
 *
 * A serial, plain-C restatement of the reference's per-timestep simulation loop
 * (toruseo/UXsim v1.14.0b8, World.exec_simulation).  It exists so that tests/ can check the CUDA
 * backend bit-for-bit; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (uxsim_b200/) never links, imports or calls it.
 *
 * It follows the C++ engine (uxsim/trafficpp/traffi.cpp) function by function -- each function below
 * cites the reference lines it restates -- using the reference's own data model (explicit
 * leader/follower links, per-link deques, per-node incoming lists, serial id-ordered node transfer), NOT
 * the GPU's ring/position layout, so that the two implementations share as little structure as possible.
 *
 * Deliberate, documented departures from the reference (the same definitions are used by the CUDA path):
 *   R1  Random numbers: Philox4x32-10 keyed on (seed, timestep, entity, purpose, draw#) (include/ux_rng.h)
 *       instead of per-node/per-link std::mt19937 (utils.h:33-36).  Stochastic parity with the reference is
 *       therefore statistical; deterministic scenarios (no draw decides anything) are bit-exact.
 *   R2  Shortest paths: per-destination backward label-correcting search on the averaged node-pair costs
 *       (dist_k[i] = min_j c_ij + dist_k[j], next = smallest j attaining it) instead of per-source forward
 *       heap Dijkstra (traffi.cpp:1317-1385).  Same costs, same tie-free next hops; equal-cost ties and
 *       last-bit rounding of path sums differ (the reference's two engines already differ there).
 *       Only destinations that appear in the demand are searched (others can never be asked for).
 *   R3  Mean platoon speed for the instantaneous travel time (traffi.cpp:608-618) is summed in the fixed
 *       order "32 interleaved partial sums, then a 5-level butterfly", which a warp reproduces exactly.
 *   R4  distance_traveled is accumulated once per link traversal (x_exit - x_entry) instead of once per
 *       step; identical up to FP rounding (not used by any decision).
 *
 * PARITY PIN: tests/test_oracle_golden.py checks this file against tests/golden/ *.npz (cumulative curves and
 * arrival steps produced by the reference's pure-Python engine, generator tests/golden/make_golden.py) and,
 * when oracle/_ref is built, against the reference's own traffi.cpp (oracle/ref_harness.cpp).
 */
#define UX_PREFIX uxo_
#include "../include/uxsim_b200.h"
#include "../include/ux_rng.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* data model                                                                                       */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    double x, y;
    double *sig; int nsig;
    double sig_offset, sig_t; int sig_phase;
    double fc, fc_remain; int lanes;
    int *in_links; int n_in;
    int *out_links; int n_out;
    int *genq; int gq_head, gq_tail, gq_cap;
    int *incoming; int n_incoming, incoming_cap;
    int *sig_log; int n_sig_log;
    double stat_count;
} ONode;

typedef struct {
    int start, end, lanes;
    double length, vmax, kappa, delta, delta_per_lane, tau, w, capacity, merge_priority;
    double cap_out, cap_out_remain, cap_in, cap_in_remain;
    int *sg; int nsg;
    double penalty; double *toll; int ntoll;
    int *q; int qh, qt, qcap;            /* vehicles on the link, front = q[qh] */
    double *arr, *dep, *tt_inst, *tt_real;
    int *ev_s; double *ev_tt; int nev, evcap, ev_applied;
    int *arrived; int n_arrived, arrived_cap;
    double ave_v_sum, stat_count, trips_completed;
} OLink;

typedef struct {
    int orig, dest, depart_ts;
    double departure_time, arrival_time, travel_time, distance, entry_x;
    int link;
    double x, x_next, x_old, v, move_remain;
    int lane, leader, follower, state, flag_wait_end, flag_abort, trip_abort, next_link;
    double arrival_time_link;
    int *pref; int npref; int *avoid; int navoid; int *route; int nroute;
    int traveled_links; unsigned char *visited;
    /* log (traffi.h:251-264) */
    int *log_link, *log_lane; double *log_x, *log_v, *log_s; int log_size, log_cap;
    int log_first_ts, log_last_ts, log_wait_count, log_end_count;
} OVeh;

struct ux_world {
    ux_world_params p;
    double delta_t; int total_ts; int route_ts; int timestep; double time;
    ONode *nodes; int nn, ncap;
    OLink *links; int nl, lcap;
    OVeh *vehs; int64_t nv, vcap;
    int initialized;
    /* route choice */
    double *pref;             /* [N][L] */
    unsigned char *pref_active; /* [N] */
    unsigned char *dest_active; /* [N] destinations present in the demand */
    double *rdist; int *rnext; /* [N][N], indexed [i*N+k] */
    int *pair_first;          /* per link: first link id of the same (start,end) pair */
    double *pair_cost;        /* per link (valid at pair_first): averaged cost */
    int *in_pair_list; int *in_pair_off; /* CSR: for node j, the pair-leading links ending at j */
    /* scratch */
    int *snapshot; int snapshot_cap;
    /* DTA solver: OD route-set store (pairs sorted by (orig, dest)) and the last route-swap result */
    int64_t dta_n_od; int *dta_o, *dta_d; int64_t *dta_od_off, *dta_route_off; int *dta_links;
    int64_t *rs_off, *ra_off; int *rs_links, *ra_links; double *ra_entry_t, *cost_actual, *t_gap; int *choice;
    double n_swap, potential_n_swap, total_t_gap; int has_swap;
    /* Edie states (last edie_state call) */
    int64_t *edie_off; int *edie_nx; double *edie_tn, *edie_dn;
    /* OD / link summary (last analysis_summary call) */
    int64_t sum_k; int *sum_o, *sum_d; double *sum_od, *sum_link;
    char err[512];
};

static char g_err[512];

static int fail(ux_world *w, int code, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(w ? w->err : g_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

static void *xrealloc(void *p, size_t n) {
    void *q = realloc(p, n ? n : 1);
    if (!q) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return q;
}
#define GROW(ptr, cap, need, type)                                   \
    do {                                                             \
        if ((need) > (cap)) {                                        \
            size_t ncap_ = (cap) ? (size_t)(cap) * 2 : 8;            \
            while (ncap_ < (size_t)(need)) ncap_ *= 2;               \
            (ptr) = (type *)xrealloc((ptr), ncap_ * sizeof(type));   \
            (cap) = (int)ncap_;                                      \
        }                                                            \
    } while (0)

static int contains_int(const int *a, int n, int v) {
    for (int i = 0; i < n; i++) if (a[i] == v) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* construction                                                                                     */
/* ------------------------------------------------------------------------------------------------ */

/* World::World traffi.cpp:1165-1216 */
UX_API ux_world *uxo_create_world(const ux_world_params *params) {
    if (!params || params->delta_n <= 0 || params->tau <= 0) { fail(NULL, UX_ERR_INVALID, "bad world params"); return NULL; }
    ux_world *w = (ux_world *)calloc(1, sizeof(ux_world));
    w->p = *params;
    w->delta_t = params->tau * params->delta_n;
    w->total_ts = (int)(params->t_max / (params->tau * params->delta_n));
    w->route_ts = (int)(params->duo_update_time / (params->tau * params->delta_n));
    if (w->route_ts == 0) w->route_ts = 1; /* traffi.cpp:1208-1212 */
    if (w->p.instantaneous_TT_timestep_interval <= 0) w->p.instantaneous_TT_timestep_interval = 5;
    return w;
}

static void free_veh(OVeh *v) {
    free(v->pref); free(v->avoid); free(v->route); free(v->visited);
    free(v->log_link); free(v->log_lane); free(v->log_x); free(v->log_v); free(v->log_s);
}

UX_API void uxo_destroy_world(ux_world *w) {
    if (!w) return;
    for (int i = 0; i < w->nn; i++) {
        ONode *n = &w->nodes[i];
        free(n->sig); free(n->in_links); free(n->out_links); free(n->genq); free(n->incoming); free(n->sig_log);
    }
    for (int i = 0; i < w->nl; i++) {
        OLink *l = &w->links[i];
        free(l->sg); free(l->toll); free(l->q); free(l->arr); free(l->dep); free(l->tt_inst); free(l->tt_real);
        free(l->ev_s); free(l->ev_tt); free(l->arrived);
    }
    for (int64_t i = 0; i < w->nv; i++) free_veh(&w->vehs[i]);
    free(w->nodes); free(w->links); free(w->vehs);
    free(w->pref); free(w->pref_active); free(w->dest_active); free(w->rdist); free(w->rnext);
    free(w->pair_first); free(w->pair_cost); free(w->in_pair_list); free(w->in_pair_off); free(w->snapshot);
    free(w->dta_o); free(w->dta_d); free(w->dta_od_off); free(w->dta_route_off); free(w->dta_links);
    free(w->rs_off); free(w->ra_off); free(w->rs_links); free(w->ra_links); free(w->ra_entry_t); free(w->cost_actual); free(w->t_gap); free(w->choice);
    free(w->edie_off); free(w->edie_nx); free(w->edie_tn); free(w->edie_dn);
    free(w->sum_o); free(w->sum_d); free(w->sum_od); free(w->sum_link);
    free(w);
}

/* Node::Node traffi.cpp:47-100 */
UX_API int uxo_add_node(ux_world *w, double x, double y, const double *sig, int nsig, double sig_offset,
                        double flow_capacity, int number_of_lanes) {
    if (w->initialized) return fail(w, UX_ERR_RUNTIME, "add_node after initialize_adj_matrix");
    GROW(w->nodes, w->ncap, w->nn + 1, ONode);
    ONode *n = &w->nodes[w->nn];
    memset(n, 0, sizeof(*n));
    n->x = x; n->y = y;
    if (nsig <= 0) { static const double zero = 0.0; sig = &zero; nsig = 1; }
    n->sig = (double *)xrealloc(NULL, sizeof(double) * nsig);
    memcpy(n->sig, sig, sizeof(double) * nsig);
    n->nsig = nsig; n->sig_offset = sig_offset;
    double cycle = 0; for (int i = 0; i < nsig; i++) cycle += sig[i];
    if (nsig > 1 || (nsig == 1 && sig[0] != 0)) { /* traffi.cpp:70-85 */
        double offset = cycle - sig_offset; int i = 0; int guard = 0;
        while (1) {
            if (offset < sig[i]) { n->sig_phase = i; n->sig_t = offset; break; }
            offset -= sig[i]; i++; if (i >= nsig) i = 0;
            if (++guard > 1000000) return fail(w, UX_ERR_INVALID, "signal offset walk does not terminate");
        }
    }
    if (flow_capacity >= 0.0) { /* traffi.cpp:91-99 */
        n->fc = flow_capacity; n->fc_remain = flow_capacity * w->p.tau * w->p.delta_n;
        n->lanes = number_of_lanes > 0 ? number_of_lanes : (int)ceil(flow_capacity / 0.8);
    } else { n->fc = -1.0; n->fc_remain = 10e10; n->lanes = number_of_lanes; }
    return w->nn++;
}

/* Link::Link traffi.cpp:465-537 */
UX_API int uxo_add_link(ux_world *w, int start, int end, double vmax, double kappa, double length, int lanes,
                        double merge_priority, double cap_out, double cap_in, const int *sg, int nsg) {
    if (w->initialized) return fail(w, UX_ERR_RUNTIME, "add_link after initialize_adj_matrix");
    if (start < 0 || start >= w->nn || end < 0 || end >= w->nn) return fail(w, UX_ERR_OUT_OF_RANGE, "add_link: node id out of range");
    if (lanes < 1) return fail(w, UX_ERR_INVALID, "add_link: number_of_lanes must be >= 1");
    GROW(w->links, w->lcap, w->nl + 1, OLink);
    OLink *l = &w->links[w->nl];
    memset(l, 0, sizeof(*l));
    l->start = start; l->end = end; l->lanes = lanes; l->length = length; l->vmax = vmax;
    l->kappa = kappa <= 0.0 ? 0.2 : kappa;
    l->merge_priority = merge_priority;
    l->delta = 1.0 / l->kappa;
    l->delta_per_lane = l->delta * lanes;
    l->tau = w->p.tau / lanes;
    l->w = 1 / l->tau / l->kappa;
    l->capacity = l->vmax * l->w * l->kappa / (l->vmax + l->w);
    l->cap_out = cap_out < 0.0 ? l->capacity * 2 : cap_out;
    l->cap_out_remain = l->cap_out * w->delta_t;
    if (cap_in < 0.0) { l->cap_in = l->capacity * 2; l->cap_in_remain = l->capacity * 2; }
    else { l->cap_in = cap_in; l->cap_in_remain = cap_in * w->delta_t; }
    if (nsg <= 0) { static const int zero = 0; sg = &zero; nsg = 1; }
    l->sg = (int *)xrealloc(NULL, sizeof(int) * nsg); memcpy(l->sg, sg, sizeof(int) * nsg); l->nsg = nsg;
    ONode *a = &w->nodes[start], *b = &w->nodes[end];
    a->out_links = (int *)xrealloc(a->out_links, sizeof(int) * (a->n_out + 1)); a->out_links[a->n_out++] = w->nl;
    b->in_links = (int *)xrealloc(b->in_links, sizeof(int) * (b->n_in + 1)); b->in_links[b->n_in++] = w->nl;
    return w->nl++;
}

/* Vehicle::Vehicle traffi.cpp:723-788 */
static int64_t new_vehicle(ux_world *w, int orig, int dest, int ts) {
    if (w->nv + 1 > w->vcap) {
        int64_t nc = w->vcap ? w->vcap * 2 : 1024;
        w->vehs = (OVeh *)xrealloc(w->vehs, sizeof(OVeh) * (size_t)nc); w->vcap = nc;
    }
    OVeh *v = &w->vehs[w->nv];
    memset(v, 0, sizeof(*v));
    v->orig = orig; v->dest = dest; v->depart_ts = ts; v->departure_time = (double)ts * w->delta_t;
    v->arrival_time = -1.0; v->travel_time = -1.0; v->link = -1; v->leader = -1; v->follower = -1;
    v->state = UX_VS_HOME; v->trip_abort = 1; v->next_link = -1;
    v->log_first_ts = -1; v->log_last_ts = -1;
    return w->nv++;
}

/* add_demand traffi.cpp:1910-1941 */
UX_API int64_t uxo_add_demand(ux_world *w, int orig, int dest, double start_t, double end_t, double flow,
                              const int *links_pref, int npref) {
    if (orig < 0 || orig >= w->nn || dest < 0 || dest >= w->nn) return fail(w, UX_ERR_OUT_OF_RANGE, "add_demand: node id out of range");
    int ts_start = (int)(start_t / w->delta_t), ts_end = (int)(end_t / w->delta_t);
    double demand = 0.0; int64_t made = 0;
    for (int ts = ts_start; ts < ts_end; ts++) {
        demand += flow * w->delta_t;
        while (demand >= (double)w->p.delta_n) {
            int64_t id = new_vehicle(w, orig, dest, ts);
            if (npref > 0) {
                OVeh *v = &w->vehs[id];
                v->pref = (int *)xrealloc(NULL, sizeof(int) * npref); memcpy(v->pref, links_pref, sizeof(int) * npref); v->npref = npref;
            }
            demand -= (double)w->p.delta_n; made++;
        }
    }
    return made;
}

UX_API int64_t uxo_add_vehicles(ux_world *w, int64_t n, const int *orig, const int *dest, const int *ts) {
    for (int64_t i = 0; i < n; i++) {
        if (orig[i] < 0 || orig[i] >= w->nn || dest[i] < 0 || dest[i] >= w->nn) return fail(w, UX_ERR_OUT_OF_RANGE, "add_vehicles: node id out of range");
        new_vehicle(w, orig[i], dest[i], ts[i]);
    }
    return n;
}

UX_API int uxo_set_vehicle_links(ux_world *w, int64_t vid, int kind, const int *links, int n) {
    if (vid < 0 || vid >= w->nv) return fail(w, UX_ERR_OUT_OF_RANGE, "vehicle index out of range");
    for (int i = 0; i < n; i++) if (links[i] < 0 || links[i] >= w->nl) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    OVeh *v = &w->vehs[vid];
    int **dst[3] = {&v->route, &v->pref, &v->avoid}; int *cnt[3] = {&v->nroute, &v->npref, &v->navoid};
    if (kind < 0 || kind > 2) return fail(w, UX_ERR_INVALID, "bad kind");
    *dst[kind] = (int *)xrealloc(*dst[kind], sizeof(int) * (n > 0 ? n : 1)); if (n > 0) memcpy(*dst[kind], links, sizeof(int) * n); *cnt[kind] = n;
    if (kind == 0) { /* enforce_route also sets links_preferred, traffi.cpp:1035-1039 */
        v->pref = (int *)xrealloc(v->pref, sizeof(int) * (n > 0 ? n : 1)); if (n > 0) memcpy(v->pref, links, sizeof(int) * n); v->npref = n;
    }
    return 0;
}

/* dta_batch_enforce_routes dta_solver.cpp:109-124 */
UX_API int uxo_batch_enforce_routes_csr(ux_world *w, const int *ids, const int64_t *off, int64_t nveh) {
    if (nveh > w->nv) nveh = w->nv;
    for (int64_t i = 0; i < nveh; i++) {
        int64_t s = off[i], e = off[i + 1];
        if (e > s) { int rc = uxo_set_vehicle_links(w, i, 0, ids + s, (int)(e - s)); if (rc) return rc; }
    }
    return 0;
}

static void link_resize_series(ux_world *w, OLink *l, int old_n, int n) {
    l->arr = (double *)xrealloc(l->arr, sizeof(double) * n);
    l->dep = (double *)xrealloc(l->dep, sizeof(double) * n);
    l->tt_real = (double *)xrealloc(l->tt_real, sizeof(double) * n);
    l->tt_inst = (double *)xrealloc(l->tt_inst, sizeof(double) * n);
    for (int i = old_n; i < n; i++) { l->arr[i] = 0; l->dep[i] = 0; l->tt_real[i] = l->length / l->vmax; l->tt_inst[i] = 0; }
    (void)w;
}

/* World::set_t_max traffi.cpp:1234-1243 */
UX_API int uxo_set_t_max(ux_world *w, double t_max) {
    int old = w->initialized ? w->total_ts : 0;
    w->p.t_max = t_max;
    w->total_ts = (int)(t_max / w->delta_t);
    if (w->initialized) {
        for (int i = 0; i < w->nl; i++) link_resize_series(w, &w->links[i], old < w->total_ts ? old : w->total_ts, w->total_ts);
        for (int i = 0; i < w->nn; i++) w->nodes[i].sig_log = (int *)xrealloc(w->nodes[i].sig_log, sizeof(int) * (w->total_ts + 1));
    }
    return 0;
}

UX_API int uxo_set_toll_timeseries(ux_world *w, int link, const double *toll, int n) {
    if (link < 0 || link >= w->nl) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    OLink *l = &w->links[link];
    l->toll = (double *)xrealloc(l->toll, sizeof(double) * (n > 0 ? n : 1)); if (n > 0) memcpy(l->toll, toll, sizeof(double) * n); l->ntoll = n;
    return 0;
}

/* Node::adjust_node_capacity traffi.cpp:212-224 */
static void adjust_node_capacity(ux_world *w, ONode *n) {
    if (n->fc < 0.0 && n->n_in > 0 && n->n_out > 0) {
        double ci = w->links[n->in_links[0]].capacity, co = w->links[n->out_links[0]].capacity;
        for (int i = 0; i < n->n_in; i++) if (w->links[n->in_links[i]].capacity > ci) ci = w->links[n->in_links[i]].capacity;
        for (int i = 0; i < n->n_out; i++) if (w->links[n->out_links[i]].capacity > co) co = w->links[n->out_links[i]].capacity;
        n->fc = (ci + co) / 2.0;
        n->fc_remain = n->fc * w->delta_t;
        if (n->lanes <= 0) n->lanes = (int)ceil(n->fc / 0.8);
    }
}

/* World::initialize_adj_matrix traffi.cpp:1245-1265 (+ static routing tables for R2) */
UX_API int uxo_initialize_adj_matrix(ux_world *w) {
    if (w->initialized) return 0;
    int N = w->nn, L = w->nl;
    if (w->p.adjust_node_capacity) for (int i = 0; i < N; i++) adjust_node_capacity(w, &w->nodes[i]);
    for (int i = 0; i < L; i++) link_resize_series(w, &w->links[i], 0, w->total_ts);
    for (int i = 0; i < N; i++) w->nodes[i].sig_log = (int *)xrealloc(NULL, sizeof(int) * (w->total_ts + 1));
    w->pref = (double *)calloc((size_t)N * (L > 0 ? L : 1), sizeof(double));
    w->pref_active = (unsigned char *)calloc(N > 0 ? N : 1, 1);
    w->dest_active = (unsigned char *)calloc(N > 0 ? N : 1, 1);
    w->rdist = (double *)xrealloc(NULL, sizeof(double) * (size_t)N * N);
    w->rnext = (int *)xrealloc(NULL, sizeof(int) * (size_t)N * N);
    for (size_t i = 0; i < (size_t)N * N; i++) { w->rdist[i] = INFINITY; w->rnext[i] = -1; }
    /* node-pair grouping: links sharing (start,end) are averaged, traffi.cpp:1297-1300 */
    w->pair_first = (int *)xrealloc(NULL, sizeof(int) * (L > 0 ? L : 1));
    w->pair_cost = (double *)xrealloc(NULL, sizeof(double) * (L > 0 ? L : 1));
    for (int i = 0; i < L; i++) {
        w->pair_first[i] = i;
        ONode *a = &w->nodes[w->links[i].start];
        for (int k = 0; k < a->n_out; k++) {
            int o = a->out_links[k];
            if (o < i && w->links[o].end == w->links[i].end) { w->pair_first[i] = w->pair_first[o]; break; }
        }
    }
    w->in_pair_off = (int *)calloc(N + 1, sizeof(int));
    for (int i = 0; i < L; i++) if (w->pair_first[i] == i) w->in_pair_off[w->links[i].end + 1]++;
    for (int i = 0; i < N; i++) w->in_pair_off[i + 1] += w->in_pair_off[i];
    w->in_pair_list = (int *)xrealloc(NULL, sizeof(int) * (L > 0 ? L : 1));
    int *fill = (int *)calloc(N > 0 ? N : 1, sizeof(int));
    for (int i = 0; i < L; i++) if (w->pair_first[i] == i) { int j = w->links[i].end; w->in_pair_list[w->in_pair_off[j] + fill[j]++] = i; }
    free(fill);
    w->initialized = 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* logging                                                                                          */
/* ------------------------------------------------------------------------------------------------ */

/* Vehicle::log_data traffi.cpp:1060-1110 */
static void log_data(ux_world *w, int vid, int snap_leader) {
    OVeh *v = &w->vehs[vid];
    if (v->state == UX_VS_RUN) {
        if (v->link >= 0) {
            OLink *l = &w->links[v->link];
            l->ave_v_sum += v->v; l->stat_count += 1.0;
        }
    } else if (v->state == UX_VS_WAIT) {
        w->nodes[v->orig].stat_count += 1.0;
    }
    if (!w->p.vehicle_log_mode) return;
    if (v->log_size == 0) v->log_first_ts = w->timestep;
    v->log_last_ts = w->timestep;
    if (v->state == UX_VS_WAIT) v->log_wait_count++;
    if (v->state == UX_VS_END || v->state == UX_VS_ABORT) v->log_end_count++;
    if (v->log_size + 1 > v->log_cap) {
        int nc = v->log_cap ? v->log_cap * 2 : 16;
        v->log_link = (int *)xrealloc(v->log_link, sizeof(int) * nc); v->log_lane = (int *)xrealloc(v->log_lane, sizeof(int) * nc);
        v->log_x = (double *)xrealloc(v->log_x, sizeof(double) * nc); v->log_v = (double *)xrealloc(v->log_v, sizeof(double) * nc);
        v->log_s = (double *)xrealloc(v->log_s, sizeof(double) * nc); v->log_cap = nc;
    }
    int i = v->log_size++;
    if (v->state == UX_VS_RUN) {
        double spacing = -1.0;
        if (snap_leader >= 0) { /* traffi.cpp:1081-1095 */
            OVeh *ld = &w->vehs[snap_leader];
            if (snap_leader < vid) { if (ld->state != UX_VS_END && ld->state != UX_VS_ABORT) spacing = ld->x - v->x; }
            else spacing = ld->x_old - v->x;
        }
        v->log_link[i] = v->link; v->log_x[i] = v->x; v->log_v[i] = v->v; v->log_lane[i] = v->lane; v->log_s[i] = spacing;
    } else {
        v->log_link[i] = -1; v->log_x[i] = -1; v->log_v[i] = -1; v->log_lane[i] = -1; v->log_s[i] = -1.0;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* route choice of one vehicle                                                                      */
/* ------------------------------------------------------------------------------------------------ */

/* random_choice utils.h:62-87 with the Philox stream (R1) */
static int weighted_pick(const double *wts, int n, double u) {
    double wsum = 0.0;
    for (int i = 0; i < n; i++) wsum += wts[i];
    if (wsum <= 0.0) { int k = (int)(u * (double)n); return k >= n ? n - 1 : k; }
    double r = u * wsum, accum = 0.0;
    for (int i = 0; i < n; i++) { accum += wts[i]; if (r <= accum) return i; }
    return n - 1;
}

/* Vehicle::route_next_link_choice traffi.cpp:934-1033.  `entity`/`purpose`/`draw` select the stream. */
static int route_next_link_choice(ux_world *w, int vid, const int *linkset, int nset, uint32_t entity, uint32_t purpose, uint32_t draw) {
    OVeh *v = &w->vehs[vid];
    if (nset == 0) { v->next_link = -1; return 0; }
    if (v->nroute > 0) {
        int traveled = v->traveled_links;
        if (traveled >= v->nroute) return fail(w, UX_ERR_OUT_OF_RANGE, "Vehicle %d: specified route has no more links (index %d out of range for route of length %d)", vid, traveled, v->nroute);
        int forced = v->route[traveled];
        if (!contains_int(linkset, nset, forced)) return fail(w, UX_ERR_INVALID, "Vehicle %d: specified route is inconsistent; link %d is not an outgoing link of the current node", vid, forced);
        v->next_link = forced; return 0;
    }
    int buf_a[64], buf_b[64]; int *out = buf_a, *tmp = buf_b; int n = 0;
    if (nset > 64) return fail(w, UX_ERR_RUNTIME, "node out-degree > 64 not supported");
    for (int i = 0; i < nset; i++) out[n++] = linkset[i];
    if (v->npref > 0) {
        int m = 0; for (int i = 0; i < n; i++) if (contains_int(v->pref, v->npref, out[i])) tmp[m++] = out[i];
        if (m > 0) { int *t = out; out = tmp; tmp = t; n = m; }
    }
    if (v->navoid > 0) {
        int m = 0; for (int i = 0; i < n; i++) if (!contains_int(v->avoid, v->navoid, out[i])) tmp[m++] = out[i];
        if (m != n) { int *t = out; out = tmp; tmp = t; n = m; }
        if (n == 0) return fail(w, UX_ERR_INVALID, "Vehicle %d: links_avoid excludes all outgoing links of the current node", vid);
    }
    if (w->p.no_cyclic_routing) {
        int m = 0;
        for (int i = 0; i < n; i++) { int e = w->links[out[i]].end; if (!v->visited || !v->visited[e]) tmp[m++] = out[i]; }
        if (m > 0) { int *t = out; out = tmp; tmp = t; n = m; }
    }
    if (n == 0) { v->next_link = -1; return 0; }
    double pw[64];
    for (int i = 0; i < n; i++) pw[i] = w->pref[(size_t)v->dest * w->nl + out[i]];
    if (w->p.hard_deterministic_mode) {
        int best = 0; for (int i = 1; i < n; i++) if (pw[i] > pw[best]) best = i;
        v->next_link = out[best];
    } else {
        double u = ux_rng_uniform(w->p.random_seed, (uint32_t)w->timestep, entity, purpose, draw);
        v->next_link = out[weighted_pick(pw, n, u)];
    }
    return 0;
}

static void mark_entered(ux_world *w, OVeh *v, OLink *ol) {
    if (w->p.no_cyclic_routing) {
        if (!v->visited) v->visited = (unsigned char *)calloc(w->nn, 1);
        v->visited[ol->start] = 1;
    }
    v->traveled_links++;
}

/* ------------------------------------------------------------------------------------------------ */
/* link queue helpers                                                                               */
/* ------------------------------------------------------------------------------------------------ */
static int q_size(const OLink *l) { return l->qt - l->qh; }
static void q_push(OLink *l, int vid) {
    if (l->qt + 1 > l->qcap) {
        if (l->qh > 0) { memmove(l->q, l->q + l->qh, sizeof(int) * (l->qt - l->qh)); l->qt -= l->qh; l->qh = 0; }
        GROW(l->q, l->qcap, l->qt + 1, int);
    }
    l->q[l->qt++] = vid;
}
static int can_accept(const ux_world *w, const OLink *ol) { /* traffi.cpp:130-142 / 285-297 */
    int ok = 0;
    if (q_size(ol) < ol->lanes) ok = 1;
    else { int idx = q_size(ol) - ol->lanes; if (w->vehs[ol->q[ol->qh + idx]].x > ol->delta_per_lane * w->p.delta_n) ok = 1; }
    if (ol->cap_in_remain < w->p.delta_n) ok = 0;
    return ok;
}
/* lane + leader assignment on entering a link, traffi.cpp:168-183 / 385-398 */
static void enter_link(ux_world *w, int vid, int lid) {
    OVeh *v = &w->vehs[vid]; OLink *ol = &w->links[lid];
    if (q_size(ol) > 0) v->lane = (w->vehs[ol->q[ol->qt - 1]].lane + 1) % ol->lanes; else v->lane = 0;
    v->leader = -1;
    if (q_size(ol) >= ol->lanes) { int ld = ol->q[ol->qh + q_size(ol) - ol->lanes]; v->leader = ld; w->vehs[ld].follower = vid; }
}

/* ------------------------------------------------------------------------------------------------ */
/* per-timestep stages                                                                              */
/* ------------------------------------------------------------------------------------------------ */

/* R3: fixed-order mean-speed sum */
static double strided_sum32(const ux_world *w, const OLink *l) {
    double part[32]; int n = q_size(l);
    for (int k = 0; k < 32; k++) part[k] = 0.0;
    for (int i = 0; i < n; i++) part[i & 31] += w->vehs[l->q[l->qh + i]].v;
    for (int off = 16; off >= 1; off >>= 1) { double nx[32]; for (int k = 0; k < 32; k++) nx[k] = part[k] + part[k ^ off]; memcpy(part, nx, sizeof(nx)); }
    return part[0];
}

/* Link::set_travel_time traffi.cpp:599-622 */
static void set_travel_time(ux_world *w, OLink *l) {
    int t = w->timestep;
    if (t % w->p.instantaneous_TT_timestep_interval != 0) { if (t > 0) l->tt_inst[t] = l->tt_inst[t - 1]; return; }
    int n = q_size(l);
    if (n > 0) {
        double avg_v = strided_sum32(w, l) / (double)n;
        l->tt_inst[t] = avg_v > 0.0 ? l->length / avg_v : l->length / (l->vmax / 100.0);
    } else l->tt_inst[t] = l->length / l->vmax;
}

/* Link::update traffi.cpp:542-565 */
static void link_update(ux_world *w, OLink *l) {
    int t = w->timestep;
    set_travel_time(w, l);
    if (t != 0) { l->arr[t] = l->arr[t - 1]; l->dep[t] = l->dep[t - 1]; }
    if (l->cap_out < 10e9) { if (l->cap_out_remain < w->p.delta_n * l->lanes) l->cap_out_remain += l->cap_out * w->delta_t; }
    else l->cap_out_remain = 10e9;
    if (l->cap_in < 10e9) { if (l->cap_in_remain < w->p.delta_n * l->lanes) l->cap_in_remain += l->cap_in * w->delta_t; }
    else l->cap_in_remain = 10e9;
}

/* Node::generate traffi.cpp:105-189 */
static int node_generate(ux_world *w, int nid) {
    ONode *nd = &w->nodes[nid];
    if (nd->gq_head == nd->gq_tail || nd->n_out == 0) return 0;
    int total_lanes = 0;
    for (int i = 0; i < nd->n_out; i++) total_lanes += w->links[nd->out_links[i]].lanes;
    for (int i = 0; i < total_lanes; i++) {
        if (nd->gq_head == nd->gq_tail) break;
        int vid = nd->genq[nd->gq_head]; OVeh *v = &w->vehs[vid];
        int rc = route_next_link_choice(w, vid, nd->out_links, nd->n_out, (uint32_t)nid, UX_RNG_GENERATE, (uint32_t)i);
        if (rc) return rc;
        if (v->next_link < 0) break;
        OLink *ol = &w->links[v->next_link];
        if (!can_accept(w, ol)) break;
        nd->gq_head++;
        v->state = UX_VS_RUN; v->link = v->next_link; v->x = 0.0; v->v = ol->vmax; v->entry_x = 0.0;
        v->arrival_time_link = (double)w->timestep * w->delta_t; /* record_travel_time(nullptr, t) */
        mark_entered(w, v, ol);
        enter_link(w, vid, v->link);
        q_push(ol, vid);
        ol->arr[w->timestep] += w->p.delta_n;
        ol->cap_in_remain -= w->p.delta_n;
    }
    return 0;
}

/* Node::signal_update traffi.cpp:194-207 */
static void node_signal_update(ux_world *w, ONode *nd) {
    if (nd->nsig > 1) {
        if (nd->sig_t > nd->sig[nd->sig_phase]) { nd->sig_phase++; nd->sig_t = 0; if (nd->sig_phase >= nd->nsig) nd->sig_phase = 0; }
        nd->sig_t += w->delta_t;
    }
    nd->sig_log[nd->n_sig_log++] = nd->sig_phase;
}

/* Node::flow_capacity_update traffi.cpp:226-234 */
static void node_flow_capacity_update(ux_world *w, ONode *nd) {
    if (nd->fc >= 0.0) { if (nd->fc_remain < w->p.delta_n * (nd->lanes > 0 ? nd->lanes : 1)) nd->fc_remain += nd->fc * w->delta_t; }
    else nd->fc_remain = 10e10;
}

static void record_tt_event(ux_world *w, OLink *l, double arrival_time_link, double tt) { /* traffi.cpp:357-362, 893-898 */
    int s = (int)(arrival_time_link / w->delta_t); if (s < 0) s = 0;
    if (l->nev + 1 > l->evcap) { int nc = l->evcap ? l->evcap * 2 : 16; l->ev_s = (int *)xrealloc(l->ev_s, sizeof(int) * nc); l->ev_tt = (double *)xrealloc(l->ev_tt, sizeof(double) * nc); l->evcap = nc; }
    l->ev_s[l->nev] = s; l->ev_tt[l->nev] = tt; l->nev++;
}

/* Vehicle::end_trip traffi.cpp:886-927 */
static void end_trip(ux_world *w, int vid) {
    OVeh *v = &w->vehs[vid]; OLink *l = &w->links[v->link];
    v->state = UX_VS_END;
    l->trips_completed += 1.0;
    l->dep[w->timestep] += w->p.delta_n;
    record_tt_event(w, l, v->arrival_time_link, ((double)w->timestep + 1.0) * w->delta_t - v->arrival_time_link);
    v->arrival_time_link = (double)w->timestep * w->delta_t;
    v->arrival_time = (double)w->timestep * w->delta_t;
    v->travel_time = v->arrival_time - v->departure_time;
    v->distance += v->x - v->entry_x; /* R4 */
    l->qh++; /* pop_front: end_trip is only ever called for the front vehicle */
    if (v->follower >= 0) w->vehs[v->follower].leader = -1;
    v->link = -1; v->x = 0.0; v->flag_wait_end = 0;
    if (v->flag_abort) { v->state = UX_VS_ABORT; v->arrival_time = -1.0; v->travel_time = -1.0; }
    log_data(w, vid, -1);
}

static void sort_ints(int *a, int n) { /* insertion sort: lists are tiny */
    for (int i = 1; i < n; i++) { int k = a[i], j = i - 1; while (j >= 0 && a[j] > k) { a[j + 1] = a[j]; j--; } a[j + 1] = k; }
}

/* Node::transfer traffi.cpp:239-445 */
static int node_transfer(ux_world *w, int nid) {
    ONode *nd = &w->nodes[nid];
    nd->n_incoming = 0;
    for (int i = 0; i < nd->n_in; i++) {
        OLink *il = &w->links[nd->in_links[i]];
        for (int k = 0; k < il->n_arrived; k++) { GROW(nd->incoming, nd->incoming_cap, nd->n_incoming + 1, int); nd->incoming[nd->n_incoming++] = il->arrived[k]; }
        il->n_arrived = 0;
    }
    if (nd->n_incoming > 1) sort_ints(nd->incoming, nd->n_incoming);
    int expanded[512]; int nexp = 0; int seen[512]; int nseen = 0;
    for (int k = 0; k < nd->n_incoming; k++) {
        int nl = w->vehs[nd->incoming[k]].next_link;
        if (nl >= 0 && !contains_int(seen, nseen, nl)) {
            if (nseen >= 512 || nexp + w->links[nl].lanes > 512) return fail(w, UX_ERR_RUNTIME, "transfer: too many out-link slots");
            seen[nseen++] = nl;
            for (int i = 0; i < w->links[nl].lanes; i++) expanded[nexp++] = nl;
        }
    }
    uint32_t draw = 0;
    if (!w->p.hard_deterministic_mode) { /* Fisher-Yates, traffi.cpp:276-282 */
        for (int i = nexp - 1; i > 0; i--) {
            int j = ux_rng_below(w->p.random_seed, (uint32_t)w->timestep, (uint32_t)nid, UX_RNG_TRANSFER, draw++, i + 1);
            int t = expanded[i]; expanded[i] = expanded[j]; expanded[j] = t;
        }
    }
    for (int s = 0; s < nexp; s++) {
        OLink *ol = &w->links[expanded[s]];
        int ok = can_accept(w, ol);
        if (nd->fc_remain < w->p.delta_n) ok = 0;
        if (!ok) continue;
        int mv[512]; double mp[512]; int nm = 0;
        for (int k = 0; k < nd->n_incoming; k++) {
            int vid = nd->incoming[k]; OVeh *v = &w->vehs[vid]; OLink *il = &w->links[v->link];
            if (v->next_link == expanded[s] && il->q[il->qh] == vid && il->cap_out_remain >= w->p.delta_n &&
                (contains_int(il->sg, il->nsg, nd->sig_phase) || nd->nsig <= 1)) { mv[nm] = vid; mp[nm] = il->merge_priority; nm++; }
        }
        if (nm == 0) continue;
        int pick;
        if (w->p.hard_deterministic_mode) { pick = 0; for (int i = 1; i < nm; i++) if (mp[i] > mp[pick]) pick = i; }
        else pick = weighted_pick(mp, nm, ux_rng_uniform(w->p.random_seed, (uint32_t)w->timestep, (uint32_t)nid, UX_RNG_TRANSFER, draw++));
        int vid = mv[pick]; OVeh *v = &w->vehs[vid]; int ilid = v->link; OLink *il = &w->links[ilid];
        il->cap_out_remain -= w->p.delta_n; ol->cap_in_remain -= w->p.delta_n;
        if (nd->fc >= 0.0) nd->fc_remain -= w->p.delta_n;
        il->dep[w->timestep] += w->p.delta_n; ol->arr[w->timestep] += w->p.delta_n;
        record_tt_event(w, il, v->arrival_time_link, (double)w->timestep * w->delta_t - v->arrival_time_link);
        v->arrival_time_link = (double)w->timestep * w->delta_t;
        v->distance += v->x - v->entry_x; /* R4 */
        il->qh++; /* pop_front */
        v->link = expanded[s]; v->x = 0.0;
        mark_entered(w, v, ol);
        if (v->follower >= 0) w->vehs[v->follower].leader = -1;
        v->follower = -1;
        enter_link(w, vid, expanded[s]);
        double x_next = v->move_remain * ol->vmax / il->vmax; /* traffi.cpp:401-419 */
        if (v->leader >= 0) {
            double x_cong = w->vehs[v->leader].x_old - ol->delta_per_lane * w->p.delta_n;
            if (x_cong < v->x) x_cong = v->x;
            if (x_next > x_cong) x_next = x_cong;
        }
        if (x_next >= ol->length) x_next = ol->length;
        if (x_next < 0.0) x_next = 0.0;
        v->x = x_next; v->entry_x = x_next;
        v->v += v->x / w->delta_t;
        v->move_remain = 0.0;
        if (q_size(il) > 0 && w->vehs[il->q[il->qh]].flag_wait_end) end_trip(w, il->q[il->qh]);
        q_push(ol, vid);
        for (int k = 0; k < nd->n_incoming; k++) if (nd->incoming[k] == vid) { memmove(nd->incoming + k, nd->incoming + k + 1, sizeof(int) * (nd->n_incoming - k - 1)); nd->n_incoming--; break; }
    }
    for (int i = 0; i < nd->n_in; i++) { /* traffi.cpp:432-441 */
        OLink *il = &w->links[nd->in_links[i]];
        for (int lane = 0; lane < il->lanes; lane++) {
            if (q_size(il) > 0 && w->vehs[il->q[il->qh]].flag_wait_end) end_trip(w, il->q[il->qh]); else break;
        }
    }
    nd->n_incoming = 0;
    return 0;
}

/* Vehicle::car_follow_newell traffi.cpp:798-823 */
static void car_follow(ux_world *w, OVeh *v) {
    OLink *l = &w->links[v->link];
    v->x_next = v->x + l->vmax * w->delta_t;
    if (v->leader >= 0) {
        double gap = w->vehs[v->leader].x - l->delta_per_lane * w->p.delta_n;
        if (gap < v->x) gap = v->x;
        if (v->x_next > gap) v->x_next = gap;
    }
    if (v->x_next < v->x) v->x_next = v->x;
    if (v->x_next > l->length) { v->move_remain = v->x_next - l->length; v->x_next = l->length; }
}

/* Vehicle::update, RUN branch, traffi.cpp:844-878 */
static int vehicle_update_run(ux_world *w, int vid, int snap_leader) {
    OVeh *v = &w->vehs[vid];
    log_data(w, vid, snap_leader);
    if (v->state != UX_VS_RUN) return 0;
    OLink *l = &w->links[v->link];
    v->v = (v->x_next - v->x) / w->delta_t;
    v->x_old = v->x;
    v->x = v->x_next;
    if (fabs(v->x - l->length) < 1e-9) {
        ONode *en = &w->nodes[l->end];
        if (l->end == v->dest) {
            v->flag_wait_end = 1;
            if (l->q[l->qh] == vid) end_trip(w, vid);
        } else if (en->n_out == 0 && v->trip_abort == 1) {
            v->flag_abort = 1; v->next_link = -1; v->flag_wait_end = 1;
            if (l->q[l->qh] == vid) end_trip(w, vid);
        } else {
            int rc = route_next_link_choice(w, vid, en->out_links, en->n_out, (uint32_t)vid, UX_RNG_VEHROUTE, 0u);
            if (rc) return rc;
            GROW(l->arrived, l->arrived_cap, l->n_arrived + 1, int);
            l->arrived[l->n_arrived++] = vid;
        }
    }
    return 0;
}

/* World::update_adj_time_matrix traffi.cpp:1267-1306 (as per-pair running averages) */
static void update_edge_costs(ux_world *w) {
    int L = w->nl, t = w->timestep;
    int *cnt = (int *)calloc(L > 0 ? L : 1, sizeof(int));
    for (int i = 0; i < L; i++) w->pair_cost[i] = 0.0;
    for (int i = 0; i < L; i++) {
        OLink *l = &w->links[i];
        double tt = l->tt_inst[t];
        if (tt <= 0.0) tt = l->length / l->vmax;
        double penalty = l->penalty;
        if (l->ntoll > 0 && t < l->ntoll) penalty = l->toll[t];
        double c;
        if (w->p.hard_deterministic_mode) c = tt + penalty;
        else { double u = ux_rng_uniform(w->p.random_seed, (uint32_t)t, (uint32_t)i, UX_RNG_LINKNOISE, 0u); c = tt * (1.0 + (w->p.route_choice_uncertainty) * u) + penalty; }
        int g = w->pair_first[i];
        double n = (double)cnt[g];
        w->pair_cost[g] = w->pair_cost[g] * n / (n + 1.0) + c / (n + 1.0);
        cnt[g] += 1;
        if (l->cap_in == 0.0) w->pair_cost[g] = INFINITY;
    }
    free(cnt);
}

/* R2: backward label-correcting search towards destination k */
static void route_search_dest(ux_world *w, int k, int *queue, unsigned char *inq) {
    int N = w->nn;
    double *dist = w->rdist; /* dist[i*N+k] */
    for (int i = 0; i < N; i++) { dist[(size_t)i * N + k] = INFINITY; w->rnext[(size_t)i * N + k] = -1; }
    dist[(size_t)k * N + k] = 0.0; w->rnext[(size_t)k * N + k] = k;
    int qh = 0, qt = 0, cnt = 0;
    memset(inq, 0, N);
    queue[qt] = k; qt = (qt + 1) % N; cnt++; inq[k] = 1;
    while (cnt > 0) {
        int j = queue[qh]; qh = (qh + 1) % N; cnt--; inq[j] = 0;
        double dj = dist[(size_t)j * N + k];
        for (int e = w->in_pair_off[j]; e < w->in_pair_off[j + 1]; e++) {
            int g = w->in_pair_list[e]; double c = w->pair_cost[g];
            if (!(c > 0.0) || isinf(c)) continue; /* adj > 0 filter traffi.cpp:1331; inf never improves */
            int i = w->links[g].start;
            double nd = c + dj;
            if (nd < dist[(size_t)i * N + k]) {
                dist[(size_t)i * N + k] = nd;
                if (!inq[i]) { queue[qt] = i; qt = (qt + 1) % N; cnt++; inq[i] = 1; }
            }
        }
    }
    /* next hop: smallest node j with c_ij + dist[j] == dist[i] (exactly; dist is the fixed point) */
    for (int g = 0; g < w->nl; g++) {
        if (w->pair_first[g] != g) continue;
        double c = w->pair_cost[g];
        if (!(c > 0.0) || isinf(c)) continue;
        int i = w->links[g].start, j = w->links[g].end;
        if (i == k) continue;
        double dj = dist[(size_t)j * N + k];
        if (isinf(dj)) continue;
        if (c + dj == dist[(size_t)i * N + k]) {
            int cur = w->rnext[(size_t)i * N + k];
            if (cur < 0 || j < cur) w->rnext[(size_t)i * N + k] = j;
        }
    }
}

/* World::route_choice_duo / _gradual traffi.cpp:1486-1553 */
static void route_choice_duo(ux_world *w, double weight) {
    int N = w->nn, L = w->nl;
    for (int k = 0; k < N; k++) {
        if (!w->dest_active[k]) continue;
        double wt = w->pref_active[k] ? weight : 1.0;
        int reinforced = 0;
        for (int l = 0; l < L; l++) {
            int i = w->links[l].start, j = w->links[l].end;
            double *p = &w->pref[(size_t)k * L + l];
            if (w->rnext[(size_t)i * N + k] == j) { *p = (1.0 - wt) * *p + wt; reinforced = 1; }
            else *p = (1.0 - wt) * *p;
        }
        if (reinforced) w->pref_active[k] = 1;
    }
}

static void route_update(ux_world *w) {
    int N = w->nn;
    update_edge_costs(w);
    int *queue = (int *)xrealloc(NULL, sizeof(int) * (N > 0 ? N : 1));
    unsigned char *inq = (unsigned char *)xrealloc(NULL, N > 0 ? N : 1);
    for (int k = 0; k < N; k++) if (w->dest_active[k]) route_search_dest(w, k, queue, inq);
    free(queue); free(inq);
}

/* World::main_loop traffi.cpp:1642-1866 */
UX_API int uxo_main_loop(ux_world *w, double duration_t, double until_t) {
    if (!w->initialized) { int rc = uxo_initialize_adj_matrix(w); if (rc) return rc; }
    int start_ts = w->timestep, end_ts;
    if (duration_t < 0 && until_t < 0) end_ts = w->total_ts;
    else if (duration_t >= 0 && until_t < 0) end_ts = (int)floor((duration_t + w->time) / w->delta_t) + 1;
    else if (duration_t < 0 && until_t >= 0) end_ts = (int)floor(until_t / w->delta_t) + 1;
    else return fail(w, UX_ERR_RUNTIME, "Cannot specify both `duration_t` and `until_t` parameters for `World.main_loop`");
    if (end_ts > w->total_ts) end_ts = w->total_ts;
    if (end_ts <= start_ts) return 0;

    /* departure buckets traffi.cpp:1663-1676 (bucket heads as a linked list through next_in_bucket) */
    int T = w->total_ts;
    int64_t *bucket_head = (int64_t *)xrealloc(NULL, sizeof(int64_t) * (T > 0 ? T : 1));
    int64_t *bucket_tail = (int64_t *)xrealloc(NULL, sizeof(int64_t) * (T > 0 ? T : 1));
    int64_t *nxt = (int64_t *)xrealloc(NULL, sizeof(int64_t) * (w->nv > 0 ? w->nv : 1));
    for (int i = 0; i < T; i++) bucket_head[i] = bucket_tail[i] = -1;
    for (int64_t i = 0; i < w->nv; i++) {
        OVeh *v = &w->vehs[i];
        w->dest_active[v->dest] = 1;
        if (v->state != UX_VS_HOME) continue;
        int ts0 = (int)(v->departure_time / w->delta_t);
        while ((double)ts0 * w->delta_t < v->departure_time) ts0++;
        if (ts0 < start_ts) ts0 = start_ts;
        if (ts0 < T) { nxt[i] = -1; if (bucket_tail[ts0] < 0) bucket_head[ts0] = i; else nxt[bucket_tail[ts0]] = i; bucket_tail[ts0] = i; }
    }
    int rc = 0;
    for (w->timestep = start_ts; w->timestep < end_ts; w->timestep++) {
        w->time = w->timestep * w->delta_t;
        for (int i = 0; i < w->nl; i++) link_update(w, &w->links[i]);
        for (int i = 0; i < w->nn; i++) {
            rc = node_generate(w, i); if (rc) goto done;
            node_signal_update(w, &w->nodes[i]);
            node_flow_capacity_update(w, &w->nodes[i]);
        }
        for (int i = 0; i < w->nn; i++) { rc = node_transfer(w, i); if (rc) goto done; }
        for (int li = 0; li < w->nl; li++) { OLink *l = &w->links[li]; for (int k = l->qh; k < l->qt; k++) car_follow(w, &w->vehs[l->q[k]]); }
        for (int li = 0; li < w->nl; li++) { /* RUN pass over a queue snapshot traffi.cpp:1763-1781 */
            OLink *l = &w->links[li];
            int n = q_size(l); if (n == 0) continue;
            GROW(w->snapshot, w->snapshot_cap, n, int);
            memcpy(w->snapshot, l->q + l->qh, sizeof(int) * n);
            for (int i = 0; i < n; i++) {
                int sl = (i - l->lanes >= 0) ? w->snapshot[i - l->lanes] : -1;
                rc = vehicle_update_run(w, w->snapshot[i], sl); if (rc) goto done;
            }
        }
        for (int i = 0; i < w->nn; i++) { ONode *nd = &w->nodes[i]; for (int k = nd->gq_head; k < nd->gq_tail; k++) log_data(w, nd->genq[k], -1); } /* WAIT pass 1789-1794 */
        if (w->timestep < T) { /* HOME departures 1804-1808 */
            for (int64_t i = bucket_head[w->timestep]; i >= 0; i = nxt[i]) {
                OVeh *v = &w->vehs[i];
                log_data(w, (int)i, -1);
                if ((double)w->timestep * w->delta_t >= v->departure_time) {
                    v->state = UX_VS_WAIT;
                    ONode *o = &w->nodes[v->orig];
                    if (o->gq_tail + 1 > o->gq_cap) {
                        if (o->gq_head > 0) { memmove(o->genq, o->genq + o->gq_head, sizeof(int) * (o->gq_tail - o->gq_head)); o->gq_tail -= o->gq_head; o->gq_head = 0; }
                        GROW(o->genq, o->gq_cap, o->gq_tail + 1, int);
                    }
                    o->genq[o->gq_tail++] = (int)i;
                }
            }
        }
        if (w->p.route_choice_update_gradual) { /* 1833-1854 */
            if (w->timestep % w->route_ts == 0) route_update(w);
            route_choice_duo(w, w->p.duo_update_weight * (w->delta_t / w->p.duo_update_time));
        } else if (w->timestep % w->route_ts == 0) { route_update(w); route_choice_duo(w, w->p.duo_update_weight); }
    }
done:
    free(bucket_head); free(bucket_tail); free(nxt);
    return rc;
}

UX_API int uxo_check_simulation_ongoing(ux_world *w) { return w->timestep < w->total_ts; }

/* ------------------------------------------------------------------------------------------------ */
/* interventions                                                                                    */
/* ------------------------------------------------------------------------------------------------ */
UX_API int uxo_set_link_param(ux_world *w, int link, int field, double value) {
    if (link < 0 || link >= w->nl) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    OLink *l = &w->links[link];
    switch (field) {
    case UX_LINK_VMAX: if (value >= 0) { l->vmax = value; l->w = 1.0 / l->tau / l->kappa; l->capacity = l->vmax * l->w * l->kappa / (l->vmax + l->w); l->delta = 1.0 / l->kappa; } break;
    case UX_LINK_KAPPA: if (value >= 0) { l->kappa = value; l->w = 1.0 / l->tau / l->kappa; l->capacity = l->vmax * l->w * l->kappa / (l->vmax + l->w); l->delta = 1.0 / l->kappa; l->delta_per_lane = l->delta * l->lanes; } break;
    case UX_LINK_CAPACITY_OUT: l->cap_out = value; break;
    case UX_LINK_CAPACITY_IN: l->cap_in = value; break;
    case UX_LINK_CAPACITY_OUT_REMAIN: l->cap_out_remain = value; break;
    case UX_LINK_CAPACITY_IN_REMAIN: l->cap_in_remain = value; break;
    case UX_LINK_MERGE_PRIORITY: l->merge_priority = value; break;
    case UX_LINK_ROUTE_CHOICE_PENALTY: l->penalty = value; break;
    default: return fail(w, UX_ERR_INVALID, "unknown link field %d", field);
    }
    return 0;
}
UX_API int uxo_set_link_signal_group(ux_world *w, int link, const int *sg, int n) {
    if (link < 0 || link >= w->nl) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    OLink *l = &w->links[link];
    l->sg = (int *)xrealloc(l->sg, sizeof(int) * (n > 0 ? n : 1)); if (n > 0) memcpy(l->sg, sg, sizeof(int) * n); l->nsg = n;
    return 0;
}
UX_API int uxo_set_node_signal(ux_world *w, int node, const double *sig, int nsig, int phase, double sig_t) {
    if (node < 0 || node >= w->nn) return fail(w, UX_ERR_OUT_OF_RANGE, "node id out of range");
    ONode *n = &w->nodes[node];
    if (sig && nsig > 0) {
        n->sig = (double *)xrealloc(n->sig, sizeof(double) * nsig); memcpy(n->sig, sig, sizeof(double) * nsig); n->nsig = nsig;
        if (n->sig_phase >= nsig) { n->sig_phase = 0; n->sig_t = 0; } /* uxsim_cpp_wrapper.py:96-99 */
    }
    if (phase >= 0) n->sig_phase = phase;
    if (sig_t >= 0) n->sig_t = sig_t;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* results                                                                                          */
/* ------------------------------------------------------------------------------------------------ */
UX_API int64_t uxo_get_int(ux_world *w, int what) {
    switch (what) {
    case UX_I_NUM_NODES: return w->nn;
    case UX_I_NUM_LINKS: return w->nl;
    case UX_I_NUM_VEHICLES: return w->nv;
    case UX_I_TIMESTEP: return w->timestep;
    case UX_I_TOTAL_TIMESTEPS: return w->total_ts;
    case UX_I_NUM_REPLICAS: return 1;
    case UX_I_ROUTE_UPDATE_STEPS: return w->route_ts;
    case UX_I_PLATOON_UPDATES: { double s = 0; for (int i = 0; i < w->nl; i++) s += w->links[i].stat_count; return (int64_t)s; }
    case UX_I_TRIPS_COMPLETED: { double s = 0; for (int i = 0; i < w->nl; i++) s += w->links[i].trips_completed; return (int64_t)s; }
    case UX_I_LOG_ENTRIES: { int64_t s = 0; for (int64_t i = 0; i < w->nv; i++) s += w->vehs[i].log_size; return s; }
    case UX_I_KERNEL_LAUNCHES: return 0;
    case UX_I_NUM_ACTIVE_DESTS: { int s = 0; if (w->dest_active) for (int i = 0; i < w->nn; i++) s += w->dest_active[i]; return s; }
    case UX_I_TRANSFER_LEVELS: return 1;
    case UX_I_LINK_EVENTS: { int64_t s = 0; for (int i = 0; i < w->nl; i++) s += w->links[i].nev; return s; }
    case UX_I_MAX_DEPART_TS: { int m = 0; for (int64_t i = 0; i < w->nv; i++) if (w->vehs[i].depart_ts > m) m = w->vehs[i].depart_ts; return m; }
    default: return fail(w, UX_ERR_INVALID, "unknown int query %d", what);
    }
}
UX_API double uxo_get_double(ux_world *w, int what) {
    switch (what) {
    case UX_D_DELTA_T: return w->delta_t;
    case UX_D_TIME: return w->timestep * w->delta_t;
    case UX_D_T_MAX: return w->p.t_max;
    case UX_D_AVE_V_SUM: { double s = 0; for (int i = 0; i < w->nl; i++) s += w->links[i].ave_v_sum; return s; }
    case UX_D_DTA_N_SWAP: return w->n_swap;
    case UX_D_DTA_POTENTIAL_N_SWAP: return w->potential_n_swap;
    case UX_D_DTA_TOTAL_T_GAP: return w->total_t_gap;
    case UX_D_STAT_SAMPLE_COUNT: { double s = 0; for (int i = 0; i < w->nl; i++) s += w->links[i].stat_count; for (int i = 0; i < w->nn; i++) s += w->nodes[i].stat_count; return s; }
    default: return NAN;
    }
}

/* Link::ensure_traveltime_real traffi.cpp:631-648 */
static void ensure_traveltime_real(ux_world *w, OLink *l) {
    int n = w->total_ts;
    for (int k = l->ev_applied; k < l->nev; k++) {
        int s = l->ev_s[k], e = n;
        if (k + 1 < l->nev) { int sn = l->ev_s[k + 1]; if (sn < n) e = sn; }
        for (int i = s; i < e; i++) l->tt_real[i] = l->ev_tt[k];
    }
    l->ev_applied = l->nev;
}

/* Vehicle::log_t_at / log_state_at traffi.cpp:1117-1133 */
static double log_t_at(const ux_world *w, const OVeh *v, int i) {
    int tail = v->log_size - v->log_end_count;
    long long ts = i < tail ? (long long)v->log_first_ts + i : (long long)v->log_last_ts;
    return (double)ts * w->delta_t;
}
static int log_state_at(const OVeh *v, int i) {
    if (v->log_end_count > 0 && i >= v->log_size - v->log_end_count) return v->state;
    if (i == 0) return UX_VS_HOME;
    if (i < 1 + v->log_wait_count) return UX_VS_WAIT;
    return UX_VS_RUN;
}
static int ltl_count(const OVeh *v) {
    int c = 1, prev = -999;
    for (int i = 0; i < v->log_size; i++) { int lid = v->log_link[i]; if (lid >= 0 && lid != prev) { c++; prev = lid; } }
    if (v->state == UX_VS_END) c++;
    return c;
}
static int enter_count(const OVeh *v) {
    int c = 0, prev = -999;
    for (int i = 0; i < v->log_size; i++) { int lid = v->log_link[i]; if (lid >= 0 && lid != prev) { c++; prev = lid; } }
    return c;
}

static int veh_effective_state(const OVeh *v) { /* traffi.cpp:1947-1958 */
    int s = v->state;
    if (s == UX_VS_END && (v->flag_abort || (v->arrival_time < 0 && v->travel_time <= 0))) s = UX_VS_ABORT;
    return s;
}

UX_API int64_t uxo_array_size(ux_world *w, int what, int *dtype) {
    int64_t N = w->nn, L = w->nl, T = w->total_ts, P = w->nv; int dt = UX_DT_F64; int64_t n = -1;
    switch (what) {
    case UX_A_CUM_ARRIVAL: case UX_A_CUM_DEPARTURE: case UX_A_TRAVELTIME_INSTANT: case UX_A_TRAVELTIME_ACTUAL: n = L * T; break;
    case UX_A_VEH_STATE: case UX_A_VEH_ARRIVAL_TS: case UX_A_VEH_DEPART_TS: case UX_A_VEH_ORIG: case UX_A_VEH_DEST: case UX_A_VEH_LINK: case UX_A_VEH_LANE: n = P; dt = UX_DT_I32; break;
    case UX_A_VEH_TRAVEL_TIME: case UX_A_VEH_DISTANCE: case UX_A_VEH_X: case UX_A_VEH_V: case UX_A_VEH_LINK_ARRIVAL_TIME: n = P; break;
    case UX_A_SIGNAL_LOG: n = N * T; dt = UX_DT_I32; break;
    case UX_A_NODE_SIGNAL_PHASE: n = N; dt = UX_DT_I32; break;
    case UX_A_NODE_SIGNAL_T: case UX_A_NODE_FLOW_CAPACITY: n = N; break;
    case UX_A_NODE_NUM_LANES: n = N; dt = UX_DT_I32; break;
    case UX_A_LINK_NUM_VEHICLES: n = L; dt = UX_DT_I32; break;
    case UX_A_LINK_CAPACITY_IN_REMAIN: case UX_A_LINK_CAPACITY_OUT_REMAIN: case UX_A_LINK_STAT_COUNT: case UX_A_LINK_TRIPS_COMPLETED: case UX_A_LINK_AVE_V_SUM: n = L; break;
    case UX_A_ROUTE_PREF: n = N * L; break;
    case UX_A_ROUTE_DIST: n = N * N; break;
    case UX_A_ROUTE_NEXT: n = N * N; dt = UX_DT_I32; break;
    case UX_A_LOG_OFFSETS: case UX_A_LTL_OFFSETS: n = P + 1; dt = UX_DT_I64; break;
    case UX_A_LOG_T: case UX_A_LOG_X: case UX_A_LOG_V: case UX_A_LOG_S: n = uxo_get_int(w, UX_I_LOG_ENTRIES); break;
    case UX_A_LOG_STATE: case UX_A_LOG_LANE: case UX_A_LOG_LINK: n = uxo_get_int(w, UX_I_LOG_ENTRIES); dt = UX_DT_I32; break;
    case UX_A_LOG_N_MISSING: n = P; dt = UX_DT_I32; break;
    case UX_A_LTL_T: case UX_A_LTL_ID: { n = 0; for (int64_t i = 0; i < P; i++) n += ltl_count(&w->vehs[i]); dt = what == UX_A_LTL_T ? UX_DT_F64 : UX_DT_I32; break; }
    case UX_A_ENTER_LINK: case UX_A_ENTER_VEH: case UX_A_ENTER_TIME: { n = 0; for (int64_t i = 0; i < P; i++) n += enter_count(&w->vehs[i]); dt = what == UX_A_ENTER_TIME ? UX_DT_F64 : UX_DT_I32; break; }
    case UX_A_EDIE_OFFSETS: n = w->edie_off ? L + 1 : 0; dt = UX_DT_I64; break;
    case UX_A_EDIE_NX: n = w->edie_off ? L : 0; dt = UX_DT_I32; break;
    case UX_A_EDIE_TN: case UX_A_EDIE_DN: n = w->edie_off ? w->edie_off[L] : 0; break;
    case UX_A_OD_ORIG: case UX_A_OD_DEST: n = w->sum_od ? w->sum_k : 0; dt = UX_DT_I32; break;
    case UX_A_OD_STATS: n = w->sum_od ? w->sum_k * 6 : 0; break;
    case UX_A_LINKC_STATS: n = w->sum_link ? L * 4 : 0; break;
    case UX_A_DTA_OD_ORIG: case UX_A_DTA_OD_DEST: n = w->dta_n_od; dt = UX_DT_I32; break;
    case UX_A_DTA_OD_OFFSETS: n = w->dta_n_od + 1; dt = UX_DT_I64; break;
    case UX_A_DTA_ROUTE_OFFSETS: n = (w->dta_od_off ? w->dta_od_off[w->dta_n_od] : 0) + 1; dt = UX_DT_I64; break;
    case UX_A_DTA_ROUTE_LINKS: n = w->dta_od_off ? w->dta_route_off[w->dta_od_off[w->dta_n_od]] : 0; dt = UX_DT_I32; break;
    case UX_A_DTA_RS_OFFSETS: case UX_A_DTA_RA_OFFSETS: n = P + 1; dt = UX_DT_I64; break;
    case UX_A_DTA_RS_LINKS: n = w->has_swap ? w->rs_off[P] : 0; dt = UX_DT_I32; break;
    case UX_A_DTA_RA_LINKS: n = w->has_swap ? w->ra_off[P] : 0; dt = UX_DT_I32; break;
    case UX_A_DTA_RA_ENTRY_T: n = w->has_swap ? w->ra_off[P] : 0; break;
    case UX_A_DTA_COST_ACTUAL: case UX_A_DTA_T_GAP: n = P; break;
    case UX_A_DTA_CHOICE: n = P; dt = UX_DT_I32; break;
    default: return fail(w, UX_ERR_INVALID, "unknown array %d", what);
    }
    if (dtype) *dtype = dt;
    return n;
}

UX_API int64_t uxo_get_array(ux_world *w, int what, int replica, void *dst, int64_t capacity) {
    int dt; int64_t n = uxo_array_size(w, what, &dt);
    if (n < 0) return n;
    if (replica != 0) return fail(w, UX_ERR_OUT_OF_RANGE, "oracle has a single replica");
    if (capacity < n) return fail(w, UX_ERR_INVALID, "get_array: capacity %lld < %lld", (long long)capacity, (long long)n);
    if (!w->initialized) { int rc = uxo_initialize_adj_matrix(w); if (rc) return rc; }
    double *d = (double *)dst; int *ii = (int *)dst; int64_t *ll = (int64_t *)dst;
    int64_t N = w->nn, L = w->nl, T = w->total_ts, P = w->nv;
    switch (what) {
    case UX_A_CUM_ARRIVAL: for (int64_t l = 0; l < L; l++) memcpy(d + l * T, w->links[l].arr, sizeof(double) * T); break;
    case UX_A_CUM_DEPARTURE: for (int64_t l = 0; l < L; l++) memcpy(d + l * T, w->links[l].dep, sizeof(double) * T); break;
    case UX_A_TRAVELTIME_INSTANT: for (int64_t l = 0; l < L; l++) memcpy(d + l * T, w->links[l].tt_inst, sizeof(double) * T); break;
    case UX_A_TRAVELTIME_ACTUAL: for (int64_t l = 0; l < L; l++) { ensure_traveltime_real(w, &w->links[l]); memcpy(d + l * T, w->links[l].tt_real, sizeof(double) * T); } break;
    case UX_A_VEH_STATE: for (int64_t i = 0; i < P; i++) ii[i] = veh_effective_state(&w->vehs[i]); break;
    case UX_A_VEH_ARRIVAL_TS: for (int64_t i = 0; i < P; i++) { const OVeh *v = &w->vehs[i]; ii[i] = (v->arrival_time >= 0 && veh_effective_state(v) != UX_VS_ABORT) ? (int)(v->arrival_time / w->delta_t) : -1; } break;
    case UX_A_VEH_DEPART_TS: for (int64_t i = 0; i < P; i++) ii[i] = (int)(w->vehs[i].departure_time / w->delta_t); break;
    case UX_A_VEH_TRAVEL_TIME: for (int64_t i = 0; i < P; i++) d[i] = veh_effective_state(&w->vehs[i]) == UX_VS_ABORT ? -1.0 : w->vehs[i].travel_time; break;
    case UX_A_VEH_LINK_ARRIVAL_TIME: for (int64_t i = 0; i < P; i++) d[i] = w->vehs[i].arrival_time_link; break;
    case UX_A_VEH_DISTANCE: for (int64_t i = 0; i < P; i++) { const OVeh *v = &w->vehs[i]; d[i] = v->distance + (v->state == UX_VS_RUN ? v->x - v->entry_x : 0.0); } break;
    case UX_A_VEH_ORIG: for (int64_t i = 0; i < P; i++) ii[i] = w->vehs[i].orig; break;
    case UX_A_VEH_DEST: for (int64_t i = 0; i < P; i++) ii[i] = w->vehs[i].dest; break;
    case UX_A_VEH_LINK: for (int64_t i = 0; i < P; i++) ii[i] = w->vehs[i].link; break;
    case UX_A_VEH_X: for (int64_t i = 0; i < P; i++) d[i] = w->vehs[i].x; break;
    case UX_A_VEH_V: for (int64_t i = 0; i < P; i++) d[i] = w->vehs[i].v; break;
    case UX_A_VEH_LANE: for (int64_t i = 0; i < P; i++) ii[i] = w->vehs[i].lane; break;
    case UX_A_SIGNAL_LOG: for (int64_t k = 0; k < N; k++) for (int64_t t = 0; t < T; t++) ii[k * T + t] = t < w->nodes[k].n_sig_log ? w->nodes[k].sig_log[t] : 0; break;
    case UX_A_NODE_SIGNAL_PHASE: for (int64_t k = 0; k < N; k++) ii[k] = w->nodes[k].sig_phase; break;
    case UX_A_NODE_SIGNAL_T: for (int64_t k = 0; k < N; k++) d[k] = w->nodes[k].sig_t; break;
    case UX_A_NODE_FLOW_CAPACITY: for (int64_t k = 0; k < N; k++) d[k] = w->nodes[k].fc; break;
    case UX_A_NODE_NUM_LANES: for (int64_t k = 0; k < N; k++) ii[k] = w->nodes[k].lanes; break;
    case UX_A_LINK_NUM_VEHICLES: for (int64_t l = 0; l < L; l++) ii[l] = q_size(&w->links[l]); break;
    case UX_A_LINK_CAPACITY_IN_REMAIN: for (int64_t l = 0; l < L; l++) d[l] = w->links[l].cap_in_remain; break;
    case UX_A_LINK_CAPACITY_OUT_REMAIN: for (int64_t l = 0; l < L; l++) d[l] = w->links[l].cap_out_remain; break;
    case UX_A_LINK_STAT_COUNT: for (int64_t l = 0; l < L; l++) d[l] = w->links[l].stat_count; break;
    case UX_A_LINK_TRIPS_COMPLETED: for (int64_t l = 0; l < L; l++) d[l] = w->links[l].trips_completed; break;
    case UX_A_LINK_AVE_V_SUM: for (int64_t l = 0; l < L; l++) d[l] = w->links[l].ave_v_sum; break;
    case UX_A_ROUTE_PREF: memcpy(d, w->pref, sizeof(double) * N * L); break;
    case UX_A_ROUTE_DIST: case UX_A_ROUTE_NEXT: {
        /* RouteChoice.dist / .next are all-pairs tables in the reference (traffi.cpp:1317-1385); the loop only needs -- and only
           searches -- the destinations that occur in the demand.  The getter completes the other columns on the edge costs of the
           last route update (the loop never reads them). */
        if (w->timestep > 0 && N > 0) {
            int *queue = (int *)xrealloc(NULL, sizeof(int) * N);
            unsigned char *inq = (unsigned char *)xrealloc(NULL, N);
            for (int k = 0; k < N; k++) if (!w->dest_active[k]) route_search_dest(w, k, queue, inq);
            free(queue); free(inq);
        }
        if (what == UX_A_ROUTE_DIST) memcpy(d, w->rdist, sizeof(double) * N * N); else memcpy(ii, w->rnext, sizeof(int) * N * N);
        break; }
    case UX_A_LOG_OFFSETS: { int64_t s = 0; for (int64_t i = 0; i < P; i++) { ll[i] = s; s += w->vehs[i].log_size; } ll[P] = s; break; }
    case UX_A_LOG_N_MISSING: for (int64_t i = 0; i < P; i++) { const OVeh *v = &w->vehs[i]; ii[i] = (v->log_size > 0 && log_t_at(w, v, 0) > 0) ? (int)(log_t_at(w, v, 0) / w->delta_t) : 0; } break;
    case UX_A_LOG_T: case UX_A_LOG_X: case UX_A_LOG_V: case UX_A_LOG_S: case UX_A_LOG_STATE: case UX_A_LOG_LANE: case UX_A_LOG_LINK: {
        int64_t s = 0;
        for (int64_t i = 0; i < P; i++) {
            const OVeh *v = &w->vehs[i];
            for (int k = 0; k < v->log_size; k++, s++) {
                int st = log_state_at(v, k); int run = st == UX_VS_RUN;
                switch (what) {
                case UX_A_LOG_T: d[s] = log_t_at(w, v, k); break;
                case UX_A_LOG_X: d[s] = run ? v->log_x[k] : -1; break;
                case UX_A_LOG_V: d[s] = run ? v->log_v[k] : -1; break;
                case UX_A_LOG_S: d[s] = v->log_s[k]; break;
                case UX_A_LOG_STATE: ii[s] = st; break;
                case UX_A_LOG_LANE: ii[s] = v->log_lane[k]; break;
                default: ii[s] = v->log_link[k]; break;
                }
            }
        }
        break; }
    case UX_A_LTL_OFFSETS: { int64_t s = 0; for (int64_t i = 0; i < P; i++) { ll[i] = s; s += ltl_count(&w->vehs[i]); } ll[P] = s; break; }
    case UX_A_LTL_T: case UX_A_LTL_ID: { /* traffi.cpp:2056-2077 */
        int64_t s = 0;
        for (int64_t i = 0; i < P; i++) {
            const OVeh *v = &w->vehs[i];
            if (what == UX_A_LTL_T) d[s] = v->departure_time; else ii[s] = -1;
            s++;
            int prev = -999;
            for (int k = 0; k < v->log_size; k++) { int lid = v->log_link[k]; if (lid >= 0 && lid != prev) { if (what == UX_A_LTL_T) d[s] = log_t_at(w, v, k); else ii[s] = lid; s++; prev = lid; } }
            if (v->state == UX_VS_END) { if (what == UX_A_LTL_T) d[s] = v->arrival_time >= 0 ? v->arrival_time : -1.0; else ii[s] = -2; s++; }
        }
        break; }
    case UX_A_ENTER_LINK: case UX_A_ENTER_TIME: case UX_A_ENTER_VEH: { /* traffi.cpp:1965-1983 */
        int64_t s = 0;
        for (int64_t i = 0; i < P; i++) {
            const OVeh *v = &w->vehs[i]; int prev = -999;
            for (int k = 0; k < v->log_size; k++) { int lid = v->log_link[k]; if (lid >= 0 && lid != prev) { if (what == UX_A_ENTER_LINK) ii[s] = lid; else if (what == UX_A_ENTER_TIME) d[s] = log_t_at(w, v, k); else ii[s] = (int)i; s++; prev = lid; } }
        }
        break; }
    case UX_A_EDIE_OFFSETS: memcpy(dst, w->edie_off, sizeof(int64_t) * n); break;
    case UX_A_EDIE_NX: memcpy(dst, w->edie_nx, sizeof(int) * n); break;
    case UX_A_EDIE_TN: memcpy(dst, w->edie_tn, sizeof(double) * n); break;
    case UX_A_EDIE_DN: memcpy(dst, w->edie_dn, sizeof(double) * n); break;
    case UX_A_OD_ORIG: memcpy(dst, w->sum_o, sizeof(int) * n); break;
    case UX_A_OD_DEST: memcpy(dst, w->sum_d, sizeof(int) * n); break;
    case UX_A_OD_STATS: memcpy(dst, w->sum_od, sizeof(double) * n); break;
    case UX_A_LINKC_STATS: memcpy(dst, w->sum_link, sizeof(double) * n); break;
    case UX_A_DTA_OD_ORIG: memcpy(dst, w->dta_o, sizeof(int) * n); break;
    case UX_A_DTA_OD_DEST: memcpy(dst, w->dta_d, sizeof(int) * n); break;
    case UX_A_DTA_OD_OFFSETS: if (w->dta_od_off) memcpy(dst, w->dta_od_off, sizeof(int64_t) * n); else ((int64_t *)dst)[0] = 0; break;
    case UX_A_DTA_ROUTE_OFFSETS: if (w->dta_route_off) memcpy(dst, w->dta_route_off, sizeof(int64_t) * n); else ((int64_t *)dst)[0] = 0; break;
    case UX_A_DTA_ROUTE_LINKS: memcpy(dst, w->dta_links, sizeof(int) * n); break;
    case UX_A_DTA_RS_OFFSETS: case UX_A_DTA_RA_OFFSETS: case UX_A_DTA_RS_LINKS: case UX_A_DTA_RA_LINKS: case UX_A_DTA_RA_ENTRY_T:
    case UX_A_DTA_COST_ACTUAL: case UX_A_DTA_T_GAP: case UX_A_DTA_CHOICE: {
        if (!w->has_swap) return fail(w, UX_ERR_RUNTIME, "no route swap result");
        const void *src = what == UX_A_DTA_RS_OFFSETS ? (void *)w->rs_off : what == UX_A_DTA_RA_OFFSETS ? (void *)w->ra_off : what == UX_A_DTA_RS_LINKS ? (void *)w->rs_links
            : what == UX_A_DTA_RA_LINKS ? (void *)w->ra_links : what == UX_A_DTA_RA_ENTRY_T ? (void *)w->ra_entry_t : what == UX_A_DTA_COST_ACTUAL ? (void *)w->cost_actual
            : what == UX_A_DTA_T_GAP ? (void *)w->t_gap : (void *)w->choice;
        memcpy(dst, src, (dt == UX_DT_I32 ? 4 : 8) * (size_t)n);
        break; }
    default: return fail(w, UX_ERR_INVALID, "unknown array %d", what);
    }
    return n;
}

UX_API int64_t uxo_get_device_array(ux_world *w, int what, int replica, void **dptr, int *dtype) {
    (void)what; (void)replica; (void)dptr; (void)dtype;
    return fail(w, UX_ERR_RUNTIME, "the oracle has no device arrays");
}

/* per-link ensemble statistics: {volume, sum tt_actual, sum tt_actual^2, platoons on link} */
UX_API int64_t uxo_ensemble_link_stats(ux_world *w, double *dst, int64_t capacity) {
    int64_t L = w->nl, T = w->total_ts;
    if (capacity < L * 4) return fail(w, UX_ERR_INVALID, "capacity too small");
    for (int64_t l = 0; l < L; l++) {
        OLink *k = &w->links[l]; ensure_traveltime_real(w, k);
        double s1 = 0, s2 = 0; for (int64_t t = 0; t < T; t++) { s1 += k->tt_real[t]; s2 += k->tt_real[t] * k->tt_real[t]; }
        dst[l * 4 + 0] = T > 0 ? k->arr[T - 1] : 0.0; dst[l * 4 + 1] = s1; dst[l * 4 + 2] = s2; dst[l * 4 + 3] = (double)q_size(k);
    }
    return L * 4;
}
UX_API int64_t uxo_ensemble_link_stats_device(ux_world *w, void **dptr) { (void)dptr; return fail(w, UX_ERR_RUNTIME, "the oracle has no device arrays"); }


/* ------------------------------------------------------------------------------------------------ */
/* DTA solver batch operations: restatement of dta_solver.cpp + traffi.cpp:650-689,  */
/* 1317-1481.  Random streams are the reference's own here (std::mt19937 seeded per call), so these   */
/* functions are bit-exact against oracle/_ref for every output.                                      */
/* ------------------------------------------------------------------------------------------------ */

/* std::mt19937 = MT19937 (Matsumoto & Nishimura 1998), default parameters */
typedef struct { uint32_t s[624]; int idx; } mt19937;
static void mt_seed(mt19937 *m, uint32_t seed) {
    m->s[0] = seed;
    for (int i = 1; i < 624; i++) m->s[i] = 1812433253u * (m->s[i - 1] ^ (m->s[i - 1] >> 30)) + (uint32_t)i;
    m->idx = 624;
}
static uint32_t mt_next(mt19937 *m) {
    if (m->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (m->s[i] & 0x80000000u) | (m->s[(i + 1) % 624] & 0x7fffffffu);
            m->s[i] = m->s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->idx = 0;
    }
    uint32_t y = m->s[m->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
/* std::uniform_real_distribution<double>(a, b) over mt19937 as libstdc++ evaluates it: generate_canonical<double, 53>
 * takes two 32-bit draws, sum = d0 + d1 * 2^32, divided by 2^64 (clamped below 1), then * (b - a) + a. */
static double mt_uniform(mt19937 *m, double a, double b) {
    double sum = 0.0, tmp = 1.0;
    for (int k = 0; k < 2; k++) { sum += (double)mt_next(m) * tmp; tmp *= 4294967296.0; }
    double r = sum / tmp;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r * (b - a) + a;
}

/* min-heap of (dist, node) pairs ordered like std::pair<double,int> under std::greater (traffi.cpp:1353) */
typedef struct { double d; int n; } HeapItem;
static int heap_less(HeapItem a, HeapItem b) { return a.d < b.d || (a.d == b.d && a.n < b.n); }
static void heap_push(HeapItem **h, int *n, int *cap, HeapItem it) {
    if (*n >= *cap) { *cap = *cap ? *cap * 2 : 64; *h = (HeapItem *)xrealloc(*h, sizeof(HeapItem) * (size_t)*cap); }
    int i = (*n)++;
    while (i > 0) { int p = (i - 1) / 2; if (!heap_less(it, (*h)[p])) break; (*h)[i] = (*h)[p]; i = p; }
    (*h)[i] = it;
}
static HeapItem heap_pop(HeapItem *h, int *n) {
    HeapItem top = h[0], last = h[--(*n)];
    int i = 0;
    while (1) {
        int c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && heap_less(h[c + 1], h[c])) c++;
        if (!heap_less(h[c], last)) break;
        h[i] = h[c]; i = c;
    }
    if (*n > 0) h[i] = last;
    return top;
}

/* World::route_search_all(adj, 0) traffi.cpp:1317-1385: forward Dijkstra from every source over the cells adj > 0,
 * neighbours in ascending node order, strict improvement, next hop inherited from the parent. */
static void dta_route_search_all(int N, const double *adj, double *dist, int *next) {
    int *nb_off = (int *)xrealloc(NULL, sizeof(int) * (size_t)(N + 1)), cnt = 0;
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) if (adj[(size_t)i * N + j] > 0.0) cnt++;
    int *nb = (int *)xrealloc(NULL, sizeof(int) * (size_t)(cnt + 1));
    cnt = 0;
    for (int i = 0; i < N; i++) { nb_off[i] = cnt; for (int j = 0; j < N; j++) if (adj[(size_t)i * N + j] > 0.0) nb[cnt++] = j; }
    nb_off[N] = cnt;
    unsigned char *visited = (unsigned char *)xrealloc(NULL, (size_t)N + 1);
    HeapItem *heap = NULL; int hn = 0, hcap = 0;
    for (int start = 0; start < N; start++) {
        double *ds = dist + (size_t)start * N; int *nh = next + (size_t)start * N;
        for (int i = 0; i < N; i++) { ds[i] = INFINITY; nh[i] = -1; visited[i] = 0; }
        ds[start] = 0.0; nh[start] = start;
        HeapItem it0 = {0.0, start}; heap_push(&heap, &hn, &hcap, it0);
        while (hn > 0) {
            HeapItem it = heap_pop(heap, &hn);
            int cur = it.n;
            if (visited[cur]) continue;
            visited[cur] = 1;
            double dcur = ds[cur];
            for (int q = nb_off[cur]; q < nb_off[cur + 1]; q++) {
                int nx = nb[q];
                double nd = dcur + adj[(size_t)cur * N + nx];
                if (nd < ds[nx]) { ds[nx] = nd; nh[nx] = (cur == start) ? nx : nh[cur]; HeapItem pi = {nd, nx}; heap_push(&heap, &hn, &hcap, pi); }
            }
        }
    }
    free(nb_off); free(nb); free(visited); free(heap);
}

typedef struct { int n; int **r; int *len; } RouteList;

static void dta_store_free(ux_world *w) {
    free(w->dta_o); free(w->dta_d); free(w->dta_od_off); free(w->dta_route_off); free(w->dta_links);
    w->dta_o = w->dta_d = w->dta_links = NULL; w->dta_od_off = w->dta_route_off = NULL; w->dta_n_od = 0;
}

/* dta_enumerate_od_route_sets_ids dta_solver.cpp:16-24 = World::enumerate_k_random_routes_cpp traffi.cpp:1387-1481 */
UX_API int64_t uxo_dta_enumerate_route_sets(ux_world *w, int k, unsigned int seed) {
    if (!w->initialized) { int rc = uxo_initialize_adj_matrix(w); if (rc) return rc; }
    int N = w->nn, L = w->nl;
    int *pair_link = (int *)xrealloc(NULL, sizeof(int) * (size_t)N * N + 4);
    for (size_t i = 0; i < (size_t)N * N; i++) pair_link[i] = -1;
    for (int l = 0; l < L; l++) pair_link[(size_t)w->links[l].start * N + w->links[l].end] = l;  /* last link wins */
    RouteList *rl = (RouteList *)calloc((size_t)N * N + 1, sizeof(RouteList));
    unsigned char *present = (unsigned char *)calloc((size_t)N * N + 1, 1);
    double *adj = (double *)calloc((size_t)N * N + 1, sizeof(double)), *dist = (double *)xrealloc(NULL, sizeof(double) * (size_t)N * N + 8);
    int *next = (int *)xrealloc(NULL, sizeof(int) * (size_t)N * N + 4), *path = (int *)xrealloc(NULL, sizeof(int) * (size_t)(N + 1));
    mt19937 rng; mt_seed(&rng, seed);
    int counter = 0, iteration = 0; double avg_old = 0.0;
    while (1) {
        for (int l = 0; l < L; l++) {
            const OLink *ln = &w->links[l];
            double base = ln->length / ln->vmax;
            adj[(size_t)ln->start * N + ln->end] = iteration == 0 ? base : base * mt_uniform(&rng, 0.0, 100.0);
        }
        dta_route_search_all(N, adj, dist, next);
        long long total_routes = 0, total_pairs = 0;
        for (int src = 0; src < N; src++) for (int dst = 0; dst < N; dst++) {
            if (src == dst || next[(size_t)src * N + dst] == -1) continue;
            RouteList *R = &rl[(size_t)src * N + dst];
            present[(size_t)src * N + dst] = 1;
            total_pairs++; total_routes += R->n;
            if (R->n >= k) continue;
            int len = 0, cur = src, valid = 1;
            while (cur != dst) {
                int nx = next[(size_t)cur * N + dst];
                if (nx == -1 || nx == cur) { valid = 0; break; }
                int lid = pair_link[(size_t)cur * N + nx];
                if (lid == -1 || len >= N) { valid = 0; break; }
                path[len++] = lid; cur = nx;
            }
            if (!valid || len == 0) continue;
            int dup = 0;
            for (int q = 0; q < R->n && !dup; q++) if (R->len[q] == len && memcmp(R->r[q], path, sizeof(int) * (size_t)len) == 0) dup = 1;
            if (dup) continue;
            R->r = (int **)xrealloc(R->r, sizeof(int *) * (size_t)(R->n + 1)); R->len = (int *)xrealloc(R->len, sizeof(int) * (size_t)(R->n + 1));
            R->r[R->n] = (int *)xrealloc(NULL, sizeof(int) * (size_t)len); memcpy(R->r[R->n], path, sizeof(int) * (size_t)len); R->len[R->n] = len; R->n++;
        }
        double avg = total_pairs > 0 ? (double)total_routes / (double)total_pairs : 0.0;
        if (avg == avg_old) counter++; else { counter = 0; avg_old = avg; }
        iteration++;
        if (counter > 10) break;
    }
    /* flatten: pairs in (orig, dest) order, including connected pairs that ended up with no route */
    dta_store_free(w);
    int64_t n_od = 0, n_routes = 0, n_links = 0;
    for (size_t i = 0; i < (size_t)N * N; i++) if (present[i]) { n_od++; n_routes += rl[i].n; for (int q = 0; q < rl[i].n; q++) n_links += rl[i].len[q]; }
    w->dta_n_od = n_od;
    w->dta_o = (int *)xrealloc(NULL, sizeof(int) * (size_t)n_od); w->dta_d = (int *)xrealloc(NULL, sizeof(int) * (size_t)n_od);
    w->dta_od_off = (int64_t *)xrealloc(NULL, sizeof(int64_t) * (size_t)(n_od + 1)); w->dta_route_off = (int64_t *)xrealloc(NULL, sizeof(int64_t) * (size_t)(n_routes + 1));
    w->dta_links = (int *)xrealloc(NULL, sizeof(int) * (size_t)n_links);
    int64_t p = 0, r = 0, e = 0;
    w->dta_od_off[0] = 0; w->dta_route_off[0] = 0;
    for (int src = 0; src < N; src++) for (int dst = 0; dst < N; dst++) {
        size_t i = (size_t)src * N + dst;
        if (!present[i]) continue;
        w->dta_o[p] = src; w->dta_d[p] = dst;
        for (int q = 0; q < rl[i].n; q++) { memcpy(w->dta_links + e, rl[i].r[q], sizeof(int) * (size_t)rl[i].len[q]); e += rl[i].len[q]; w->dta_route_off[++r] = e; free(rl[i].r[q]); }
        free(rl[i].r); free(rl[i].len);
        w->dta_od_off[++p] = r;
    }
    free(pair_link); free(rl); free(present); free(adj); free(dist); free(next); free(path);
    return n_routes;
}

/* World.make_od_route_set_store bindings.cpp:467-489 */
UX_API int uxo_dta_set_route_sets(ux_world *w, int64_t n_od, const int *orig, const int *dest, const int64_t *od_off, const int64_t *route_off, const int *links) {
    for (int64_t p = 1; p < n_od; p++)
        if (orig[p] < orig[p - 1] || (orig[p] == orig[p - 1] && dest[p] <= dest[p - 1])) return fail(w, UX_ERR_INVALID, "route sets must be sorted by (orig, dest) without duplicates");
    dta_store_free(w);
    int64_t n_routes = od_off[n_od], n_links = route_off[n_routes];
    w->dta_n_od = n_od;
    w->dta_o = (int *)xrealloc(NULL, sizeof(int) * (size_t)n_od); memcpy(w->dta_o, orig, sizeof(int) * (size_t)n_od);
    w->dta_d = (int *)xrealloc(NULL, sizeof(int) * (size_t)n_od); memcpy(w->dta_d, dest, sizeof(int) * (size_t)n_od);
    w->dta_od_off = (int64_t *)xrealloc(NULL, sizeof(int64_t) * (size_t)(n_od + 1)); memcpy(w->dta_od_off, od_off, sizeof(int64_t) * (size_t)(n_od + 1));
    w->dta_route_off = (int64_t *)xrealloc(NULL, sizeof(int64_t) * (size_t)(n_routes + 1)); memcpy(w->dta_route_off, route_off, sizeof(int64_t) * (size_t)(n_routes + 1));
    w->dta_links = (int *)xrealloc(NULL, sizeof(int) * (size_t)n_links); memcpy(w->dta_links, links, sizeof(int) * (size_t)n_links);
    return 0;
}

UX_API int uxo_dta_share_route_sets(ux_world *w, ux_world *from) {
    if (!from) return fail(w, UX_ERR_INVALID, "dta_share_route_sets: no source world");
    if (from == w) return 0;
    if (!from->dta_od_off) { dta_store_free(w); return 0; }
    return uxo_dta_set_route_sets(w, from->dta_n_od, from->dta_o, from->dta_d, from->dta_od_off, from->dta_route_off, from->dta_links);
}
UX_API int uxo_dta_next_day(ux_world *w) { return fail(w, UX_ERR_RUNTIME, "the oracle builds a new world per day (as the reference does)"); }
UX_API int uxo_dta_enforce_routes(ux_world *w, int replica, const int *ids, const int64_t *off, int64_t nveh) {
    if (replica != 0) return fail(w, UX_ERR_OUT_OF_RANGE, "oracle has a single replica");
    return uxo_batch_enforce_routes_csr(w, ids, off, nveh);
}

/* dta_get_toll / dta_get_actual_travel_time dta_solver.h:44-65 */
static double dta_toll(const ux_world *w, const OLink *ln, double t) {
    if (ln->ntoll > 0) { int ts = (int)(t / w->delta_t); if (ts < 0) ts = 0; if (ts >= ln->ntoll) ts = ln->ntoll - 1; return ln->toll[ts]; }
    return ln->penalty;
}
static double dta_actual_tt(ux_world *w, OLink *ln, double t) {
    ensure_traveltime_real(w, ln);
    int n = w->total_ts;
    if (n == 0) return ln->length / ln->vmax;
    int idx = (int)(t / w->delta_t);
    if (idx < 0) idx = 0;
    if (idx >= n) idx = n - 1;
    return ln->tt_real[idx];
}
/* Link::estimate_congestion_externality traffi.cpp:650-689 */
static double dta_externality(ux_world *w, OLink *ln, int ts) {
    ensure_traveltime_real(w, ln);
    int n = w->total_ts; double dt = w->delta_t, fftt = ln->length / ln->vmax * 1.01;
    int ts_end = ts;
    for (int i = ts; i < n; i++) {
        double tt_i = ln->tt_real[i];
        int fut = i + (int)(tt_i / dt);
        if (fut >= n) break;
        if (tt_i > fftt && ln->arr[fut] - ln->dep[fut] > 0) ts_end = i; else break;
    }
    double veh_count = 0;
    if (ts >= 0 && ts < n && ts_end >= 0 && ts_end < n) veh_count = ln->arr[ts_end] - ln->arr[ts];
    int ts_out = 0;
    if (ts >= 0 && ts < n) ts_out = (int)(ts + (int)(ln->tt_real[ts] / dt)) + 1;
    if (ts_out >= n) ts_out = n - 1;
    int ts_out_leader = ts_out;
    for (int i = ts_out; i > 0; i--) if (i < n && ts_out < n && ln->dep[i] < ln->dep[ts_out]) { ts_out_leader = i; break; }
    return (double)(ts_out - ts_out_leader) * dt * veh_count;
}
/* dta_compute_route_cost_due / _dso dta_solver.cpp:72-104 */
static double dta_route_cost(ux_world *w, int mode, const int *links, int n, double departure_time) {
    double tt_total = 0.0, extra = 0.0, t = departure_time;
    for (int j = 0; j < n; j++) {
        OLink *ln = &w->links[links[j]];
        double ltt = dta_actual_tt(w, ln, t);
        if (mode == 0) extra += dta_toll(w, ln, t); else extra += dta_externality(w, ln, (int)(t / w->delta_t));
        tt_total += ltt; t += ltt;
    }
    return tt_total + extra;
}

/* dta_route_swap_due / dta_route_swap_dso dta_solver.cpp:130-373 */
UX_API int uxo_dta_route_swap(ux_world *w, int replica, int mode, double swap_prob, int swap_num, int has_external, unsigned int rng_seed) {
    if (replica != 0) return fail(w, UX_ERR_OUT_OF_RANGE, "oracle has a single replica");
    if (mode != 0 && swap_num >= 0) return fail(w, UX_ERR_INVALID, "swap_num >= 0 uses std::shuffle, which this restatement does not reproduce");
    int64_t P = w->nv;
    free(w->rs_off); free(w->ra_off); free(w->rs_links); free(w->ra_links); free(w->ra_entry_t); free(w->cost_actual); free(w->t_gap); free(w->choice);
    w->rs_off = (int64_t *)calloc((size_t)P + 1, sizeof(int64_t)); w->ra_off = (int64_t *)calloc((size_t)P + 1, sizeof(int64_t));
    w->cost_actual = (double *)calloc((size_t)P + 1, sizeof(double)); w->t_gap = (double *)calloc((size_t)P + 1, sizeof(double)); w->choice = (int *)calloc((size_t)P + 1, sizeof(int));
    /* traveled routes (dta_get_traveled_route dta_solver.cpp:30-57) */
    int64_t tot = 0;
    for (int64_t i = 0; i < P; i++) { w->ra_off[i] = tot; tot += enter_count(&w->vehs[i]); }
    w->ra_off[P] = tot;
    w->ra_links = (int *)xrealloc(NULL, sizeof(int) * (size_t)tot + 4); w->ra_entry_t = (double *)xrealloc(NULL, sizeof(double) * (size_t)tot + 8);
    for (int64_t i = 0; i < P; i++) {
        const OVeh *v = &w->vehs[i]; int prev = -999; int64_t s = w->ra_off[i];
        for (int k = 0; k < v->log_size; k++) { int lid = v->log_link[k]; if (lid >= 0 && lid != prev) { w->ra_links[s] = lid; w->ra_entry_t[s] = log_t_at(w, v, k); s++; prev = lid; } }
    }
    mt19937 rng; mt_seed(&rng, rng_seed);
    unsigned char *cand = NULL;
    if (mode != 0) { cand = (unsigned char *)calloc((size_t)P + 1, 1); for (int64_t i = 0; i < P; i++) cand[i] = mt_uniform(&rng, 0.0, 1.0) < swap_prob; }
    w->n_swap = 0; w->potential_n_swap = 0; w->total_t_gap = 0;
    for (int64_t vi = 0; vi < P; vi++) {
        const OVeh *v = &w->vehs[vi];
        const int *tr = w->ra_links + w->ra_off[vi]; const double *et = w->ra_entry_t + w->ra_off[vi]; int ntr = (int)(w->ra_off[vi + 1] - w->ra_off[vi]);
        double arrival = (v->state == UX_VS_END && v->arrival_time >= 0) ? v->arrival_time : -1.0;
        w->cost_actual[vi] = ntr > 0 ? arrival - et[0] : 0;
        w->choice[vi] = -2;
        if (v->state != UX_VS_END) continue;
        w->choice[vi] = -1;
        int64_t lo = 0, hi = w->dta_n_od, pair = -1;
        while (lo < hi) { int64_t mid = (lo + hi) / 2; if (w->dta_o[mid] < v->orig || (w->dta_o[mid] == v->orig && w->dta_d[mid] < v->dest)) lo = mid + 1; else hi = mid; }
        if (lo < w->dta_n_od && w->dta_o[lo] == v->orig && w->dta_d[lo] == v->dest) pair = lo;
        if (pair < 0) continue;
        int64_t r0 = w->dta_od_off[pair], r1 = w->dta_od_off[pair + 1];
        if (has_external) {
            int found = 0;
            for (int64_t r = r0; r < r1 && !found; r++) { int64_t len = w->dta_route_off[r + 1] - w->dta_route_off[r]; if (len == ntr && memcmp(w->dta_links + w->dta_route_off[r], tr, sizeof(int) * (size_t)ntr) == 0) found = 1; }
            if (!found && r1 > r0) { w->choice[vi] = (int)r0; continue; }
        }
        int flag_swap = mode == 0 ? (mt_uniform(&rng, 0.0, 1.0) < swap_prob) : cand[vi];
        double departure_t = ntr > 0 ? et[0] : v->departure_time, cost_current = 0.0;
        if (mode == 0) {
            double t = departure_t;
            for (int j = 0; j < ntr; j++) { OLink *ln = &w->links[tr[j]]; double ltt = dta_actual_tt(w, ln, t); cost_current += dta_toll(w, ln, et[j]); cost_current += ltt; t += ltt; }
        } else cost_current = dta_route_cost(w, 1, tr, ntr, departure_t);
        int changed = 0, changed_idx = -1; double t_gap = 0.0, pot = 0.0;
        for (int64_t r = r0; r < r1; r++) {
            double cost_alt = dta_route_cost(w, mode, w->dta_links + w->dta_route_off[r], (int)(w->dta_route_off[r + 1] - w->dta_route_off[r]), departure_t);
            if (cost_alt < cost_current) { t_gap = cost_current - cost_alt; pot = w->p.delta_n; if (flag_swap) { changed = 1; changed_idx = (int)r; cost_current = cost_alt; } }
        }
        w->potential_n_swap += pot; w->total_t_gap += t_gap; w->t_gap[vi] = t_gap;
        if (changed && changed_idx >= 0) { w->choice[vi] = changed_idx; w->n_swap += w->p.delta_n; }
    }
    free(cand);
    /* routes_specified */
    tot = 0;
    for (int64_t vi = 0; vi < P; vi++) {
        w->rs_off[vi] = tot;
        int c = w->choice[vi];
        tot += c == -2 ? 0 : c == -1 ? w->ra_off[vi + 1] - w->ra_off[vi] : w->dta_route_off[c + 1] - w->dta_route_off[c];
    }
    w->rs_off[P] = tot;
    w->rs_links = (int *)xrealloc(NULL, sizeof(int) * (size_t)tot + 4);
    for (int64_t vi = 0; vi < P; vi++) {
        int c = w->choice[vi]; int64_t n = w->rs_off[vi + 1] - w->rs_off[vi];
        if (n > 0) memcpy(w->rs_links + w->rs_off[vi], c == -1 ? w->ra_links + w->ra_off[vi] : w->dta_links + w->dta_route_off[c], sizeof(int) * (size_t)n);
    }
    w->has_swap = 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Edie states: restatement of Analyzer.compute_accurate_traj + compute_edie_state  */
/* (analyzer.py:205-330).                                                                            */
/* ------------------------------------------------------------------------------------------------ */
#define EDIE_FX 16777216.0 /* 2^24: contributions are rounded to this fixed point (include/uxsim_b200.h) */
static long long edie_fx(double v) { return (long long)floor(v * EDIE_FX + 0.5); }
/* Python's float floor division and modulo for a positive divisor (floatobject.c float_divmod) */
static double py_mod(double x, double y) { double m = fmod(x, y); if (m != 0.0 && ((y < 0) != (m < 0))) m += y; return m; }
static double py_floordiv(double x, double y) {
    double m = fmod(x, y), div = (x - m) / y;
    if (m != 0.0 && ((y < 0) != (m < 0))) div -= 1.0;
    if (div == 0.0) return 0.0;
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return fl;
}
/* one trajectory segment (x0,t0) -> (x1,t1) on a link whose grid is nt x nx cells of dt x dx (analyzer.py:272-312) */
static void edie_segment(long long *tn, long long *dn, int nt, int nx, double dt, double dx, double x0, double t0, double x1, double t1) {
    double v0 = (t1 - t0 != 0) ? (x1 - x0) / (t1 - t0) : 0.0;
    int tt = (int)py_floordiv(t0, dt), xx = (int)py_floordiv(x0, dx);
    if (tt < 0) tt += nt;  /* Python's negative index: an extrapolated start before t = 0 lands in the last rows */
    if (tt < 0 || !(tt < nt && xx < nx)) return;
    if (v0 > 0) {
        double xl0 = dx - py_mod(x0, dx), xl1 = py_mod(x1, dx), tl0 = xl0 / v0, tl1 = xl1 / v0;
        double x1c = py_floordiv(x1, dx);
        if ((double)xx == x1c) { dn[(size_t)tt * nx + xx] += edie_fx(x1 - x0); tn[(size_t)tt * nx + xx] += edie_fx(t1 - t0); }
        else {
            int jj = (int)(x1c - xx + 1);
            for (int j = 0; j < jj; j++) {
                if (xx + j >= nx) continue;
                size_t c = (size_t)tt * nx + xx + j;
                if (j == 0) { dn[c] += edie_fx(xl0); tn[c] += edie_fx(tl0); }
                else if (j == jj - 1) { dn[c] += edie_fx(xl1); tn[c] += edie_fx(tl1); }
                else { dn[c] += edie_fx(dx); tn[c] += edie_fx(dx / v0); }
            }
        }
    } else tn[(size_t)tt * nx + xx] += edie_fx(t1 - t0);
}

UX_API int uxo_edie_state(ux_world *w, int replica, const double *edie_dt, const double *edie_dx) {
    if (replica != 0) return fail(w, UX_ERR_OUT_OF_RANGE, "oracle has a single replica");
    if (!w->p.vehicle_log_mode) return fail(w, UX_ERR_RUNTIME, "edie_state needs the per-step vehicle logs (vehicle_log_mode)");
    int L = w->nl;
    free(w->edie_off); free(w->edie_nx); free(w->edie_tn); free(w->edie_dn);
    w->edie_off = (int64_t *)calloc((size_t)L + 1, sizeof(int64_t)); w->edie_nx = (int *)calloc((size_t)L + 1, sizeof(int));
    for (int l = 0; l < L; l++) {
        if (!(edie_dt[l] > 0) || !(edie_dx[l] > 0)) return fail(w, UX_ERR_INVALID, "edie_state: cell sizes must be positive");
        int nt = (int)(w->p.t_max / edie_dt[l]), nx = (int)(w->links[l].length / edie_dx[l]);
        w->edie_nx[l] = nx; w->edie_off[l + 1] = w->edie_off[l] + (int64_t)nt * nx;
    }
    int64_t cells = w->edie_off[L];
    long long *tn = (long long *)calloc((size_t)cells + 1, sizeof(long long)), *dn = (long long *)calloc((size_t)cells + 1, sizeof(long long));
    for (int64_t vi = 0; vi < w->nv; vi++) {
        const OVeh *v = &w->vehs[vi];
        int i = 0;
        while (i < v->log_size) {  /* a maximal run of records on one link = one trajectory on that link (analyzer.py:217-231) */
            int l = v->log_link[i];
            if (l < 0) { i++; continue; }
            int j = i;
            while (j + 1 < v->log_size && v->log_link[j + 1] == l) j++;
            const OLink *ln = &w->links[l];
            int nx = w->edie_nx[l], nt = nx > 0 ? (int)((w->edie_off[l + 1] - w->edie_off[l]) / nx) : 0;
            long long *ltn = tn + w->edie_off[l], *ldn = dn + w->edie_off[l];
            double dt = edie_dt[l], dx = edie_dx[l], u = ln->vmax;
            if (nx > 0 && nt > 0) {
                double xf = v->log_x[i], tf = log_t_at(w, v, i);
                if (xf != 0 && xf / u > w->delta_t * 0.01) edie_segment(ltn, ldn, nt, nx, dt, dx, 0.0, tf - xf / u, xf, tf);  /* extrapolated entry */
                for (int k = i; k < j; k++) edie_segment(ltn, ldn, nt, nx, dt, dx, v->log_x[k], log_t_at(w, v, k), v->log_x[k + 1], log_t_at(w, v, k + 1));
                double xe = v->log_x[j], te = log_t_at(w, v, j);
                if (ln->length - u * w->delta_t <= xe && xe < ln->length) {  /* extrapolated exit */
                    double rem = ln->length - xe;
                    if (rem / u > w->delta_t * 0.01) edie_segment(ltn, ldn, nt, nx, dt, dx, xe, te, ln->length, te + rem / u);
                }
            }
            i = j + 1;
        }
    }
    w->edie_tn = (double *)calloc((size_t)cells + 1, sizeof(double)); w->edie_dn = (double *)calloc((size_t)cells + 1, sizeof(double));
    for (int64_t c = 0; c < cells; c++) { w->edie_tn[c] = (double)tn[c] / EDIE_FX * w->p.delta_n; w->edie_dn[c] = (double)dn[c] / EDIE_FX * w->p.delta_n; }
    free(tn); free(dn);
    return 0;
}

/* ------------------------------------------------------------------------------------------------- */
/* Analyzer.od_analysis (analyzer.py:121-181) and Analyzer.link_analysis_coarse (183-203), restated:     */
/* per OD pair the lists of travel times (completed trips) and of distances travelled (all vehicles),    */
/* np.average / np.std over them; per link volume, remaining vehicles, mean / std of the positive        */
/* entries of traveltime_actual.  Pairs are reported ascending in (orig * N + dest).                     */
/* ------------------------------------------------------------------------------------------------- */
static int cmp_pair_key(const void *a, const void *b) {
    const int64_t *x = (const int64_t *)a, *y = (const int64_t *)b;  /* {key, vehicle id} */
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : x[1] > y[1];
}
UX_API int uxo_analysis_summary(ux_world *w, int replica) {
    if (replica != 0) return fail(w, UX_ERR_OUT_OF_RANGE, "replica %d out of range", replica);
    if (w->timestep == 0) return fail(w, UX_ERR_RUNTIME, "analysis_summary needs a run");
    int64_t P = w->nv; int L = w->nl, T = w->total_ts, N = w->nn;
    free(w->sum_o); free(w->sum_d); free(w->sum_od); free(w->sum_link);
    int64_t *kv = (int64_t *)calloc((size_t)P * 2 + 2, sizeof(int64_t));
    for (int64_t i = 0; i < P; i++) { kv[2 * i] = (int64_t)w->vehs[i].orig * N + w->vehs[i].dest; kv[2 * i + 1] = i; }
    qsort(kv, (size_t)P, 2 * sizeof(int64_t), cmp_pair_key);
    int64_t K = 0;
    for (int64_t q = 0; q < P; q++) if (q == 0 || kv[2 * q] != kv[2 * q - 2]) K++;
    w->sum_k = K;
    w->sum_o = (int *)calloc((size_t)K + 1, sizeof(int)); w->sum_d = (int *)calloc((size_t)K + 1, sizeof(int));
    w->sum_od = (double *)calloc((size_t)K * 6 + 1, sizeof(double)); w->sum_link = (double *)calloc((size_t)L * 4 + 1, sizeof(double));
    int64_t k = -1;
    for (int64_t q = 0; q < P;) {
        int64_t e = q; while (e < P && kv[2 * e] == kv[2 * q]) e++;
        k++;
        w->sum_o[k] = (int)(kv[2 * q] / N); w->sum_d[k] = (int)(kv[2 * q] % N);
        double n = 0, nc = 0, stt = 0, sd = 0;
        for (int64_t j = q; j < e; j++) {  /* od_trips, od_dist, od_trips_comp, od_tt (analyzer.py:164-176) */
            const OVeh *v = &w->vehs[kv[2 * j + 1]];
            double d = v->distance + (v->state == UX_VS_RUN ? v->x - v->entry_x : 0.0);
            n += 1; sd += d;
            if (v->travel_time != -1.0) { nc += 1; stt += v->travel_time; }
        }
        double md = sd / n, mt = nc > 0 ? stt / nc : 0.0, vd = 0, vt = 0;
        for (int64_t j = q; j < e; j++) {  /* np.std: population standard deviation */
            const OVeh *v = &w->vehs[kv[2 * j + 1]];
            double d = v->distance + (v->state == UX_VS_RUN ? v->x - v->entry_x : 0.0);
            vd += (d - md) * (d - md);
            if (v->travel_time != -1.0) vt += (v->travel_time - mt) * (v->travel_time - mt);
        }
        double *o = w->sum_od + k * 6;
        o[0] = n; o[1] = nc; o[2] = nc > 0 ? mt : -1.0; o[3] = nc > 0 ? sqrt(vt / nc) : -1.0; o[4] = md; o[5] = sqrt(vd / n);
        q = e;
    }
    free(kv);
    for (int l = 0; l < L; l++) {  /* link_analysis_coarse (analyzer.py:183-203) */
        OLink *lk = &w->links[l];
        double vol = lk->dep[T - 1], mean = -1.0, sdev = -1.0;
        if (vol != 0.0) {
            ensure_traveltime_real(w, lk);
            double s = 0, n = 0;
            for (int i = 0; i < T; i++) if (lk->tt_real[i] > 0.0) { s += lk->tt_real[i]; n += 1; }
            if (n > 0) {
                mean = s / n;
                double qq = 0;
                for (int i = 0; i < T; i++) if (lk->tt_real[i] > 0.0) qq += (lk->tt_real[i] - mean) * (lk->tt_real[i] - mean);
                sdev = sqrt(qq / n);
            }
        }
        double *o = w->sum_link + (size_t)l * 4;
        o[0] = vol; o[1] = lk->arr[T - 1] - vol; o[2] = mean; o[3] = sdev;
    }
    return 0;
}

#define PFX(n) uxo_##n
/* vectorised ingest: add_node / add_link / add_demand in index order (include/uxsim_b200.h) */
UX_API int64_t PFX(add_nodes)(ux_world *w, int64_t n, const double *x, const double *y, const int *so, const double *sig,
                              const double *sig_offset, const double *fc, const int *lanes) {
    for (int64_t i = 0; i < n; i++) { int rc = PFX(add_node)(w, x[i], y[i], sig + so[i], so[i + 1] - so[i], sig_offset[i], fc[i], lanes[i]); if (rc < 0) return rc; }
    return n;
}
UX_API int64_t PFX(add_links)(ux_world *w, int64_t n, const int *a, const int *b, const double *vmax, const double *kappa, const double *len,
                              const int *lanes, const double *mp, const double *co, const double *ci, const int *go, const int *sg) {
    for (int64_t i = 0; i < n; i++) { int rc = PFX(add_link)(w, a[i], b[i], vmax[i], kappa[i], len[i], lanes[i], mp[i], co[i], ci[i], sg + go[i], go[i + 1] - go[i]); if (rc < 0) return rc; }
    return n;
}
UX_API int64_t PFX(add_demands)(ux_world *w, int64_t n, const int *o, const int *d, const double *t0, const double *t1, const double *flow) {
    int64_t made = 0;
    for (int64_t i = 0; i < n; i++) { int64_t m = PFX(add_demand)(w, o[i], d[i], t0[i], t1[i], flow[i], 0, 0); if (m < 0) return m; made += m; }
    return made;
}
UX_API int PFX(set_option)(ux_world *w, int option, double value) { (void)w; (void)option; (void)value; return 0; }

UX_API const char *uxo_last_error(ux_world *w) { return w ? w->err : g_err; }
