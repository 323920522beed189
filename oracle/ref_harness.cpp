// ref_harness.cpp -- TEST INFRASTRUCTURE, NOT THE PRODUCT.
//
// Exposes the reference's OWN C++ engine (toruseo/UXsim uxsim/trafficpp/traffi.cpp, compiled from the
// sources where they lie under /root/reference -- nothing is copied into this repo) behind the same C-ABI
// as the product (include/uxsim_b200.h, prefix uxr_).  It plays the role of the reference's nanobind
// module (bindings.cpp, which cannot be built here: nanobind is absent) for two purposes only:
//   * pinning the oracle (oracle/uxsim_oracle.c) and the CUDA backend against the real reference engine;
//   * the CPU baseline of bench.py (`cpu_baseline.kind == "reference"`, `--impl reference`).
// Build: `make -C oracle ref` -> oracle/_ref/libuxsim_ref.so (git-ignored; travels to the GPU box).
//
// Every function is a thin forwarding call; the cited bindings.cpp lines show the equivalent nanobind glue.
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include <exception>
#include <stdexcept>

#include "traffi.cpp"      // resolved via -I/root/reference/uxsim/trafficpp (as bindings.cpp:18 does)
#include "dta_solver.cpp"  // bindings.cpp:19

#define UX_PREFIX uxr_
#include "../include/uxsim_b200.h"

struct ux_world {
    World *W = nullptr;
    ux_world_params p;
    std::string err;
    bool initialized = false;
    // DTA solver state (bindings.cpp keeps these as Python-owned objects)
    ODRouteSetStore store;
    std::vector<std::pair<int, int>> od_keys;  // sorted (orig, dest)
    RouteSwapResult swap;
    std::vector<double> swap_entry_t;          // TraveledRouteInfo::entry_times, flat in route_actual order
    bool has_swap = false;
};

static std::string g_err;

static int fail(ux_world *w, int code, const std::string &msg) {
    (w ? w->err : g_err) = msg;
    return code;
}

#define UXR_TRY try {
#define UXR_CATCH(w)                                                                                  \
    }                                                                                                 \
    catch (const std::out_of_range &e) { return fail(w, UX_ERR_OUT_OF_RANGE, e.what()); }             \
    catch (const std::invalid_argument &e) { return fail(w, UX_ERR_INVALID, e.what()); }              \
    catch (const std::exception &e) { return fail(w, UX_ERR_RUNTIME, e.what()); }

extern "C" {

// create_world bindings.cpp:150-181
UX_API ux_world *uxr_create_world(const ux_world_params *p) {
    try {
        auto *h = new ux_world();
        h->p = *p;
        h->W = new World("", p->t_max, p->delta_n, p->tau, p->duo_update_time, p->duo_update_weight,
                         p->route_choice_uncertainty, 0 /*print_mode*/, (long long)p->random_seed,
                         p->vehicle_log_mode != 0, p->hard_deterministic_mode != 0,
                         p->route_choice_update_gradual != 0, p->no_cyclic_routing != 0);
        // uxsim_cpp_wrapper.py:894-898
        h->W->adjust_node_capacity = p->adjust_node_capacity != 0;
        if (p->instantaneous_TT_timestep_interval > 0)
            h->W->instantaneous_TT_timestep_interval = p->instantaneous_TT_timestep_interval;
        h->W->num_threads = p->num_threads != 0 ? p->num_threads : 1;
        h->W->show_progress = 0;
        return h;
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

UX_API void uxr_destroy_world(ux_world *w) {
    if (!w) return;
    delete w->W;
    delete w;
}

// add_node bindings.cpp:186-190
UX_API int uxr_add_node(ux_world *w, double x, double y, const double *sig, int nsig, double sig_offset,
                        double flow_capacity, int number_of_lanes) {
    UXR_TRY
    std::vector<double> s = nsig > 0 ? std::vector<double>(sig, sig + nsig) : std::vector<double>{0};
    int id = w->W->node_id;
    new Node(w->W, std::to_string(id), x, y, s, sig_offset, flow_capacity, number_of_lanes);
    return id;
    UXR_CATCH(w)
}

// add_link bindings.cpp:192-207
UX_API int uxr_add_link(ux_world *w, int start, int end, double vmax, double kappa, double length, int lanes,
                        double merge_priority, double cap_out, double cap_in, const int *sg, int nsg) {
    UXR_TRY
    if (start < 0 || start >= w->W->node_id || end < 0 || end >= w->W->node_id)
        return fail(w, UX_ERR_OUT_OF_RANGE, "add_link: node id out of range");
    std::vector<int> g = nsg > 0 ? std::vector<int>(sg, sg + nsg) : std::vector<int>{0};
    int id = w->W->link_id;
    new Link(w->W, std::to_string(id), std::to_string(start), std::to_string(end), vmax, kappa, length, lanes,
             merge_priority, cap_out, cap_in, g);
    return id;
    UXR_CATCH(w)
}

// add_demand bindings.cpp:209-216 -> traffi.cpp:1910-1941
UX_API int64_t uxr_add_demand(ux_world *w, int orig, int dest, double start_t, double end_t, double flow,
                              const int *links_pref, int npref) {
    UXR_TRY
    if (orig < 0 || orig >= w->W->node_id || dest < 0 || dest >= w->W->node_id)
        return fail(w, UX_ERR_OUT_OF_RANGE, "add_demand: node id out of range");
    std::vector<std::string> pref;
    for (int i = 0; i < npref; i++) pref.push_back(std::to_string(links_pref[i]));
    int before = w->W->vehicle_id;
    add_demand(w->W, std::to_string(orig), std::to_string(dest), start_t, end_t, flow, pref);
    return w->W->vehicle_id - before;
    UXR_CATCH(w)
}

// CppWorld.addVehicle uxsim_cpp_wrapper.py:959-961: one add_demand per platoon
UX_API int64_t uxr_add_vehicles(ux_world *w, int64_t n, const int *orig, const int *dest, const int *ts) {
    UXR_TRY
    double dt = w->W->delta_t;
    for (int64_t i = 0; i < n; i++) {
        double t = (double)ts[i] * dt;
        int before = w->W->vehicle_id;
        add_demand(w->W, std::to_string(orig[i]), std::to_string(dest[i]), t, t + dt, w->W->delta_n / dt, {});
        if (w->W->vehicle_id - before != 1) {
            // FP corner of the wrapper's flow trick: fall back to constructing the platoon directly
            while (w->W->vehicle_id - before < 1)
                new Vehicle(w->W, std::to_string(w->W->vehicle_id), t, std::to_string(orig[i]), std::to_string(dest[i]));
        }
    }
    return n;
    UXR_CATCH(w)
}

UX_API int uxr_set_vehicle_links(ux_world *w, int64_t vid, int kind, const int *links, int n) {
    UXR_TRY
    Vehicle *v = w->W->get_vehicle_by_index((int)vid);
    std::vector<Link *> ls;
    for (int i = 0; i < n; i++) {
        if (links[i] < 0 || links[i] >= w->W->link_id) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
        ls.push_back(w->W->links[links[i]]);
    }
    if (kind == 0) v->enforce_route(ls);
    else if (kind == 1) v->links_preferred = ls;
    else if (kind == 2) v->links_avoid = ls;
    else return fail(w, UX_ERR_INVALID, "bad kind");
    return 0;
    UXR_CATCH(w)
}

UX_API int uxr_batch_enforce_routes_csr(ux_world *w, const int *ids, const int64_t *off, int64_t nveh) {
    UXR_TRY
    if (nveh > (int64_t)w->W->vehicles.size()) nveh = (int64_t)w->W->vehicles.size();
    for (int64_t i = 0; i < nveh; i++) {
        if (off[i + 1] > off[i]) {
            int rc = uxr_set_vehicle_links(w, i, 0, ids + off[i], (int)(off[i + 1] - off[i]));
            if (rc) return rc;
        }
    }
    return 0;
    UXR_CATCH(w)
}

UX_API int uxr_set_t_max(ux_world *w, double t_max) {
    UXR_TRY
    w->p.t_max = t_max;
    w->W->set_t_max(t_max);
    return 0;
    UXR_CATCH(w)
}

UX_API int uxr_set_toll_timeseries(ux_world *w, int link, const double *toll, int n) {
    UXR_TRY
    if (link < 0 || link >= w->W->link_id) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    w->W->links[link]->toll_timeseries.assign(toll, toll + n);
    return 0;
    UXR_CATCH(w)
}

UX_API int uxr_initialize_adj_matrix(ux_world *w) {
    UXR_TRY
    w->W->set_t_max(w->p.t_max);
    w->W->initialize_adj_matrix();
    w->initialized = true;
    return 0;
    UXR_CATCH(w)
}

UX_API int uxr_main_loop(ux_world *w, double duration_t, double until_t) {
    UXR_TRY
    if (!w->initialized) { int rc = uxr_initialize_adj_matrix(w); if (rc) return rc; }
    w->W->main_loop(duration_t, until_t);
    return 0;
    UXR_CATCH(w)
}

UX_API int uxr_check_simulation_ongoing(ux_world *w) { return w->W->check_simulation_ongoing() ? 1 : 0; }

UX_API int uxr_set_link_param(ux_world *w, int link, int field, double value) {
    UXR_TRY
    if (link < 0 || link >= w->W->link_id) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    Link *l = w->W->links[link];
    switch (field) {
    case UX_LINK_VMAX: l->change_free_flow_speed(value); break;
    case UX_LINK_KAPPA: l->change_jam_density(value); break;
    case UX_LINK_CAPACITY_OUT: l->capacity_out = value; break;
    case UX_LINK_CAPACITY_IN: l->capacity_in = value; break;
    case UX_LINK_CAPACITY_OUT_REMAIN: l->capacity_out_remain = value; break;
    case UX_LINK_CAPACITY_IN_REMAIN: l->capacity_in_remain = value; break;
    case UX_LINK_MERGE_PRIORITY: l->merge_priority = value; break;
    case UX_LINK_ROUTE_CHOICE_PENALTY: l->route_choice_penalty = value; break;
    default: return fail(w, UX_ERR_INVALID, "unknown link field");
    }
    return 0;
    UXR_CATCH(w)
}

UX_API int uxr_set_link_signal_group(ux_world *w, int link, const int *sg, int n) {
    UXR_TRY
    if (link < 0 || link >= w->W->link_id) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    w->W->links[link]->signal_group.assign(sg, sg + n);
    return 0;
    UXR_CATCH(w)
}

UX_API int uxr_set_node_signal(ux_world *w, int node, const double *sig, int nsig, int phase, double sig_t) {
    UXR_TRY
    if (node < 0 || node >= w->W->node_id) return fail(w, UX_ERR_OUT_OF_RANGE, "node id out of range");
    Node *n = w->W->nodes[node];
    if (sig && nsig > 0) {
        n->signal_intervals.assign(sig, sig + nsig);
        if (n->signal_phase >= nsig) { n->signal_phase = 0; n->signal_t = 0; }
    }
    if (phase >= 0) n->signal_phase = phase;
    if (sig_t >= 0) n->signal_t = sig_t;
    return 0;
    UXR_CATCH(w)
}

UX_API int64_t uxr_get_int(ux_world *w, int what) {
    World *W = w->W;
    switch (what) {
    case UX_I_NUM_NODES: return W->node_id;
    case UX_I_NUM_LINKS: return W->link_id;
    case UX_I_NUM_VEHICLES: return W->vehicle_id;
    case UX_I_TIMESTEP: return (int64_t)W->timestep;
    case UX_I_TOTAL_TIMESTEPS: return (int64_t)W->total_timesteps;
    case UX_I_NUM_REPLICAS: return 1;
    case UX_I_ROUTE_UPDATE_STEPS: return (int64_t)W->timestep_for_route_update;
    case UX_I_PLATOON_UPDATES: { double s = 0; for (double c : W->link_stat_count) s += c; return (int64_t)s; }
    case UX_I_TRIPS_COMPLETED: return (int64_t)W->trips_completed_count();
    case UX_I_LOG_ENTRIES: { int64_t s = 0; for (auto *v : W->vehicles) s += (int64_t)v->log_size; return s; }
    case UX_I_KERNEL_LAUNCHES: return 0;
    case UX_I_NUM_ACTIVE_DESTS: return W->node_id;
    case UX_I_TRANSFER_LEVELS: return W->node_id;
    case UX_I_LINK_EVENTS: { int64_t s = 0; for (auto *l : W->links) s += (int64_t)l->traveltime_real_events.size(); return s; }
    case UX_I_MAX_DEPART_TS: { int m = 0; for (auto *v : W->vehicles) m = std::max(m, (int)(v->departure_time / W->delta_t)); return m; }
    default: return fail(w, UX_ERR_INVALID, "unknown int query");
    }
}

UX_API double uxr_get_double(ux_world *w, int what) {
    World *W = w->W;
    switch (what) {
    case UX_D_DTA_N_SWAP: return w->swap.n_swap;
    case UX_D_DTA_POTENTIAL_N_SWAP: return w->swap.potential_n_swap;
    case UX_D_DTA_TOTAL_T_GAP: return w->swap.total_t_gap;
    case UX_D_DELTA_T: return W->delta_t;
    case UX_D_TIME: return (double)W->timestep * W->delta_t;
    case UX_D_T_MAX: return W->t_max;
    case UX_D_AVE_V_SUM: return W->ave_v_sum();
    case UX_D_STAT_SAMPLE_COUNT: return W->stat_sample_count();
    default: return 0.0 / 0.0;
    }
}

static int eff_state(const Vehicle *v) {  // traffi.cpp:1947-1958
    int s = v->state;
    if (s == vsEND && (v->flag_trip_aborted || (v->arrival_time < 0 && v->travel_time <= 0))) s = vsABORT;
    return s;
}

UX_API int64_t uxr_array_size(ux_world *w, int what, int *dtype) {
    World *W = w->W;
    int64_t N = W->node_id, L = W->link_id, T = (int64_t)W->total_timesteps, P = W->vehicle_id;
    int dt = UX_DT_F64; int64_t n = -1;
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
    case UX_A_LOG_T: case UX_A_LOG_X: case UX_A_LOG_V: case UX_A_LOG_S: n = uxr_get_int(w, UX_I_LOG_ENTRIES); break;
    case UX_A_LOG_STATE: case UX_A_LOG_LANE: case UX_A_LOG_LINK: n = uxr_get_int(w, UX_I_LOG_ENTRIES); dt = UX_DT_I32; break;
    case UX_A_LOG_N_MISSING: n = P; dt = UX_DT_I32; break;
    case UX_A_LTL_T: case UX_A_LTL_ID: { auto fl = W->build_all_vehicle_logs_flat_compact(); n = (int64_t)fl.ltl_t.size(); dt = what == UX_A_LTL_T ? UX_DT_F64 : UX_DT_I32; break; }
    case UX_A_ENTER_LINK: case UX_A_ENTER_VEH: case UX_A_ENTER_TIME: { n = (int64_t)W->build_enter_log_data().size(); dt = what == UX_A_ENTER_TIME ? UX_DT_F64 : UX_DT_I32; break; }
    case UX_A_DTA_OD_ORIG: case UX_A_DTA_OD_DEST: n = (int64_t)w->od_keys.size(); dt = UX_DT_I32; break;
    case UX_A_DTA_OD_OFFSETS: n = (int64_t)w->od_keys.size() + 1; dt = UX_DT_I64; break;
    case UX_A_DTA_ROUTE_OFFSETS: { n = 1; for (auto &k : w->od_keys) n += (int64_t)w->store.sets[k].size(); dt = UX_DT_I64; break; }
    case UX_A_DTA_ROUTE_LINKS: { n = 0; for (auto &k : w->od_keys) for (auto &r : w->store.sets[k]) n += (int64_t)r.size(); dt = UX_DT_I32; break; }
    case UX_A_DTA_RS_OFFSETS: case UX_A_DTA_RA_OFFSETS: n = P + 1; dt = UX_DT_I64; break;
    case UX_A_DTA_RS_LINKS: { n = 0; for (auto &r : w->swap.routes_specified) n += (int64_t)r.size(); dt = UX_DT_I32; break; }
    case UX_A_DTA_RA_LINKS: { n = 0; for (auto &r : w->swap.route_actual) n += (int64_t)r.size(); dt = UX_DT_I32; break; }
    case UX_A_DTA_RA_ENTRY_T: n = (int64_t)w->swap_entry_t.size(); break;
    case UX_A_DTA_COST_ACTUAL: n = P; break;
    default: return fail(w, UX_ERR_INVALID, "unknown array");
    }
    if (dtype) *dtype = dt;
    return n;
}

UX_API int64_t uxr_get_array(ux_world *w, int what, int replica, void *dst, int64_t capacity) {
    UXR_TRY
    World *W = w->W;
    int dt; int64_t n = uxr_array_size(w, what, &dt);
    if (n < 0) return n;
    if (replica != 0) return fail(w, UX_ERR_OUT_OF_RANGE, "reference harness has a single replica");
    if (capacity < n) return fail(w, UX_ERR_INVALID, "get_array: capacity too small");
    double *d = (double *)dst; int *ii = (int *)dst; int64_t *ll = (int64_t *)dst;
    int64_t N = W->node_id, L = W->link_id, T = (int64_t)W->total_timesteps, P = W->vehicle_id;
    switch (what) {
    // Link.get_*_np bindings.cpp:819-835
    case UX_A_CUM_ARRIVAL: for (int64_t l = 0; l < L; l++) memcpy(d + l * T, W->links[l]->arrival_curve.data(), sizeof(double) * T); break;
    case UX_A_CUM_DEPARTURE: for (int64_t l = 0; l < L; l++) memcpy(d + l * T, W->links[l]->departure_curve.data(), sizeof(double) * T); break;
    case UX_A_TRAVELTIME_INSTANT: for (int64_t l = 0; l < L; l++) memcpy(d + l * T, W->links[l]->traveltime_instant.data(), sizeof(double) * T); break;
    case UX_A_TRAVELTIME_ACTUAL: for (int64_t l = 0; l < L; l++) { W->links[l]->ensure_traveltime_real(); memcpy(d + l * T, W->links[l]->traveltime_real.data(), sizeof(double) * T); } break;
    case UX_A_VEH_STATE: for (int64_t i = 0; i < P; i++) ii[i] = eff_state(W->vehicles[i]); break;
    case UX_A_VEH_ARRIVAL_TS: for (int64_t i = 0; i < P; i++) { auto *v = W->vehicles[i]; ii[i] = (v->arrival_time >= 0 && eff_state(v) != vsABORT) ? (int)(v->arrival_time / W->delta_t) : -1; } break;
    case UX_A_VEH_DEPART_TS: for (int64_t i = 0; i < P; i++) ii[i] = (int)(W->vehicles[i]->departure_time / W->delta_t); break;
    case UX_A_VEH_TRAVEL_TIME: for (int64_t i = 0; i < P; i++) d[i] = eff_state(W->vehicles[i]) == vsABORT ? -1.0 : W->vehicles[i]->travel_time; break;
    case UX_A_VEH_DISTANCE: for (int64_t i = 0; i < P; i++) d[i] = W->vehicles[i]->distance_traveled; break;
    case UX_A_VEH_LINK_ARRIVAL_TIME: for (int64_t i = 0; i < P; i++) d[i] = W->vehicles[i]->arrival_time_link; break;
    case UX_A_VEH_ORIG: for (int64_t i = 0; i < P; i++) ii[i] = W->vehicles[i]->orig->id; break;
    case UX_A_VEH_DEST: for (int64_t i = 0; i < P; i++) ii[i] = W->vehicles[i]->dest->id; break;
    case UX_A_VEH_LINK: for (int64_t i = 0; i < P; i++) ii[i] = W->vehicles[i]->link ? W->vehicles[i]->link->id : -1; break;
    case UX_A_VEH_X: for (int64_t i = 0; i < P; i++) d[i] = W->vehicles[i]->x; break;
    case UX_A_VEH_V: for (int64_t i = 0; i < P; i++) d[i] = W->vehicles[i]->v; break;
    case UX_A_VEH_LANE: for (int64_t i = 0; i < P; i++) ii[i] = W->vehicles[i]->lane; break;
    case UX_A_SIGNAL_LOG: for (int64_t k = 0; k < N; k++) for (int64_t t = 0; t < T; t++) ii[k * T + t] = t < (int64_t)W->nodes[k]->signal_log.size() ? W->nodes[k]->signal_log[t] : 0; break;
    case UX_A_NODE_SIGNAL_PHASE: for (int64_t k = 0; k < N; k++) ii[k] = W->nodes[k]->signal_phase; break;
    case UX_A_NODE_SIGNAL_T: for (int64_t k = 0; k < N; k++) d[k] = W->nodes[k]->signal_t; break;
    case UX_A_NODE_FLOW_CAPACITY: for (int64_t k = 0; k < N; k++) d[k] = W->nodes[k]->flow_capacity; break;
    case UX_A_NODE_NUM_LANES: for (int64_t k = 0; k < N; k++) ii[k] = W->nodes[k]->number_of_lanes; break;
    case UX_A_LINK_NUM_VEHICLES: for (int64_t l = 0; l < L; l++) ii[l] = (int)W->links[l]->vehicles.size(); break;
    case UX_A_LINK_CAPACITY_IN_REMAIN: for (int64_t l = 0; l < L; l++) d[l] = W->links[l]->capacity_in_remain; break;
    case UX_A_LINK_CAPACITY_OUT_REMAIN: for (int64_t l = 0; l < L; l++) d[l] = W->links[l]->capacity_out_remain; break;
    case UX_A_LINK_STAT_COUNT: for (int64_t l = 0; l < L; l++) d[l] = W->link_stat_count[l]; break;
    case UX_A_LINK_TRIPS_COMPLETED: for (int64_t l = 0; l < L; l++) d[l] = W->link_trips_completed[l]; break;
    case UX_A_LINK_AVE_V_SUM: for (int64_t l = 0; l < L; l++) d[l] = W->link_ave_v_sum[l]; break;
    case UX_A_ROUTE_PREF: for (int64_t k = 0; k < N; k++) for (int64_t l = 0; l < L; l++) d[k * L + l] = k < (int64_t)W->route_preference.size() ? W->route_preference[k][l] : 0.0; break;
    case UX_A_ROUTE_DIST: for (int64_t i = 0; i < N; i++) for (int64_t k = 0; k < N; k++) d[i * N + k] = i < (int64_t)W->route_dist.size() ? W->route_dist[i][k] : 0.0; break;
    case UX_A_ROUTE_NEXT: for (int64_t i = 0; i < N; i++) for (int64_t k = 0; k < N; k++) ii[i * N + k] = i < (int64_t)W->route_next.size() ? W->route_next[i][k] : -1; break;
    case UX_A_DTA_OD_ORIG: for (size_t i = 0; i < w->od_keys.size(); i++) ii[i] = w->od_keys[i].first; break;
    case UX_A_DTA_OD_DEST: for (size_t i = 0; i < w->od_keys.size(); i++) ii[i] = w->od_keys[i].second; break;
    case UX_A_DTA_OD_OFFSETS: { int64_t a = 0; ll[0] = 0; for (size_t i = 0; i < w->od_keys.size(); i++) { a += (int64_t)w->store.sets[w->od_keys[i]].size(); ll[i + 1] = a; } break; }
    case UX_A_DTA_ROUTE_OFFSETS: { int64_t a = 0, q = 0; ll[q++] = 0; for (auto &k : w->od_keys) for (auto &r : w->store.sets[k]) { a += (int64_t)r.size(); ll[q++] = a; } break; }
    case UX_A_DTA_ROUTE_LINKS: { int64_t q = 0; for (auto &k : w->od_keys) for (auto &r : w->store.sets[k]) for (int l : r) ii[q++] = l; break; }
    case UX_A_DTA_RS_OFFSETS: case UX_A_DTA_RA_OFFSETS: {
        if (!w->has_swap) return fail(w, UX_ERR_RUNTIME, "no route swap result");
        auto &rr = what == UX_A_DTA_RS_OFFSETS ? w->swap.routes_specified : w->swap.route_actual;
        int64_t a = 0; ll[0] = 0; for (int64_t i = 0; i < P; i++) { a += (int64_t)rr[i].size(); ll[i + 1] = a; } break; }
    case UX_A_DTA_RS_LINKS: case UX_A_DTA_RA_LINKS: {
        auto &rr = what == UX_A_DTA_RS_LINKS ? w->swap.routes_specified : w->swap.route_actual;
        int64_t q = 0; for (auto &r : rr) for (int l : r) ii[q++] = l; break; }
    case UX_A_DTA_RA_ENTRY_T: memcpy(d, w->swap_entry_t.data(), sizeof(double) * w->swap_entry_t.size()); break;
    case UX_A_DTA_COST_ACTUAL: if (!w->has_swap) return fail(w, UX_ERR_RUNTIME, "no route swap result"); memcpy(d, w->swap.cost_actual.data(), sizeof(double) * P); break;
    case UX_A_ENTER_LINK: case UX_A_ENTER_TIME: case UX_A_ENTER_VEH: {
        auto e = W->build_enter_log_data();
        for (size_t i = 0; i < e.size(); i++) { if (what == UX_A_ENTER_LINK) ii[i] = e[i].link_id; else if (what == UX_A_ENTER_TIME) d[i] = e[i].time; else ii[i] = e[i].vehicle_index; }
        break; }
    default: { // World::build_all_vehicle_logs_flat_compact bindings.cpp:579-627
        auto fl = W->build_all_vehicle_logs_flat_compact();
        switch (what) {
        case UX_A_LOG_OFFSETS: for (size_t i = 0; i < fl.offsets.size(); i++) ll[i] = (int64_t)fl.offsets[i]; break;
        case UX_A_LTL_OFFSETS: for (size_t i = 0; i < fl.ltl_offsets.size(); i++) ll[i] = (int64_t)fl.ltl_offsets[i]; break;
        case UX_A_LOG_T: memcpy(d, fl.log_t.data(), sizeof(double) * fl.log_t.size()); break;
        case UX_A_LOG_X: memcpy(d, fl.log_x.data(), sizeof(double) * fl.log_x.size()); break;
        case UX_A_LOG_V: memcpy(d, fl.log_v.data(), sizeof(double) * fl.log_v.size()); break;
        case UX_A_LOG_S: memcpy(d, fl.log_s.data(), sizeof(double) * fl.log_s.size()); break;
        case UX_A_LOG_STATE: memcpy(ii, fl.log_state.data(), sizeof(int) * fl.log_state.size()); break;
        case UX_A_LOG_LANE: memcpy(ii, fl.log_lane.data(), sizeof(int) * fl.log_lane.size()); break;
        case UX_A_LOG_LINK: memcpy(ii, fl.log_link.data(), sizeof(int) * fl.log_link.size()); break;
        case UX_A_LOG_N_MISSING: memcpy(ii, fl.n_missing.data(), sizeof(int) * fl.n_missing.size()); break;
        case UX_A_LTL_T: memcpy(d, fl.ltl_t.data(), sizeof(double) * fl.ltl_t.size()); break;
        case UX_A_LTL_ID: memcpy(ii, fl.ltl_id.data(), sizeof(int) * fl.ltl_id.size()); break;
        default: return fail(w, UX_ERR_INVALID, "unknown array");
        }
    }
    }
    return n;
    UXR_CATCH(w)
}

UX_API int64_t uxr_get_device_array(ux_world *w, int, int, void **, int *) { return fail(w, UX_ERR_RUNTIME, "the reference harness has no device arrays"); }

UX_API int64_t uxr_ensemble_link_stats(ux_world *w, double *dst, int64_t capacity) {
    UXR_TRY
    World *W = w->W; int64_t L = W->link_id, T = (int64_t)W->total_timesteps;
    if (capacity < L * 4) return fail(w, UX_ERR_INVALID, "capacity too small");
    for (int64_t l = 0; l < L; l++) {
        Link *k = W->links[l]; k->ensure_traveltime_real();
        double s1 = 0, s2 = 0; for (int64_t t = 0; t < T; t++) { s1 += k->traveltime_real[t]; s2 += k->traveltime_real[t] * k->traveltime_real[t]; }
        dst[l * 4 + 0] = T > 0 ? k->arrival_curve[T - 1] : 0.0; dst[l * 4 + 1] = s1; dst[l * 4 + 2] = s2; dst[l * 4 + 3] = (double)k->vehicles.size();
    }
    return L * 4;
    UXR_CATCH(w)
}
UX_API int64_t uxr_ensemble_link_stats_device(ux_world *w, void **) { return fail(w, UX_ERR_RUNTIME, "the reference harness has no device arrays"); }

#define PFX(n) uxr_##n
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
    for (int64_t i = 0; i < n; i++) { int64_t m = PFX(add_demand)(w, o[i], d[i], t0[i], t1[i], flow[i], nullptr, 0); if (m < 0) return m; made += m; }
    return made;
}
UX_API int PFX(set_option)(ux_world *w, int option, double value) { (void)w; (void)option; (void)value; return 0; }

// ---- DTA solver batch operations: forwarding calls into the reference's dta_solver.cpp
static void index_store(ux_world *w) {
    w->od_keys.clear();
    for (auto &kv : w->store.sets) w->od_keys.push_back(kv.first);
    std::sort(w->od_keys.begin(), w->od_keys.end());
}
// enumerate_od_route_sets_ids_cpp bindings.cpp:461-465
UX_API int64_t uxr_dta_enumerate_route_sets(ux_world *w, int k, unsigned int seed) {
    UXR_TRY
    if (!w->initialized) { int rc = uxr_initialize_adj_matrix(w); if (rc) return rc; }
    w->store = dta_enumerate_od_route_sets_ids(w->W, k, seed);
    index_store(w);
    int64_t n = 0; for (auto &kv : w->store.sets) n += (int64_t)kv.second.size();
    return n;
    UXR_CATCH(w)
}
// make_od_route_set_store bindings.cpp:467-489
UX_API int uxr_dta_set_route_sets(ux_world *w, int64_t n_od, const int *orig, const int *dest, const int64_t *od_off, const int64_t *route_off, const int *links) {
    UXR_TRY
    w->store.sets.clear();
    for (int64_t p = 0; p < n_od; p++) {
        std::vector<std::vector<int>> routes;
        for (int64_t r = od_off[p]; r < od_off[p + 1]; r++) routes.emplace_back(links + route_off[r], links + route_off[r + 1]);
        w->store.sets[{orig[p], dest[p]}] = std::move(routes);
    }
    index_store(w);
    return 0;
    UXR_CATCH(w)
}
UX_API int uxr_dta_share_route_sets(ux_world *w, ux_world *from) {
    if (!from) return fail(w, UX_ERR_INVALID, "dta_share_route_sets: no source world");
    if (from != w) { w->store = from->store; w->od_keys = from->od_keys; }
    return 0;
}
UX_API int uxr_analysis_summary(ux_world *w, int) { return fail(w, UX_ERR_RUNTIME, "od_analysis / link_analysis_coarse are computed by the reference's Python analyzer (analyzer.py:121-203), not by its C++ engine"); }
UX_API int uxr_edie_state(ux_world *w, int, const double *, const double *) { return fail(w, UX_ERR_RUNTIME, "Edie states are computed by the reference's Python analyzer (analyzer.py:249-330), not by its C++ engine"); }
UX_API int uxr_dta_next_day(ux_world *w) { return fail(w, UX_ERR_RUNTIME, "the reference builds a new World per day (DTAsolvers_cpp_adapter.py:318-330)"); }
UX_API int uxr_dta_enforce_routes(ux_world *w, int replica, const int *ids, const int64_t *off, int64_t nveh) {
    if (replica != 0) return fail(w, UX_ERR_OUT_OF_RANGE, "reference harness has a single replica");
    return uxr_batch_enforce_routes_csr(w, ids, off, nveh);
}
// route_swap_due / route_swap_dso bindings.cpp:661-706
UX_API int uxr_dta_route_swap(ux_world *w, int replica, int mode, double swap_prob, int swap_num, int has_external, unsigned int rng_seed) {
    UXR_TRY
    if (replica != 0) return fail(w, UX_ERR_OUT_OF_RANGE, "reference harness has a single replica");
    if (mode == 0) w->swap = dta_route_swap_due(w->W, w->store, swap_prob, has_external != 0, rng_seed);
    else w->swap = dta_route_swap_dso(w->W, w->store, swap_prob, swap_num, has_external != 0, rng_seed);
    w->swap_entry_t.clear();
    for (Vehicle *v : w->W->vehicles) { auto tr = dta_get_traveled_route(v); w->swap_entry_t.insert(w->swap_entry_t.end(), tr.entry_times.begin(), tr.entry_times.end()); }
    w->has_swap = true;
    return 0;
    UXR_CATCH(w)
}

UX_API const char *uxr_last_error(ux_world *w) { return w ? w->err.c_str() : g_err.c_str(); }

}  // extern "C"
