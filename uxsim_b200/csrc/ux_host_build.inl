// ux_host_build.inl -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// C-ABI, part 1 (inside extern "C"): world creation, scenario ingest, finalisation, upload of the device state, mid-run vehicle additions.


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
    if (option == UX_OPT_ABORT_VEHICLE) {
        long long vid = (long long)value;
        if (vid < 0 || vid >= (long long)w->vehs.size()) return fail(w, UX_ERR_OUT_OF_RANGE, "vehicle index out of range");
        if (w->uploaded) return fail(w, UX_ERR_RUNTIME, "UX_OPT_ABORT_VEHICLE must be set before the first main_loop");
        HVeh &v = w->vehs[vid];
        if (!v.aborted) { v.aborted = true; v.ts0 = v.ts; v.ts = 0x3fffffff; }  // never released from its origin
        return 0;
    }
    if (option == UX_OPT_ROUTE_TABLES_AT) { w->route_tables_at = value < 0.0 ? -1 : (int)value; return 0; }
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
    for (auto &l : w->links) l.tt0 = l.length / l.vmax;  // traveltime_actual before the first exit event (traffi.cpp:519)
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
        // Ensembles (k_node_h): heavy nodes -- many lanes in or out -- are stepped by a warp each, light ones by a thread per replica;
        // inside a height the heavy ones come first, so that the block table is a few long runs of one form
        std::vector<char> heavy(N, 0);
        if (w->node_rl) for (int a = 0; a < N; a++) {
            int li = 0, lo = 0;
            for (int l : w->nodes[a].in) li += w->links[l].lanes;
            for (int l : w->nodes[a].out) lo += w->links[l].lanes;
            heavy[a] = std::max(li, lo) > w->node_heavy_lanes;
        }
        std::stable_sort(lvl_nodes.begin(), lvl_nodes.end(), [&](int a, int b) {
            if (height[a] != height[b]) return height[a] > height[b];
            if (heavy[a] != heavy[b]) return heavy[a] > heavy[b];
            return weight[a] > weight[b]; });
        w->h_node_blks.clear();
        if (w->node_rl) {  // {form, first position in lvl_nodes, nodes, replica | lane group}, four nodes per block
            const int groups = (w->R + 31) / 32;
            for (int q = 0; q < N;) {
                int e = q + 1;
                while (e < N && e - q < (heavy[lvl_nodes[q]] ? NRL_HEAVY_WARPS : NRL_WARPS) && heavy[lvl_nodes[e]] == heavy[lvl_nodes[q]] && height[lvl_nodes[e]] == height[lvl_nodes[q]]) e++;
                if (heavy[lvl_nodes[q]]) for (int r = 0; r < w->R; r++) w->h_node_blks.push_back(make_int4(0, q, e - q, r));
                else for (int g = 0; g < groups; g++) w->h_node_blks.push_back(make_int4(1, q, e - q, g));
                q = e;
            }
        }
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
    if (w->node_rl) {  // block table of k_node_h (a changed table invalidates the captured graphs: their grids are part of them)
        if (w->h_node_blks.size() > w->node_blks_cap) { w->node_blks_cap = w->h_node_blks.size() + 64; w->d_node_blks = dalloc<int4>(w, w->node_blks_cap, false); w->rewindable = w->rewindable && !w->uploaded; }
        if (!w->d_node_blks) return fail(w, UX_ERR_CUDA, "device allocation failed");
        CUDA_OK(cudaMemcpyAsync(w->d_node_blks, w->h_node_blks.data(), sizeof(int4) * w->h_node_blks.size(), cudaMemcpyHostToDevice, w->stream));
#ifndef UX_EMU
        if ((int)w->h_node_blks.size() != w->n_node_blks) { for (auto &g : w->graphs) cudaGraphExecDestroy(g.second.first); w->graphs.clear(); }
#endif
        w->n_node_blks = (int)w->h_node_blks.size();
    }
    lvl_off.resize(w->nodes.size() + 2, lvl_off.back());
    CUDA_OK(cudaMemcpyAsync(w->d_l_flags, lflags.data(), sizeof(int) * L, cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_lvl_off, lvl_off.data(), sizeof(int) * lvl_off.size(), cudaMemcpyHostToDevice, w->stream));
    CUDA_OK(cudaMemcpyAsync(w->d_lvl_nodes, lvl_nodes.data(), sizeof(int) * lvl_nodes.size(), cudaMemcpyHostToDevice, w->stream));
    return 0;
}

// Signal plans (per node) and signal groups (per link) are CSR tables; a setter that changes the LENGTH of one list after the
// upload rebuilds both tables here (new arrays from the slabs, pointers swapped in every replica's DevWorld).
static int upload_signal_tables(ux_world *w) {
    w->rewindable = false;  // the new tables come from the slab tail, inside the range a day rewind would zero
    std::vector<int> sg_off(w->links.size() + 1, 0), sg, sig_off(w->nodes.size() + 1, 0);
    std::vector<double> n_sig;
    for (size_t i = 0; i < w->links.size(); i++) { sg.insert(sg.end(), w->links[i].sg.begin(), w->links[i].sg.end()); sg_off[i + 1] = (int)sg.size(); }
    for (size_t i = 0; i < w->nodes.size(); i++) { n_sig.insert(n_sig.end(), w->nodes[i].sig.begin(), w->nodes[i].sig.end()); sig_off[i + 1] = (int)n_sig.size(); }
    w->d_sg_off = dupload(w, sg_off); w->d_sg = dupload(w, sg); w->d_n_sig_off = dupload(w, sig_off); w->d_n_sig = dupload(w, n_sig);
    if (!w->d_sg_off || !w->d_sg || !w->d_n_sig_off || !w->d_n_sig) return fail(w, UX_ERR_CUDA, "device allocation of the signal tables failed");
    for (auto &d : w->hw) { d.l_sg_off = w->d_sg_off; d.l_sg = w->d_sg; d.n_sig_off = w->d_n_sig_off; d.n_sig = w->d_n_sig; }
    CUDA_OK(cudaMemcpyAsync(w->dworlds, w->hw.data(), sizeof(DevWorld) * w->hw.size(), cudaMemcpyHostToDevice, w->stream));
    return 0;
}

// the non-zero initial values of the replica-minor state, broadcast over the replicas (upload, and every day rewind)
static void launch_init_dyn(ux_world *w) {
    int L = (int)w->links.size(), N = (int)w->nodes.size();
    long long tot = (long long)(L + N) * w->R;
    if (tot > 0) UX_LAUNCH(k_init_dyn, dim3((unsigned)((tot + 255) / 256)), dim3(256), w->stream, w->dworlds, w->R, w->d_init_co, w->d_init_ci, w->d_init_ph, w->d_init_st, w->d_init_fc);
    long long per = std::max((long long)w->P_uploaded, (long long)w->n_active * N);
    if (per > 0) UX_LAUNCH(k_init_rep, dim3((unsigned)((per + 255) / 256), w->R), dim3(256), w->stream, w->dworlds);
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
    {   // a separate launch for the scalar bookkeeping pays off once it has ~40k threads of work (B200: Chicago x32 -3 %, x128 -7 %; Sioux Falls x64 +9 % if forced)
        const char *e = getenv("UXSIM_B200_PRE_SPLIT");
        w->pre_split = e ? atoi(e) != 0 : ((long long)(L + N) * w->R >= 40000);
        base.pre_split = w->pre_split ? 1 : 0;
        const char *e2 = getenv("UXSIM_B200_NODE");  // rl | warp
        // Node step: warp per (node, replica) (k_node), or -- large ensembles -- thread per (node, replica) (k_node_h).  Measured on the
        // B200, Chicago sketch, us per launch (warp | thread form): 32 replicas 127 | 546, 128: 320 | 647, 256: 609 | ~600, 512: 1210 | 820 --
        // the thread form is bound by the serial chain of its busiest node (flat in R), the warp form by residency (linear in R).
        // UXSIM_B200_NODE = warp | rl | hybrid (hybrid: nodes with more than UXSIM_B200_NODE_HEAVY lanes in or out by warps, inside k_node_h)
        w->node_rl = e2 ? strcmp(e2, "warp") != 0 : w->R >= 256;
        const char *e3 = getenv("UXSIM_B200_NODE_HEAVY");
        w->node_heavy_lanes = e3 ? atoi(e3) : (e2 && !strcmp(e2, "hybrid")) ? 8 : 1 << 30;
        base.node_rl = w->node_rl ? 1 : 0;
        if (w->node_rl) { w->pre_split = true; base.pre_split = 1; }  // the scalar bookkeeping always runs in k_link_pre<true> then
    }
    base.vis_words = (N + 31) / 32;
    // ---- links
    std::vector<int> l_start(L), l_end(L), l_lanes(L), l_rcap(L), sg_off(L + 1, 0), sg;
    std::vector<long long> l_roff(L);
    std::vector<double> l_length(L);
    long long C = 0;
    std::vector<int> seg_link, seg_start;
    {   // two work items per warp in k_update where queues are short and there are enough items to fill the GPU either way
        const char *e = getenv("UXSIM_B200_UPDATE_GROUP");
        w->upd_group = e ? (atoi(e) == 16 ? 16 : 32) : ((P <= 32ll * L && (long long)L * w->R >= 20000) ? 16 : 32);
    }
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
        {   // work items of the update kernel: one lane group per 16 * (lanes of a group) ring slots, at least one per link
            int nparts = (int)((cap + 16 * w->upd_group - 1) / (16 * w->upd_group));
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
    { std::vector<double> tt0(L); for (int i = 0; i < L; i++) tt0[i] = w->links[i].tt0; base.l_tt0 = dupload(w, tt0); }
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
        std::vector<int> e_inpos(Ep), e_outpos(Ep), eout_off(N + 1, 0), eout_to(Ep), order(Ep);
        for (int q = 0; q < Ep; q++) e_inpos[ein_eid[q]] = q;
        for (int e = 0; e < Ep; e++) { eout_off[e_from[e] + 1]++; order[e] = e; }
        for (int i = 0; i < N; i++) eout_off[i + 1] += eout_off[i];
        std::sort(order.begin(), order.end(), [&](int a, int b) { return e_from[a] != e_from[b] ? e_from[a] < e_from[b] : e_to[a] < e_to[b]; });
        for (int q = 0; q < Ep; q++) { e_outpos[order[q]] = q; eout_to[q] = e_to[order[q]]; }
        base.eout_off = dupload(w, eout_off); base.eout_to = dupload(w, eout_to); base.e_inpos = dupload(w, e_inpos); base.e_outpos = dupload(w, e_outpos);
    }
    // ---- nodes
    std::vector<int> out_off(N + 1, 0), in_off(N + 1, 0), out, in, sig_off(N + 1, 0), n_lanes(N), n_out_lanes(N), out_woff;
    std::vector<double> n_sig, n_fc(N);
    for (int i = 0; i < N; i++) {
        const HNode &n = w->nodes[i];
        out_off[i + 1] = out_off[i] + (int)n.out.size(); in_off[i + 1] = in_off[i] + (int)n.in.size();
        out.insert(out.end(), n.out.begin(), n.out.end()); in.insert(in.end(), n.in.begin(), n.in.end());
        sig_off[i + 1] = sig_off[i] + (int)n.sig.size(); n_sig.insert(n_sig.end(), n.sig.begin(), n.sig.end());
        n_fc[i] = n.fc; n_lanes[i] = n.lanes;
        int tl = 0; for (int l : n.out) { out_woff.push_back(tl); tl += w->links[l].lanes; } n_out_lanes[i] = tl;  // out_woff: lanes of the node's earlier out-links
    }
    base.n_out_off = dupload(w, out_off); base.n_out = dupload(w, out); base.n_in_off = dupload(w, in_off); base.n_in = dupload(w, in);
    w->d_n_sig_off = dupload(w, sig_off); w->d_n_sig = dupload(w, n_sig);
    base.n_sig_off = w->d_n_sig_off; base.n_sig = w->d_n_sig; base.n_fc = dupload(w, n_fc); base.n_lanes = dupload(w, n_lanes); base.n_out_lanes = dupload(w, n_out_lanes); base.n_out_woff = dupload(w, out_woff);
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
    // ---- dynamic state
    std::vector<double> co_rem(L), ci_rem(L), sig_t(N), fc_rem(N);
    std::vector<int> sig_ph(N);
    for (int i = 0; i < L; i++) { co_rem[i] = w->links[i].cap_out_rem0; ci_rem[i] = w->links[i].cap_in_rem0; }
    for (int i = 0; i < N; i++) { sig_t[i] = w->nodes[i].sig_t0; sig_ph[i] = w->nodes[i].sig_phase0; fc_rem[i] = w->nodes[i].fc_rem0; }
    // the non-zero initial values, kept on the device: k_init_dyn broadcasts them over the replicas (at upload and at every day rewind)
    w->d_init_co = dupload(w, co_rem); w->d_init_ci = dupload(w, ci_rem); w->d_init_ph = dupload(w, sig_ph); w->d_init_st = dupload(w, sig_t); w->d_init_fc = dupload(w, fc_rem);
    const int R = w->R, RS = (R + 7) & ~7;  // rows of 4-byte elements start on 32-byte sectors
    base.RS = RS;
    w->hw.assign(R, base);
    for (auto &kv : w->rep_routes) {  // per-replica enforced routes (ux_dta_enforce_routes): specified_route and links_preferred of that replica
        DevWorld &d = w->hw[kv.first];
        const int *o_ = dupload(w, kv.second.first), *l_ = dupload(w, kv.second.second);
        d.vl_off[0] = d.vl_off[1] = o_; d.vl_list[0] = d.vl_list[1] = l_;
    }
    // Everything allocated from here to the second mark is DYNAMIC state: zero, except what k_init_dyn and the ReInit list
    // restore.  ux_dta_next_day rewinds a finished world by zeroing these slab ranges and replaying both.
    dalloc<char>(w, 1);  // make sure the marks below sit inside a slab
    w->dyn_slab0 = w->slabs.size() - 1; w->dyn_used0 = w->slabs.back().used;
    w->reinit.clear();
    w->d_err_all = dalloc<int>(w, 2 * (size_t)R);
    if (!w->d_err_all) return fail(w, UX_ERR_CUDA, "device allocation failed");
    size_t TL = (size_t)T * L;
    // (1) replica-minor state: one array per field for ALL replicas, element (entity, replica) at [entity * RS + replica]
    LinkQ *lq = dalloc<LinkQ>(w, (size_t)L * RS); LinkC *lc = dalloc<LinkC>(w, (size_t)L * RS); LinkR *lr = dalloc<LinkR>(w, (size_t)L * RS);
    long long *stat = dalloc<long long>(w, (size_t)L * RS);
    NodeS *ns = dalloc<NodeS>(w, (size_t)(N + 2) * RS); NodeR *nr = dalloc<NodeR>(w, (size_t)(N + 2) * RS); int *node_done = dalloc<int>(w, (size_t)(N + 2) * RS);
    double *cum_arr = dalloc<double>(w, TL * RS), *cum_dep = dalloc<double>(w, TL * RS), *tt_inst = dalloc<double>(w, TL * RS), *ev_tt = dalloc<double>(w, TL * RS);
    int *ev_ord = dalloc<int>(w, TL * RS), *sig_log = dalloc<int>(w, (size_t)T * N * RS), *t_base = dalloc<int>(w, RS);
    if (!lq || !lc || !lr || !stat || !ns || !nr || !node_done || !cum_arr || !cum_dep || !tt_inst || !ev_tt || !ev_ord || !sig_log || !t_base)
        return fail(w, UX_ERR_CUDA, "device allocation failed (replica-minor state, %d replicas)", R);
    { std::vector<int> tb(RS, w->timestep); CUDA_OK(cudaMemcpyAsync(t_base, tb.data(), sizeof(int) * RS, cudaMemcpyHostToDevice, w->stream)); }
    // (2) replica-major state: rings, per-platoon arrays, route tables, logs -- contiguous per replica, and at the SAME offsets
    // inside one block per replica, so the array of replica r is the array of replica 0 + r * rep_stride bytes (the kernels
    // whose lanes are replicas compute the address instead of loading 32 pointer tables)
    long long lg_cap = 0, tv_cap = 0;
    if (base.log_mode) for (long long i = 0; i < P; i++) lg_cap += std::max(0, T - w->vehs[i].ts - 1);  // a platoon can be RUN from ts+1 to T-1
    if (w->record_traversals) {  // at most one link entry per platoon and step; in practice a few dozen per trip
        long long bound = 0;
        for (long long i = 0; i < P; i++) bound += std::max(0, T - w->vehs[i].ts - 1);
        tv_cap = std::min(bound, std::max(P * 48, (long long)1 << 20));
    }
    auto layout = [&](DevWorld &d, char *blk) -> size_t {
        size_t off = 0;
        auto take = [&](size_t bytes) -> char * { off = (off + 255) & ~(size_t)255; char *q = blk ? blk + off : nullptr; off += bytes ? bytes : 1; return q; };
#define TAKE(T_, n_) ((T_ *)take(sizeof(T_) * (size_t)(n_)))
        d.X = TAKE(double, 2 * (size_t)C); d.V = TAKE(double, C); d.VID = TAKE(int, C); d.LANE = TAKE(int, C);
        d.v_state = TAKE(int, P); d.v_next = TAKE(int, P); d.v_link = TAKE(int, P); d.v_arr_ts = TAKE(int, P); d.v_trav = TAKE(int, P);
        d.v_mr = TAKE(double, P); d.v_atl = TAKE(double, P); d.v_dist = TAKE(double, P); d.v_entry_x = TAKE(double, P); d.v_tt = TAKE(double, P);
        d.v_flags = TAKE(unsigned char, P);
        if (base.log_mode) {
            d.lg_cap = lg_cap;
            d.v_gen_ts = TAKE(int, P); d.v_end_ts = TAKE(int, P); d.v_end_kind = TAKE(unsigned char, P);
            d.lg_vid = TAKE(int, lg_cap); d.lg_t = TAKE(int, lg_cap); d.lg_link = TAKE(int, lg_cap); d.lg_lane = TAKE(int, lg_cap);
            d.lg_x = TAKE(double, lg_cap); d.lg_v = TAKE(double, lg_cap); d.lg_s = TAKE(double, lg_cap);
            d.lg_count = TAKE(unsigned long long, 1);
        }
        if (w->record_traversals) { d.tv_cap = tv_cap; d.tv_ev = TAKE(int4, tv_cap); d.tv_count = TAKE(unsigned long long, 1); }
        d.v_visited = base.no_cyclic ? TAKE(unsigned, (size_t)P * base.vis_words) : nullptr;
        d.pref = TAKE(double, (size_t)D * L); d.rdist = TAKE(double, (size_t)D * N); d.pair_cost = TAKE(double, w->n_edges);
        d.cost_in = TAKE(double, w->n_edges); d.cost_out = TAKE(double, w->n_edges); d.sssp_delta = TAKE(double, 1);
        d.cost_hist = TAKE(double, (size_t)((T + w->route_ts - 1) / w->route_ts + 1) * (size_t)w->n_edges);
        d.pref_active = TAKE(int, D); d.has_path = TAKE(int, D); d.rnext = TAKE(int, (size_t)D * N);
#undef TAKE
        return (off + 255) & ~(size_t)255;
    };
    DevWorld probe = base;
    const size_t rep_bytes = layout(probe, nullptr);
    char *rep_blk = dalloc<char>(w, rep_bytes * (size_t)R);
    if (!rep_blk) return fail(w, UX_ERR_CUDA, "device allocation failed (%d replicas x %zu bytes of per-replica state)", R, rep_bytes);
    for (int r = 0; r < R; r++) {
        DevWorld &d = w->hw[r];
        d.seed = w->p.random_seed + r; d.rep = r; d.rep_stride = (long long)rep_bytes;
        d.lq = lq + r; d.lc = lc + r; d.lr = lr + r; d.stat_count = stat + r; d.ns = ns + r; d.nr = nr + r; d.node_done = node_done + r;
        d.cum_arr = cum_arr + r; d.cum_dep = cum_dep + r; d.tt_inst = tt_inst + r; d.ev_tt = ev_tt + r; d.ev_ord = ev_ord + r; d.sig_log = sig_log + r; d.t_base = t_base + r;
        layout(d, rep_blk + (size_t)r * rep_bytes);
        d.err = w->d_err_all + 2 * r; d.n_levels = w->n_levels;
    }
    w->uniform = true;
    w->dyn_slab1 = w->slabs.size() - 1; w->dyn_used1 = w->slabs.back().used;
    w->rewindable = w->timestep == 0;
    w->dworlds = dalloc<DevWorld>(w, R, false);
    if (!w->dworlds) return fail(w, UX_ERR_CUDA, "device allocation failed");
    CUDA_OK(cudaMemcpyAsync(w->dworlds, w->hw.data(), sizeof(DevWorld) * R, cudaMemcpyHostToDevice, w->stream));
    launch_init_dyn(w);
    CUDA_OK(cudaStreamSynchronize(w->stream));
    w->uploaded = true;
    w->upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return 0;
}

static int launch_sssp(ux_world *w, cudaStream_t st) {  // returns the number of kernels launched
    int N = (int)w->nodes.size();
#ifdef UX_EMU
    UX_LAUNCH(k_sssp<false>, dim3(w->n_active, w->R), dim3(SSSP_THREADS), st, w->dworlds);
#else
    size_t bytes = (size_t)N * 12 + 16, fbytes = (size_t)((N + 3) & ~3) * 17 + 64;
    int use_smem = bytes <= 200 * 1024;
    // small graphs converge in a few dozen whole-edge sweeps (cheaper than one barrier per wavefront); large sparse graphs
    // are dominated by sweeps x E and switch to a worklist kernel: k_sssp_wave (labels in L2, ~10 destinations per SM), or
    // k_sssp_frontier (labels in shared memory, one destination per SM; UXSIM_B200_SSSP=frontier keeps it reachable for tests)
    const int mode = [] { const char *e = getenv("UXSIM_B200_SSSP"); return !e ? 0 : !strcmp(e, "sweep") ? 1 : !strcmp(e, "frontier") ? 2 : !strcmp(e, "wave") ? 3 : 0; }();
    const int wave_pad = [] { const char *e = getenv("UXSIM_B200_SSSP_WAVE_PAD"); return e ? atoi(e) : -1; }();
    bool large = N >= 2048 && N < 65536;
    // mid-size graphs (Chicago sketch: 546 nodes) already gain from the worklist order when the labels stay in shared memory
    // (817 -> 698 us per update at 32 replicas); below ~256 nodes a handful of whole-edge sweeps is cheaper
    bool mid = N >= 256 && N < 2048 && (long long)w->n_active * w->R >= 4096;  // a single scenario is one wave of blocks: the sweeps' shorter chain wins (56 vs 135 us)
    if ((mode == 3 || (mode == 0 && (large || mid))) && N < 65536) {
        // shared memory per block = queue + bitmap, padded so that the dist rows of the resident blocks stay within ~96 MB of L2
        size_t wbytes = (size_t)((N + 2 * SSSP_WAVE_THREADS + 3) & ~3) * 2 + (size_t)((N + 31) / 32) * 4 + 16;
        if (N <= 2048) {  // small / mid-size graph: labels in shared memory too; one destination per warp, four per block
            static const double factor_s = [] { const char *e = getenv("UXSIM_B200_SSSP_DELTA"); return e ? atof(e) : 2.0; }();
            static const int per_block = [] { const char *e = getenv("UXSIM_B200_SSSP_WAVE_THREADS"); return e ? atoi(e) : 0; }();  // > 0: one destination per block of that many threads
            k_cost_stats<<<dim3(1, w->R), dim3(256), 0, st>>>(w->dworlds, factor_s);
            if (per_block >= 64) {
                k_sssp_wave<true, false><<<dim3(w->n_active, w->R), dim3(per_block), wbytes + (size_t)N * 8, st>>>(w->dworlds, 0);
            } else {
                size_t ww = (((size_t)((N + 64 + 3) & ~3) * 2 + (size_t)((N + 31) / 32) * 4 + 15) / 8) + (size_t)N + 2;  // 8-byte words per warp
                {   // UXSIM_B200_SSSP_STAGE=1: the in-edge CSR and the replica's edge costs staged in shared memory, as many destinations
                    // (warps) per block as let two blocks share an SM.  Measured SLOWER on the B200 (Chicago x1024: 15.3 vs 14.3 ms per update:
                    // the kernel is bound by instruction issue, 74 % of the issue slots, not by the L1 hits the staging saves), so it is
                    // off by default and kept for the tests that compare the variants.
                    const int stage_on = [] { const char *e = getenv("UXSIM_B200_SSSP_STAGE"); return e ? atoi(e) : 0; }();  // (read per call: the tests switch it)
                    size_t sw = ((size_t)w->n_edges * 8 + (size_t)(N + 1) * 4 + (size_t)w->n_edges * 2 + 7) / 8;  // 8-byte words of the staged tables
                    long long room = (long long)110 * 1024 - (long long)sw * 8;
                    int wpb = room > 0 ? (int)std::min<long long>(SSSP_STAGE_MAX_WARPS, room / (long long)(ww * 8)) : 0;
                    if (stage_on && wpb >= 8 && (long long)w->n_active * w->R >= 16384) {
                        size_t bytes_s = (sw + ww * (size_t)wpb) * 8;
                        cudaFuncSetAttribute(k_sssp_wave<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes_s);
                        k_sssp_wave<true, true, true><<<dim3((w->n_active + wpb - 1) / wpb, w->R), dim3(32 * wpb), bytes_s, st>>>(w->dworlds, (int)ww, (int)sw);
                        return 2;
                    }
                }
                size_t bytes_w = ww * 8 * SSSP_WARPS_PER_BLOCK;
                if (bytes_w > 48 * 1024) cudaFuncSetAttribute(k_sssp_wave<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes_w);
                k_sssp_wave<true, true><<<dim3((w->n_active + SSSP_WARPS_PER_BLOCK - 1) / SSSP_WARPS_PER_BLOCK, w->R), dim3(32 * SSSP_WARPS_PER_BLOCK), bytes_w, st>>>(w->dworlds, (int)ww);
            }
            return 2;
        }
        size_t row_bytes = (size_t)N * 8, per_sm = std::max<size_t>(1, ((size_t)96 << 20) / 148 / std::max<size_t>(row_bytes, 1));
        size_t pad_to = wave_pad >= 0 ? (size_t)wave_pad : (per_sm >= 16 ? 0 : ((size_t)220 * 1024) / per_sm);
        if (pad_to > wbytes) wbytes = std::min<size_t>(pad_to, (size_t)220 * 1024);
        if (wbytes <= 220 * 1024) {
            if (wbytes > 48 * 1024) cudaFuncSetAttribute(k_sssp_wave<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wbytes);
            static const double factor = [] { const char *e = getenv("UXSIM_B200_SSSP_DELTA"); return e ? atof(e) : 2.0; }();  // bucket width / mean edge cost
            k_cost_stats<<<dim3(1, w->R), dim3(256), 0, st>>>(w->dworlds, factor);
            k_sssp_wave<false, false><<<dim3(w->n_active, w->R), dim3(SSSP_WAVE_THREADS), wbytes, st>>>(w->dworlds, 0);
            return 2;
        }
    }
    if ((mode == 2 || (mode == 0 && large)) && N < 65536 && fbytes <= 220 * 1024) {
        if (fbytes > 48 * 1024) cudaFuncSetAttribute(k_sssp_frontier, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fbytes);
        k_sssp_frontier<<<dim3(w->n_active, w->R), dim3(SSSP_THREADS), fbytes, st>>>(w->dworlds);
        return 1;
    }
    if (use_smem && bytes > 48 * 1024) cudaFuncSetAttribute(k_sssp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (use_smem) k_sssp<true><<<dim3(w->n_active, w->R), dim3(SSSP_THREADS), bytes, st>>>(w->dworlds);
    else k_sssp<false><<<dim3(w->n_active, w->R), dim3(SSSP_THREADS), 0, st>>>(w->dworlds);
#endif
    return 1;
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

static int extend_vehicles(ux_world *w) {
    cudaSetDevice(w->device);
    int N = (int)w->nodes.size(), T = w->total_ts, start_ts = w->timestep;
    long long P0 = w->P_uploaded, P = (long long)w->vehs.size();
    if (!w->rep_routes.empty()) return fail(w, UX_ERR_RUNTIME, "adding vehicles to a world with per-replica routes is not supported");
    if (w->record_traversals) return fail(w, UX_ERR_RUNTIME, "adding vehicles after the start is not supported with UX_OPT_RECORD_TRAVERSALS (the traversal log is sized at upload)");
    w->rewindable = false;  // the per-platoon arrays move out of the recorded dynamic region
    w->uniform = false;
#ifndef UX_EMU
    for (auto &g : w->graphs) cudaGraphExecDestroy(g.second.first);  // kernel choices (link-update form, lane groups) depend on the platoon count
    w->graphs.clear();
#endif
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
        d.rep_stride = 0; d.node_rl = 0;  // the regrown per-platoon arrays are separate allocations: the replica-lane node kernel is off for this world
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

