// ux_host_dta.inl -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// C-ABI, part 4: DTA solver batch operations (route-set enumeration, route swap, in-place rewind).

// ---------------------------------------------------------------------------------------------------
// DTA solver batch operations
// ---------------------------------------------------------------------------------------------------

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
    launch_init_dyn(w);  // capacity tokens, signal phases / clocks, node tokens of every replica
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

