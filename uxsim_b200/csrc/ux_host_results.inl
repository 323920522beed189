// ux_host_results.inl -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// C-ABI, part 3: scalar queries, per-platoon logs, array sizes, device-side array production.

static int64_t get_log_array(ux_world *w, int what, int replica, void *dst, int64_t capacity, int *dtype, bool size_only);
static long long dsum_ll(ux_world *w, const long long *first, int n) {
    std::vector<long long> h; fetch_field(w, first, sizeof(long long), n, h);
    long long s = 0; for (auto v : h) s += v; return s;
}
static long long dsum_i(ux_world *w, const int *first, int n) {
    std::vector<int> h; fetch_field(w, first, sizeof(LinkC), n, h);
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
    case UX_I_TRIPS_COMPLETED: return w->uploaded ? dsum_i(w, &LC(w->hw[0], 0).trips, (int)w->links.size()) : 0;
    case UX_I_LOG_ENTRIES: { if (!w->uploaded || !w->p.vehicle_log_mode) return 0; int dt_; return get_log_array(w, UX_A_LOG_T, 0, nullptr, 0, &dt_, true); }
    case UX_I_KERNEL_LAUNCHES: return w->launches;
    case UX_I_NUM_ACTIVE_DESTS: return w->n_active;
    case UX_I_TRANSFER_LEVELS: return w->n_levels;
    case UX_I_LINK_EVENTS: return w->uploaded ? dsum_i(w, &LC(w->hw[0], 0).ev_cnt, (int)w->links.size()) : 0;
    case UX_I_H2D_BYTES: return w->h2d_bytes;
    case UX_I_D2H_BYTES: return w->d2h_bytes;
    case UX_I_RING_SLOTS: return w->C;
    case UX_I_MAX_DEPART_TS: { int m = 0; for (auto &v : w->vehs) m = std::max(m, v.aborted ? v.ts0 : v.ts); return m; }
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
// RUN records per platoon (from the generation / end steps) and their offsets in per-platoon time order
static int log_offsets(ux_world *w, int replica, std::vector<long long> &roff, unsigned long long &nrec, std::vector<int> *gen_out = nullptr,
                       std::vector<int> *st_out = nullptr, std::vector<int> *arr_out = nullptr, std::vector<int> *end_out = nullptr,
                       std::vector<unsigned char> *kind_out = nullptr) {
    const DevWorld &d = w->hw[replica];
    long long P = d.P; int Tnow = w->timestep;
    std::vector<int> gen(P), endt(P), st(P), arr(P); std::vector<unsigned char> kind(P);
    nrec = 0;
    if (P) {
        CUDA_OK(cudaMemcpy(gen.data(), d.v_gen_ts, sizeof(int) * P, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(endt.data(), d.v_end_ts, sizeof(int) * P, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(kind.data(), d.v_end_kind, P, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(st.data(), d.v_state, sizeof(int) * P, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(arr.data(), d.v_arr_ts, sizeof(int) * P, cudaMemcpyDeviceToHost));
    }
    CUDA_OK(cudaMemcpy(&nrec, d.lg_count, sizeof(nrec), cudaMemcpyDeviceToHost));
    roff.assign(P + 1, 0);
    for (long long i = 0; i < P; i++) {
        long long r = 0;
        if (gen[i] >= 0) { int last = kind[i] == 1 ? endt[i] : kind[i] == 2 ? endt[i] - 1 : Tnow - 1; r = last - gen[i] + 1; if (r < 0) r = 0; }
        roff[i + 1] = roff[i] + r;
    }
    if ((unsigned long long)roff[P] != nrec) return fail(w, UX_ERR_RUNTIME, "vehicle log bookkeeping mismatch: %lld expected, %llu recorded", roff[P], nrec);
    if (gen_out) gen_out->swap(gen);
    if (st_out) st_out->swap(st);
    if (arr_out) arr_out->swap(arr);
    if (end_out) end_out->swap(endt);
    if (kind_out) kind_out->swap(kind);
    return 0;
}

static int build_logs(ux_world *w, int replica) {
    auto &G = w->logs;
    if (G.replica == replica && G.timestep == w->timestep) return 0;
    const DevWorld &d = w->hw[replica];
    long long P = d.P; int Tnow = w->timestep; double dt = w->dt;
    std::vector<int> gen, st, arr, endt; std::vector<unsigned char> kind;
    std::vector<long long> roff;
    unsigned long long nrec = 0;
    { int rc = log_offsets(w, replica, roff, nrec, &gen, &st, &arr, &endt, &kind); if (rc) return rc; }
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
    case UX_A_VEH_TRAVEL_TIME: case UX_A_VEH_DISTANCE: case UX_A_VEH_X: case UX_A_VEH_V: case UX_A_VEH_LINK_ARRIVAL_TIME: n = P; break;
    case UX_A_SIGNAL_LOG: n = N * T; dt = UX_DT_I32; break;
    case UX_A_NODE_SIGNAL_PHASE: n = N; dt = UX_DT_I32; break;
    case UX_A_NODE_SIGNAL_T: case UX_A_NODE_FLOW_CAPACITY: n = N; break;
    case UX_A_NODE_NUM_LANES: n = N; dt = UX_DT_I32; break;
    case UX_A_LINK_NUM_VEHICLES: n = L; dt = UX_DT_I32; break;
    case UX_A_LINK_CAPACITY_IN_REMAIN: case UX_A_LINK_CAPACITY_OUT_REMAIN: case UX_A_LINK_STAT_COUNT: case UX_A_LINK_TRIPS_COMPLETED: n = L; break;
    case UX_A_ROUTE_PREF: n = N * L; break;
    case UX_A_ROUTE_DIST: n = N * N; break;
    case UX_A_ROUTE_NEXT: n = N * N; dt = UX_DT_I32; break;
    case UX_A_EDIE_OFFSETS: n = (int64_t)w->edie.off.size(); dt = UX_DT_I64; break;
    case UX_A_EDIE_NX: n = (int64_t)w->edie.nx.size(); dt = UX_DT_I32; break;
    case UX_A_EDIE_TN: case UX_A_EDIE_DN: n = (int64_t)w->edie.tn.size(); break;
    case UX_A_OD_ORIG: case UX_A_OD_DEST: n = (int64_t)w->summary.o.size(); dt = UX_DT_I32; break;
    case UX_A_OD_STATS: n = (int64_t)w->summary.od.size(); break;
    case UX_A_LINKC_STATS: n = (int64_t)w->summary.link.size(); break;
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
        if (tot > 0) UX_LAUNCH(k_transpose<double>, dim3((unsigned)((tot + 255) / 256)), dim3(256), st, src, w->d_scratch, rows, cols, d.RS);
        *dptr = w->d_scratch; return 0; };
    auto transpose_i = [&](const int *src, int rows, int cols) -> int {
        int rc = ensure_scratch(w, sizeof(int) * (size_t)rows * cols + 16); if (rc) return rc;
        long long tot = (long long)rows * cols;
        if (tot > 0) UX_LAUNCH(k_transpose<int>, dim3((unsigned)((tot + 255) / 256)), dim3(256), st, src, (int *)w->d_scratch, rows, cols, d.RS);
        *dptr = w->d_scratch; return 0; };
    int rc = 0;
    switch (what) {
#ifdef UX_TIME_MAJOR
    case UX_A_CUM_ARRIVAL: rc = transpose_d(d.cum_arr, T, L); break;
    case UX_A_CUM_DEPARTURE: rc = transpose_d(d.cum_dep, T, L); break;
    case UX_A_TRAVELTIME_INSTANT: rc = transpose_d(d.tt_inst, T, L); break;
    case UX_A_SIGNAL_LOG: rc = transpose_i(d.sig_log, T, N); break;
#else  // the series are kept link-major on the device: the ABI layout, no transpose
    case UX_A_CUM_ARRIVAL: *dptr = d.cum_arr; break;
    case UX_A_CUM_DEPARTURE: *dptr = d.cum_dep; break;
    case UX_A_TRAVELTIME_INSTANT: *dptr = d.tt_inst; break;
    case UX_A_SIGNAL_LOG: *dptr = d.sig_log; break;
#endif
    case UX_A_TRAVELTIME_ACTUAL:
        rc = ensure_scratch(w, sizeof(double) * (size_t)T * L + 16); if (rc) break;
        if (L > 0) UX_LAUNCH(k_tt_actual, dim3((L + 127) / 128), dim3(128), st, w->dworlds, replica, w->d_scratch);
        *dptr = w->d_scratch; break;
    case UX_A_VEH_ARRIVAL_TS: *dptr = d.v_arr_ts; break;
    case UX_A_VEH_DEPART_TS: *dptr = (void *)d.v_depart_ts; break;
    case UX_A_VEH_TRAVEL_TIME: *dptr = d.v_tt; break;
    case UX_A_VEH_LINK_ARRIVAL_TIME: *dptr = d.v_atl; break;
    case UX_A_VEH_ORIG: *dptr = (void *)d.v_orig; break;
    case UX_A_VEH_DEST: *dptr = (void *)d.v_dest; break;
    case UX_A_VEH_LINK: *dptr = d.v_link; break;
    case UX_A_NODE_SIGNAL_PHASE: *dptr = gather_device<int>(w, &NS(d, 0).sig_phase, sizeof(NodeS), N); break;
    case UX_A_NODE_SIGNAL_T: *dptr = gather_device<double>(w, &NR(d, 0).sig_t, sizeof(NodeR), N); break;
    case UX_A_LINK_CAPACITY_IN_REMAIN: *dptr = gather_device<double>(w, &LR(d, 0).cap_in_rem, sizeof(LinkR), L); break;
    case UX_A_LINK_CAPACITY_OUT_REMAIN: *dptr = gather_device<double>(w, &LR(d, 0).cap_out_rem, sizeof(LinkR), L); break;
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


