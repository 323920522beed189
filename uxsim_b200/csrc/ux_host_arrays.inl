// ux_host_arrays.inl -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// C-ABI, part 5: get_array (host copies of results), ensemble link statistics, last_error.

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

// Analyzer.compute_edie_state (analyzer.py:205-330) on the device, from the RUN records of `replica`
UX_API int ux_edie_state(ux_world *w, int replica, const double *edie_dt, const double *edie_dx) {
    if (!w->uploaded || w->timestep == 0) return fail(w, UX_ERR_RUNTIME, "edie_state needs a run");
    if (replica < 0 || replica >= w->R) return fail(w, UX_ERR_OUT_OF_RANGE, "replica %d out of range", replica);
    if (!w->p.vehicle_log_mode) return fail(w, UX_ERR_RUNTIME, "edie_state needs the per-step vehicle logs (vehicle_log_mode)");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    int L = (int)w->links.size();
    auto &R = w->edie;
    R.replica = -1;
    R.off.assign(L + 1, 0); R.nx.assign(L, 0);
    for (int l = 0; l < L; l++) {
        if (!(edie_dt[l] > 0) || !(edie_dx[l] > 0)) return fail(w, UX_ERR_INVALID, "edie_state: cell sizes must be positive");
        int nt = (int)(w->p.t_max / edie_dt[l]), nx = (int)(w->links[l].length / edie_dx[l]);
        R.nx[l] = nx; R.off[l + 1] = R.off[l] + (long long)nt * nx;
    }
    long long cells = R.off[L];
    std::vector<long long> roff;
    unsigned long long nrec = 0;
    { int rc = log_offsets(w, replica, roff, nrec); if (rc) return rc; }
    long long P = (long long)roff.size() - 1;
    size_t bytes = 0;
    auto carve = [&](size_t b) { size_t o = bytes; bytes += (b + 255) & ~(size_t)255; return o; };
    size_t o_off = carve(8 * (size_t)(P + 1)), o_vid = carve(4 * (size_t)nrec), o_lnk = carve(4 * (size_t)nrec), o_x = carve(8 * (size_t)nrec), o_coff = carve(8 * (size_t)(L + 1)),
           o_cnx = carve(4 * (size_t)L), o_dt = carve(8 * (size_t)L), o_dx = carve(8 * (size_t)L), o_tn = carve(8 * (size_t)cells), o_dn = carve(8 * (size_t)cells),
           o_otn = carve(8 * (size_t)cells), o_odn = carve(8 * (size_t)cells);
    char *buf = nullptr;
    CUDA_OK(cudaMalloc((void **)&buf, bytes + 256));
    cudaMemcpyAsync(buf + o_off, roff.data(), 8 * (size_t)(P + 1), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(buf + o_coff, R.off.data(), 8 * (size_t)(L + 1), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(buf + o_cnx, R.nx.data(), 4 * (size_t)L, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(buf + o_dt, edie_dt, 8 * (size_t)L, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(buf + o_dx, edie_dx, 8 * (size_t)L, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(buf + o_tn, 0, o_otn - o_tn, st);
    if (nrec) {
        UX_LAUNCH(k_edie_order, dim3((unsigned)((nrec + 255) / 256)), dim3(256), st, w->dworlds, replica, (long long)nrec, (const long long *)(buf + o_off), (int *)(buf + o_vid),
                  (int *)(buf + o_lnk), (double *)(buf + o_x));
        UX_LAUNCH(k_edie_accumulate, dim3((unsigned)((nrec + 255) / 256)), dim3(256), st, w->dworlds, replica, (long long)nrec, (const long long *)(buf + o_off),
                  (const int *)(buf + o_vid), (const int *)(buf + o_lnk), (const double *)(buf + o_x), (const long long *)(buf + o_coff), (const int *)(buf + o_cnx),
                  (const double *)(buf + o_dt), (const double *)(buf + o_dx), (unsigned long long *)(buf + o_tn), (unsigned long long *)(buf + o_dn));
    }
    R.tn.assign((size_t)cells, 0.0); R.dn.assign((size_t)cells, 0.0);
    if (cells) {
        UX_LAUNCH(k_edie_finish, dim3((unsigned)((cells + 255) / 256)), dim3(256), st, cells, (const unsigned long long *)(buf + o_tn), (const unsigned long long *)(buf + o_dn),
                  w->p.delta_n, (double *)(buf + o_otn), (double *)(buf + o_odn));
        cudaMemcpyAsync(R.tn.data(), buf + o_otn, 8 * (size_t)cells, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(R.dn.data(), buf + o_odn, 8 * (size_t)cells, cudaMemcpyDeviceToHost, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(w, UX_ERR_CUDA, "CUDA error in edie_state: %s", cudaGetErrorString(e));
    w->d2h_bytes += 16 * cells; w->h2d_bytes += 8 * (P + 1) + 28 * (long long)L;
    R.replica = replica;
    return 0;
}

static int64_t get_edie_array(ux_world *w, int what, int replica, void *dst, int64_t capacity) {
    auto &R = w->edie;
    if (R.replica != replica) return fail(w, UX_ERR_RUNTIME, "no Edie state for replica %d", replica);
    int64_t n = what == UX_A_EDIE_OFFSETS ? (int64_t)R.off.size() : what == UX_A_EDIE_NX ? (int64_t)R.nx.size() : (int64_t)R.tn.size();
    if (capacity < n) return fail(w, UX_ERR_INVALID, "get_array: capacity %lld < %lld", (long long)capacity, (long long)n);
    const void *src = what == UX_A_EDIE_OFFSETS ? (const void *)R.off.data() : what == UX_A_EDIE_NX ? (const void *)R.nx.data() : what == UX_A_EDIE_TN ? (const void *)R.tn.data() : (const void *)R.dn.data();
    if (n > 0) memcpy(dst, src, (what == UX_A_EDIE_NX ? 4 : 8) * (size_t)n);
    return n;
}

// Analyzer.od_analysis / link_analysis_coarse (analyzer.py:121-203) of `replica` on the device
UX_API int ux_analysis_summary(ux_world *w, int replica) {
    if (!w->uploaded || w->timestep == 0) return fail(w, UX_ERR_RUNTIME, "analysis_summary needs a run");
    if (replica < 0 || replica >= w->R) return fail(w, UX_ERR_OUT_OF_RANGE, "replica %d out of range", replica);
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    auto &S = w->summary;
    S.replica = -1;
    const int L = (int)w->links.size(), N = (int)w->nodes.size();
    const long long P = (long long)w->vehs.size();
    // the OD pairs of the demand, ascending in (orig * N + dest), each with its platoons in id order (static: host side)
    std::vector<int> order((size_t)P);
    for (long long i = 0; i < P; i++) order[(size_t)i] = (int)i;
    auto key = [&](int v) { return (long long)w->vehs[v].orig * N + w->vehs[v].dest; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
    std::vector<long long> off;
    S.o.clear(); S.d.clear();
    for (long long q = 0; q < P; q++) {
        if (q == 0 || key(order[(size_t)q]) != key(order[(size_t)q - 1])) { off.push_back(q); S.o.push_back(w->vehs[order[(size_t)q]].orig); S.d.push_back(w->vehs[order[(size_t)q]].dest); }
    }
    off.push_back(P);
    const int K = (int)S.o.size();
    void *px = nullptr;
    if (P > 0) { int64_t m = device_array(w, UX_A_VEH_X, replica, &px, nullptr); if (m < 0) return (int)m; }
    size_t bytes = 0;
    auto carve = [&](size_t b) { size_t o = bytes; bytes += (b + 255) & ~(size_t)255; return o; };
    size_t o_off = carve(8 * (size_t)(K + 1)), o_veh = carve(4 * (size_t)P), o_od = carve(48 * (size_t)K), o_lk = carve(32 * (size_t)L);
    char *buf = nullptr;
    CUDA_OK(cudaMalloc((void **)&buf, bytes + 256));
    cudaMemcpyAsync(buf + o_off, off.data(), 8 * (size_t)(K + 1), cudaMemcpyHostToDevice, st);
    if (P) cudaMemcpyAsync(buf + o_veh, order.data(), 4 * (size_t)P, cudaMemcpyHostToDevice, st);
    S.od.assign((size_t)K * 6, 0.0); S.link.assign((size_t)L * 4, 0.0);
    if (K) {
        UX_LAUNCH(k_od_stats, dim3((unsigned)((K + 127) / 128)), dim3(128), st, w->dworlds, replica, K, (const long long *)(buf + o_off), (const int *)(buf + o_veh),
                  (const double *)px, (double *)(buf + o_od));
        cudaMemcpyAsync(S.od.data(), buf + o_od, 48 * (size_t)K, cudaMemcpyDeviceToHost, st);
    }
    if (L) {
        UX_LAUNCH(k_linkc_stats, dim3((unsigned)((L + 127) / 128)), dim3(128), st, w->dworlds, replica, (double *)(buf + o_lk));
        cudaMemcpyAsync(S.link.data(), buf + o_lk, 32 * (size_t)L, cudaMemcpyDeviceToHost, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(w, UX_ERR_CUDA, "CUDA error in analysis_summary: %s", cudaGetErrorString(e));
    w->d2h_bytes += 48ll * K + 32ll * L; w->h2d_bytes += 8ll * (K + 1) + 4 * P;
    S.replica = replica;
    return 0;
}

static int64_t get_summary_array(ux_world *w, int what, int replica, void *dst, int64_t capacity) {
    auto &S = w->summary;
    if (S.replica != replica) return fail(w, UX_ERR_RUNTIME, "no analysis summary for replica %d", replica);
    const void *src = what == UX_A_OD_ORIG ? (const void *)S.o.data() : what == UX_A_OD_DEST ? (const void *)S.d.data() : what == UX_A_OD_STATS ? (const void *)S.od.data() : (const void *)S.link.data();
    int64_t n = what == UX_A_OD_ORIG || what == UX_A_OD_DEST ? (int64_t)S.o.size() : what == UX_A_OD_STATS ? (int64_t)S.od.size() : (int64_t)S.link.size();
    if (capacity < n) return fail(w, UX_ERR_INVALID, "get_array: capacity %lld < %lld", (long long)capacity, (long long)n);
    if (n > 0) memcpy(dst, src, (what == UX_A_OD_ORIG || what == UX_A_OD_DEST ? 4 : 8) * (size_t)n);
    return n;
}

UX_API int64_t ux_get_array(ux_world *w, int what, int replica, void *dst, int64_t capacity) {
    if (is_log_array(what)) return get_log_array(w, what, replica, dst, capacity, nullptr, false);
    if (what >= UX_A_OD_ORIG && what <= UX_A_LINKC_STATS) return get_summary_array(w, what, replica, dst, capacity);
    if (what >= UX_A_DTA_OD_ORIG && what <= UX_A_DTA_RA_ENTRY_T) return get_dta_array(w, what, replica, dst, capacity);
    if (what >= UX_A_EDIE_OFFSETS && what <= UX_A_EDIE_DN) return get_edie_array(w, what, replica, dst, capacity);
    if (what == UX_A_NODE_FLOW_CAPACITY || what == UX_A_NODE_NUM_LANES) {  // static after initialize_adj_matrix: answered from the host copy
        int64_t N = (int64_t)w->nodes.size();
        if (capacity < N) return fail(w, UX_ERR_INVALID, "get_array: capacity %lld < %lld", (long long)capacity, (long long)N);
        for (int64_t i = 0; i < N; i++) { if (what == UX_A_NODE_FLOW_CAPACITY) ((double *)dst)[i] = w->nodes[i].fc; else ((int *)dst)[i] = w->nodes[i].lanes; }
        return N;
    }
    if (what == UX_A_VEH_ORIG || what == UX_A_VEH_DEST || what == UX_A_VEH_DEPART_TS) {  // static: answered from the host copy (no upload forced)
        int64_t P = (int64_t)w->vehs.size();
        if (replica < 0 || replica >= w->R) return fail(w, UX_ERR_OUT_OF_RANGE, "replica %d out of range", replica);
        if (capacity < P) return fail(w, UX_ERR_INVALID, "get_array: capacity %lld < %lld", (long long)capacity, (long long)P);
        int *oi = (int *)dst;
        for (int64_t i = 0; i < P; i++) oi[i] = what == UX_A_VEH_ORIG ? w->vehs[i].orig : what == UX_A_VEH_DEST ? w->vehs[i].dest : (w->vehs[i].aborted ? w->vehs[i].ts0 : w->vehs[i].ts);
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
        {   // replica-minor members: one gather kernel writes the replica-major [R, n] block
            const DevWorld &d0 = w->hw[0];
            const void *first = nullptr; size_t rec = 0;
            switch (what) {
            case UX_A_NODE_SIGNAL_PHASE: first = &NS(d0, 0).sig_phase; rec = sizeof(NodeS); break;
            case UX_A_NODE_SIGNAL_T: first = &NR(d0, 0).sig_t; rec = sizeof(NodeR); break;
            case UX_A_LINK_CAPACITY_IN_REMAIN: first = &LR(d0, 0).cap_in_rem; rec = sizeof(LinkR); break;
            case UX_A_LINK_CAPACITY_OUT_REMAIN: first = &LR(d0, 0).cap_out_rem; rec = sizeof(LinkR); break;
            case UX_A_LINK_STAT_COUNT: first = &LSTAT(d0, 0); rec = sizeof(long long); break;
            default: break;
            }
            if (first) {
                int rc = ensure_scratch(w, row * R + 16); if (rc) return rc;
                long long tot = (long long)n0 * R;
                if (tot > 0) {
                    if (es0 == 4) UX_LAUNCH(k_gather_all<int>, dim3((unsigned)((tot + 255) / 256)), dim3(256), w->stream, (const char *)first, rec, d0.RS, R, (long long)n0, (int *)w->d_scratch);
                    else UX_LAUNCH(k_gather_all<long long>, dim3((unsigned)((tot + 255) / 256)), dim3(256), w->stream, (const char *)first, rec, d0.RS, R, (long long)n0, (long long *)w->d_scratch);
                    CUDA_OK(cudaMemcpyAsync(dst, w->d_scratch, row * R, cudaMemcpyDeviceToHost, w->stream));
                }
                CUDA_OK(cudaStreamSynchronize(w->stream));
                if (what == UX_A_LINK_STAT_COUNT) { long long *hi = (long long *)dst; double *o = (double *)dst; for (long long i = 0; i < tot; i++) o[i] = (double)hi[i]; }  // i64 on the device, f64 in the ABI
                w->d2h_bytes += (long long)(row * R);
                return n0 * R;
            }
        }
        std::vector<const void *> src(R, nullptr);
        for (int r = 0; r < R; r++) {
            const DevWorld &d = w->hw[r];
            switch (what) {
            case UX_A_VEH_ARRIVAL_TS: src[r] = d.v_arr_ts; break;
            case UX_A_VEH_TRAVEL_TIME: src[r] = d.v_tt; break;
            case UX_A_VEH_LINK: src[r] = d.v_link; break;
            default: break;
            }
        }
        {
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
            if (w->vehs[i].aborted) s = UX_VS_ABORT;
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
        std::vector<int> hd, tl;
        if (!fetch_field(w, &LQ(d, 0).head[w->timestep & 1], sizeof(LinkQ), L, hd) || !fetch_field(w, &LQ(d, 0).tail, sizeof(LinkQ), L, tl)) return fail(w, UX_ERR_CUDA, "device read failed");
        for (int l = 0; l < L; l++) oi[l] = tl[l] - hd[l];
        break; }
    case UX_A_LINK_STAT_COUNT: { std::vector<long long> h; if (!fetch_field(w, &LSTAT(d, 0), sizeof(long long), L, h)) return fail(w, UX_ERR_CUDA, "device read failed"); for (int l = 0; l < L; l++) o[l] = (double)h[l]; break; }
    case UX_A_LINK_TRIPS_COMPLETED: { std::vector<int> h; if (!fetch_field(w, &LC(d, 0).trips, sizeof(LinkC), L, h)) return fail(w, UX_ERR_CUDA, "device read failed"); for (int l = 0; l < L; l++) o[l] = (double)h[l]; break; }
    case UX_A_ROUTE_PREF: {
        std::vector<double> h((size_t)d.D * L); if (!h.empty()) CUDA_OK(cudaMemcpy(h.data(), d.pref, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
        for (int k = 0; k < N; k++) { int r = w->dest_row[k]; for (int l = 0; l < L; l++) o[(size_t)k * L + l] = r >= 0 ? h[(size_t)r * L + l] : 0.0; }
        break; }
    case UX_A_ROUTE_DIST: case UX_A_ROUTE_NEXT: {
        // The loop keeps the shortest-path rows of the destinations that occur in the demand only; the reference's
        // RouteChoice.dist / .next are all-pairs tables (traffi.cpp:1317-1385).  When rows are missing and a search has happened, the
        // getter runs the same search once more for EVERY node as destination, on the edge costs of the last route update (they are still
        // on the device, noise included), into scratch memory -- same fixed point, so the rows the loop holds come out identical.
        const bool searched = w->timestep > 0;
        // UX_OPT_ROUTE_TABLES_AT: the tables of an earlier route update, searched on that update's edge costs
        int hist_u = -1;
        if (searched && w->route_tables_at >= 0 && d.cost_hist) {
            int last_u = (w->timestep - 1) / w->route_ts, u = std::min(w->route_tables_at / w->route_ts, last_u);
            if (u < last_u) hist_u = u;
        }
        const bool full = searched && (d.D < N || hist_u >= 0) && N > 0 && w->n_edges > 0;
        std::vector<double> hd; std::vector<int> hn;
        int D_ = d.D;
        std::vector<int> row_of = w->dest_row;
        if (full) {
            size_t bytes = sizeof(DevWorld) + (size_t)N * N * 12 + (size_t)N * 8 + 256;
            int rc = ensure_scratch(w, bytes); if (rc) return rc;
            char *base = (char *)w->d_scratch;
            DevWorld *dW = (DevWorld *)base;
            double *s_dist = (double *)(base + ((sizeof(DevWorld) + 255) & ~(size_t)255));
            int *s_next = (int *)(s_dist + (size_t)N * N), *s_rowdest = s_next + (size_t)N * N, *s_has = s_rowdest + N;
            DevWorld t = d;
            t.D = N; t.row_dest = s_rowdest; t.rdist = s_dist; t.rnext = s_next; t.has_path = s_has;
            if (hist_u >= 0) t.pair_cost = d.cost_hist + (size_t)hist_u * (size_t)w->n_edges;
            std::vector<int> ident(N); for (int i = 0; i < N; i++) ident[i] = i;
            CUDA_OK(cudaMemcpyAsync(dW, &t, sizeof(DevWorld), cudaMemcpyHostToDevice, w->stream));
            CUDA_OK(cudaMemcpyAsync(s_rowdest, ident.data(), sizeof(int) * N, cudaMemcpyHostToDevice, w->stream));
#ifdef UX_EMU
            UX_LAUNCH(k_sssp<false>, dim3(N, 1), dim3(SSSP_THREADS), w->stream, (const DevWorld *)dW);
#else
            {   // labels in shared memory when a row fits (as launch_sssp does), else in global memory read through L2
                size_t sbytes = (size_t)N * 12 + 16;
                if (sbytes <= 200 * 1024) {
                    if (sbytes > 48 * 1024) cudaFuncSetAttribute(k_sssp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbytes);
                    k_sssp<true><<<dim3(N, 1), dim3(SSSP_THREADS), sbytes, w->stream>>>((const DevWorld *)dW);
                } else k_sssp<false><<<dim3(N, 1), dim3(SSSP_THREADS), 0, w->stream>>>((const DevWorld *)dW);
            }
#endif
            D_ = N; for (int i = 0; i < N; i++) row_of[i] = i;
            if (what == UX_A_ROUTE_DIST) { hd.resize((size_t)N * N); CUDA_OK(cudaMemcpyAsync(hd.data(), s_dist, sizeof(double) * hd.size(), cudaMemcpyDeviceToHost, w->stream)); }
            else { hn.resize((size_t)N * N); CUDA_OK(cudaMemcpyAsync(hn.data(), s_next, sizeof(int) * hn.size(), cudaMemcpyDeviceToHost, w->stream)); }
            CUDA_OK(cudaStreamSynchronize(w->stream));
        } else if (what == UX_A_ROUTE_DIST) { hd.resize((size_t)d.D * N); if (!hd.empty()) CUDA_OK(cudaMemcpy(hd.data(), d.rdist, sizeof(double) * hd.size(), cudaMemcpyDeviceToHost)); }
        else { hn.resize((size_t)d.D * N); if (!hn.empty()) CUDA_OK(cudaMemcpy(hn.data(), d.rnext, sizeof(int) * hn.size(), cudaMemcpyDeviceToHost)); }
        (void)D_;
        if (what == UX_A_ROUTE_DIST) { for (int i = 0; i < N; i++) for (int k = 0; k < N; k++) { int r = row_of[k]; o[(size_t)i * N + k] = (r >= 0 && searched) ? hd[(size_t)r * N + i] : INFINITY; } }
        else { for (int i = 0; i < N; i++) for (int k = 0; k < N; k++) { int r = row_of[k]; oi[(size_t)i * N + k] = r >= 0 ? hn[(size_t)r * N + i] : -1; } }
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
#ifdef UX_EMU
        UX_LAUNCH(k_ensemble_partial, dim3((L + 63) / 64, w->R), dim3(64), w->stream, w->dworlds, w->d_scratch, w->R);
#else
        UX_LAUNCH(k_ensemble_partial, dim3((unsigned)(((size_t)L * 32 + 255) / 256), (w->R + 31) / 32), dim3(256), w->stream, w->dworlds, w->d_scratch, w->R);
#endif
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

