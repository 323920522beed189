// ux_host_run.inl -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// C-ABI, part 2: the executor (CUDA graphs per chunk shape), ux_main_loop, mid-run setters.

// enqueue the kernels of `nsteps` timesteps starting at phase (t mod route_ts) = `phase` on the stream
static void enqueue_steps(ux_world *w, int nsteps, int phase, long long *launch_count) {
    int N = (int)w->nodes.size(), R = w->R;
    cudaStream_t st = w->stream;
    dim3 g_nodes(R, (N + XFER_WARPS - 1) / XFER_WARPS), g_update((w->nseg + UPD_GROUPS_PER_BLOCK(w->upd_group) - 1) / UPD_GROUPS_PER_BLOCK(w->upd_group), R);
    long long lc = 0;
    // Large ensembles of lightly loaded networks run the link update with one thread per (link, replica) (k_update_rl: 4x fewer
    // warp instructions; B200, Chicago sketch: 102 -> 77 us at 64 replicas, 198 -> 116 us at 128); below 64 replicas, or with
    // long queues (more than 32 platoons per link on average), the warp-per-link kernel wins.  UXSIM_B200_UPDATE=rl|warp forces one.
    const char *upd_env = getenv("UXSIM_B200_UPDATE");
    bool update_rl = upd_env ? !strcmp(upd_env, "rl") : (R >= 64 && (long long)w->vehs.size() <= 32ll * (long long)w->links.size());
    for (int s = 0; s < nsteps; s++) {
        // Ensembles (from 32 replicas): the node step in replica-lane form -- k_link_pre<true> (warp = one link / node in 32 replicas,
        // incl. the mean-speed travel times) + k_node_rl (thread per node and replica).  Otherwise warp-per-node k_node, with the scalar
        // bookkeeping peeled into k_link_pre<false> for large worlds.
        const bool node_rl = w->node_rl && w->uniform;
        if (node_rl) { KT_BEGIN(UX_K_PRE) UX_LAUNCH(k_link_pre<true>, dim3((((unsigned)w->links.size() + (unsigned)N) * 32u + 255u) / 256u, (R + 31) / 32), dim3(256), st, w->dworlds, s, R); KT_END(UX_K_PRE) lc++; }
        else if (w->pre_split && R >= 8) { KT_BEGIN(UX_K_PRE) UX_LAUNCH(k_link_pre<true>, dim3((((unsigned)w->links.size() + (unsigned)N) * 32u + 255u) / 256u, (R + 31) / 32), dim3(256), st, w->dworlds, s, R); KT_END(UX_K_PRE) lc++; }  // (the mean speeds stay in k_node: W.node_rl == 0)
        else if (w->pre_split) { KT_BEGIN(UX_K_PRE) UX_LAUNCH(k_link_pre<false>, dim3(((unsigned)w->links.size() + (unsigned)N + 255) / 256, R), dim3(256), st, w->dworlds, s, R); KT_END(UX_K_PRE) lc++; }
        if (N > 0 && node_rl) {
            KT_BEGIN(UX_K_TRANSFER)
            dim3 g_nh(w->n_node_blks);
            const bool heavy = w->node_heavy_lanes < w->max_node_lanes;  // some node is stepped by a warp
#ifndef UX_EMU
            if (s == 0 && !heavy) {  // pure replica-lane form: no shared-memory staging, all of the L1 for the threads' local arrays
                cudaFuncSetAttribute(k_node_h<8, 8, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
                cudaFuncSetAttribute(k_node_h<32, 16, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
                cudaFuncSetAttribute(k_node_h<UX_MAXC, UX_MAXDEG, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
            }
#endif
#define UX_NODE_H(MC_, MD_) { if (heavy) UX_LAUNCH((k_node_h<MC_, MD_, true>), g_nh, dim3(32 * NRL_WARPS), st, w->dworlds, s, R, (const int4 *)w->d_node_blks); \
                              else UX_LAUNCH((k_node_h<MC_, MD_, false>), g_nh, dim3(32 * NRL_WARPS), st, w->dworlds, s, R, (const int4 *)w->d_node_blks); }
            if (w->max_node_lanes <= 8 && w->max_node_deg <= 8) UX_NODE_H(8, 8)
            else if (w->max_node_lanes <= 32) UX_NODE_H(32, 16)
            else UX_NODE_H(UX_MAXC, UX_MAXDEG)
#undef UX_NODE_H
            KT_END(UX_K_TRANSFER) lc++;
        } else if (N > 0) {
            KT_BEGIN(UX_K_TRANSFER)
            // one wave of blocks at two per SM (a single scenario of Chicago-sketch size): the low-residency build, no register cap to spill against
#ifdef UX_EMU
            static const int n_sm = 148;
#else
            static const int n_sm = [] { int dev = 0, n = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }();
#endif
            const bool one_wave = (long long)g_nodes.x * g_nodes.y <= 2ll * n_sm;
#define UX_NODE_W(MC_, MD_) { if (one_wave) UX_LAUNCH((k_node<MC_, MD_, 2>), g_nodes, dim3(32 * XFER_WARPS), st, w->dworlds, s); \
                              else UX_LAUNCH((k_node<MC_, MD_>), g_nodes, dim3(32 * XFER_WARPS), st, w->dworlds, s); }
            if (w->max_node_lanes <= 8 && w->max_node_deg <= 8) UX_NODE_W(8, 8)
            else if (w->max_node_lanes <= 32) UX_NODE_W(32, 16)
            else UX_NODE_W(UX_MAXC, UX_MAXDEG)
#undef UX_NODE_W
            KT_END(UX_K_TRANSFER) lc++;
        }
        if (w->nseg > 0) {
            KT_BEGIN(UX_K_UPDATE)
            if (update_rl) {
                // helper blocks (grid z) share long queues; the head-of-link log records need the whole queue in one block
                // Only where queues get long: on a lightly loaded network (the regime in which this kernel is picked by itself) no link has
                // 128 platoons over 32 replicas, and helper blocks that look and leave still cost 16 % of the launch (B200, Chicago x2048: 969 ->
                // 814 us without them; forced on Chicago deltan=5 x16 they take it from 143 to 56 us).
                static const int rl_parts_env = [] { const char *e = getenv("UXSIM_B200_RL_PARTS"); return e ? atoi(e) : 0; }();
                const bool rl_full = ((long long)w->links.size() / UPD_RL_WARPS) * ((R + 31) / 32) >= 2048;  // enough blocks to fill the GPU without splitting queues (x128: 98 us with helpers, 107 without)
                const int rl_parts = rl_parts_env > 0 ? rl_parts_env : ((rl_full && (long long)w->vehs.size() <= 32ll * (long long)w->links.size()) ? 1 : 4);
                dim3 g_rl(((int)w->links.size() + UPD_RL_WARPS - 1) / UPD_RL_WARPS, (R + 31) / 32, w->p.vehicle_log_mode ? 1 : std::max(1, rl_parts));
#ifdef UX_EMU
                UX_LAUNCH(k_update_rl, g_rl, dim3(32 * UPD_RL_WARPS), st, w->dworlds, s, R, (int)w->links.size());
#else
                const size_t rl_bytes = 32 * UPD_RL_STRIDE * 8;  // the replicas' DevWorld structs
                if (s == 0) cudaFuncSetAttribute(k_update_rl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rl_bytes);  // per device
                k_update_rl<<<g_rl, dim3(32 * UPD_RL_WARPS), rl_bytes, st>>>(w->dworlds, s, R, (int)w->links.size());
#endif
            }
            else if (w->upd_group == 16) UX_LAUNCH(k_update<16>, g_update, dim3(32 * UPD_WARPS), st, w->dworlds, s);
            else UX_LAUNCH(k_update<32>, g_update, dim3(32 * UPD_WARPS), st, w->dworlds, s);
            KT_END(UX_K_UPDATE) lc++;
        }
        bool route = ((phase + s) % w->route_ts) == 0;
        if (w->n_active > 0 && w->n_edges > 0) {
            if (route) {
                KT_BEGIN(UX_K_EDGE_COST) UX_LAUNCH(k_edge_cost, dim3((w->n_edges + 127) / 128, R), dim3(128), st, w->dworlds, s); KT_END(UX_K_EDGE_COST) lc++;
                KT_BEGIN(UX_K_SSSP) lc += launch_sssp(w, st); KT_END(UX_K_SSSP)
            }
            if (route || w->p.route_choice_update_gradual) {
                KT_BEGIN(UX_K_DUO) UX_LAUNCH(k_duo, dim3((unsigned)((w->links.size() + 128 * DUO_PER_THREAD - 1) / (128 * DUO_PER_THREAD)), w->n_active, R), dim3(128), st, w->dworlds, w->p.route_choice_update_gradual ? 1 : 0); KT_END(UX_K_DUO) lc++;
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
    case DERR_DEP_TIMEOUT: return "transfer of node %d timed out waiting for a lower dependency level";
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
    if (w->sig_tables_dirty) { int rc = upload_signal_tables(w); if (rc) return rc; w->sig_tables_dirty = false; }
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
            int code = (c0 == DERR_RING_FULL || c0 == DERR_LOG_FULL) ? UX_ERR_CAPACITY : c0 == DERR_ROUTE_EXHAUSTED ? UX_ERR_OUT_OF_RANGE : c0 == DERR_DEP_TIMEOUT ? UX_ERR_RUNTIME : UX_ERR_INVALID;
            return fail(w, code, derr_text(c0), e[2 * r + 1]);
        }
    }
#ifdef UX_PROFILE_NODE
    {   // measurement build: per-step maxima of the node-step phases (cycles), averaged over the steps of this call
        static unsigned long long h[4096][8], h2[4096][8];
        cudaMemcpyFromSymbol(h, g_node_prof, sizeof(h)); cudaMemcpyFromSymbol(h2, g_node_prof2, sizeof(h2));
        const char *nm_rl[8] = {"generate", "gather", "dep-wait", "out-links", "slot loop", "flush+writeback", "total(active)", "-"};
        const char *nm_w[8] = {"xfer gather A1-A3", "xfer dep-wait", "xfer A4+window", "xfer rng+shuffle", "xfer rounds", "xfer flush+wb", "xfer total", "kernel total/warp"};
        const char *nm_p[8] = {"pre fresh tt", "pre scalar", "pre release", "gen choice+gather", "gen window", "gen serial", "gen wb+tail", "pre total"};
        for (int tab = 0; tab < 2; tab++) {
            double sum[8] = {0}, mx[8] = {0}; int cnt = 0;
            auto &H = tab ? h2 : h;
            for (int t = start_ts; t < end_ts && t < 4096; t++) { cnt++; for (int k = 0; k < 8; k++) { sum[k] += (double)H[t][k]; if ((double)H[t][k] > mx[k]) mx[k] = (double)H[t][k]; } }
            const char **nm = tab ? nm_p : (w->node_rl ? nm_rl : nm_w);
            for (int k = 0; k < 8; k++) fprintf(stderr, "[node_prof] %-18s mean of per-step max %9.0f cycles, max %9.0f\n", nm[k], sum[k] / (cnt ? cnt : 1), mx[k]);
        }
        static unsigned long long z[4096][8]; cudaMemcpyToSymbol(g_node_prof, z, sizeof(z)); cudaMemcpyToSymbol(g_node_prof2, z, sizeof(z));
    }
#endif
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
        if (rem) for (int r = 0; r < w->R; r++) CUDA_OK(cudaMemcpy(field == UX_LINK_CAPACITY_OUT_REMAIN ? &LR(w->hw[r], link).cap_out_rem : &LR(w->hw[r], link).cap_in_rem, &rv, sizeof(double), cudaMemcpyHostToDevice));
        else w->net_dirty = true;
    }
    return 0;
}

UX_API int ux_set_link_signal_group(ux_world *w, int link, const int *sg, int n) {
    if (link < 0 || link >= (int)w->links.size()) return fail(w, UX_ERR_OUT_OF_RANGE, "link id out of range");
    bool resized = n != (int)w->links[link].sg.size();
    w->links[link].sg.assign(sg, sg + n);
    if (w->uploaded && resized) w->sig_tables_dirty = true;  // the CSR of all groups is rebuilt before the next step
    else if (w->uploaded) {
        cudaSetDevice(w->device);
        int off = 0; for (int i = 0; i < link; i++) off += (int)w->links[i].sg.size();
        CUDA_OK(cudaMemcpy(w->d_sg + off, sg, sizeof(int) * n, cudaMemcpyHostToDevice));
    }
    return 0;
}

UX_API int ux_set_node_signal(ux_world *w, int node, const double *sig, int nsig, int phase, double sig_t) {
    if (node < 0 || node >= (int)w->nodes.size()) return fail(w, UX_ERR_OUT_OF_RANGE, "node id out of range");
    HNode &n = w->nodes[node];
    bool resized = false;
    if (sig && nsig > 0) { resized = nsig != (int)n.sig.size(); n.sig.assign(sig, sig + nsig); }
    if (!w->uploaded) {
        if (sig && nsig > 0 && n.sig_phase0 >= nsig) { n.sig_phase0 = 0; n.sig_t0 = 0; }
        if (phase >= 0) n.sig_phase0 = phase;
        if (sig_t >= 0) n.sig_t0 = sig_t;
        return 0;
    }
    cudaSetDevice(w->device);
    if (resized) {
        // a different number of phases: the CSR of all signal plans is rebuilt before the next step; a running phase beyond
        // the new plan restarts at phase 0 (the reference would index past the end of the list)
        w->sig_tables_dirty = true;
        for (int r = 0; r < w->R; r++) {
            int ph = 0; CUDA_OK(cudaMemcpy(&ph, &NS(w->hw[r], node).sig_phase, sizeof(int), cudaMemcpyDeviceToHost));
            if (ph >= nsig && phase < 0) { int z = 0; double zt = 0.0; CUDA_OK(cudaMemcpy(&NS(w->hw[r], node).sig_phase, &z, sizeof(int), cudaMemcpyHostToDevice)); CUDA_OK(cudaMemcpy(&NR(w->hw[r], node).sig_t, &zt, sizeof(double), cudaMemcpyHostToDevice)); }
        }
    } else if (sig && nsig > 0) { int off = 0; for (int i = 0; i < node; i++) off += (int)w->nodes[i].sig.size(); CUDA_OK(cudaMemcpy(w->d_n_sig + off, sig, sizeof(double) * nsig, cudaMemcpyHostToDevice)); }
    for (int r = 0; r < w->R; r++) {
        if (phase >= 0) CUDA_OK(cudaMemcpy(&NS(w->hw[r], node).sig_phase, &phase, sizeof(int), cudaMemcpyHostToDevice));
        if (sig_t >= 0) CUDA_OK(cudaMemcpy(&NR(w->hw[r], node).sig_t, &sig_t, sizeof(double), cudaMemcpyHostToDevice));
    }
    return 0;
}

// ---- results ---------------------------------------------------------------------------------------------
