// ux_kernels_step.cuh -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// The two kernels of a timestep: k_node (link update, generate, signal, transfer) and k_update (car following, link head, logs).

#ifdef UX_PROFILE_NODE  // measurement build only (uxsim_b200/variants/): per step, the longest time any thread spent in each phase of k_node_rl
__device__ unsigned long long g_node_prof[4096][8];
__device__ unsigned long long g_node_prof2[4096][8];  // warp form: phases of pre_node
#define NPROF_T(v) const long long v = clock64();
#define NPROF_MAX(k, a, b) atomicMax(&g_node_prof[t & 4095][k], (unsigned long long)((b) - (a)));
#define NPROF2_MAX(k, a, b) atomicMax(&g_node_prof2[t & 4095][k], (unsigned long long)((b) - (a)));
#define NLAP_INIT long long lap_ = clock64(); const long long lap0_ = lap_;
#define NLAP(tab, k) { long long now_ = clock64(); if ((threadIdx.x & 31) == 0) atomicMax(&tab[t & 4095][k], (unsigned long long)(now_ - lap_)); lap_ = now_; }
#define NLAP_TOTAL(tab, k) if ((threadIdx.x & 31) == 0) atomicMax(&tab[t & 4095][k], (unsigned long long)(clock64() - lap0_));
#else
#define NPROF_T(v)
#define NPROF_MAX(k, a, b)
#define NPROF2_MAX(k, a, b)
#define NLAP_INIT
#define NLAP(tab, k)
#define NLAP_TOTAL(tab, k)
#endif
// ---------------------------------------------------------------------------------------------------
// K1: Link::update (traffi.cpp:542-565, 599-622) for the node's out-links, then Node::generate (105-189),
//     Node::signal_update (194-207), Node::flow_capacity_update (226-234).  One warp per node.
// ---------------------------------------------------------------------------------------------------
// Node::signal_update (traffi.cpp:194-207) and Node::flow_capacity_update (226-234) of node n at step t
// `ln` = replica offset from W's own replica (0 unless the caller's lanes are the replicas of a group whose first world is W)
UX_DEV void node_signal_capacity(const DevWorld &W, int n, int t, int ln = 0) {
    int s0 = W.n_sig_off[n], nsig = W.n_sig_off[n + 1] - s0;
    NodeS *ns = &NS(W, n) + ln; NodeR *nr = &NR(W, n) + ln;
    int ph = ns->sig_phase;
    if (nsig > 1) {
        double st = nr->sig_t;
        if (st > W.n_sig[s0 + ph]) { ph++; st = 0; if (ph >= nsig) ph = 0; }
        st += W.dt;
        nr->sig_t = st; ns->sig_phase = ph;
    }
    W.sig_log[TNI(W, t, n) + ln] = ph;
    double fc = W.n_fc[n];
    if (fc >= 0.0) { int nl = W.n_lanes[n]; if (nr->fc_rem < W.dn * (nl > 0 ? nl : 1)) nr->fc_rem += fc * W.dt; }
    else nr->fc_rem = 10e10;
}

template <int MC, int MD>
struct PreSmem {
    // generation attempts (the first platoons of the node's vertical queue)
    int choice[MC], g_vid[MC], g_err[MC], g_trav[MC], g_slot[MC];
    unsigned char g_q[MC], g_lane[MC];
    int tried, accepted;
    // out-links
    int o_hd[MD], o_tail[MD], o_tail0[MD], o_lanes[MD], o_cap[MD], o_lastlane[MD], o_win_off[MD], o_win_n[MD], o_win_f[MD];
    long long o_roff[MD];
    double o_ci[MD], o_dpl[MD], o_vmax[MD], o_cumarr[MD];
    double w_x[MC];  // window of the last `lanes` platoons of each out-link; circular once full (o_win_f = oldest entry)
};

// One warp runs the first half of a node's timestep.  Everything it reads is known when the kernel starts, so the loads of all
// parts are issued in ONE parallel phase, each part on its own lanes: the route draws of the waiting platoons (the longest chain),
// the out-links' inflow side (Link::update + tail state + window), the in-links' outflow side, the signal / node capacity step.
// Lane 0 then walks the generation attempts on shared memory only, and the lanes apply what it decided.
template <int MC, int MD>
UX_DEV void pre_node(const DevWorld &W, PreSmem<MC, MD> &G, double *tt_s, int n, int t, int lane) {
    int cur = t & 1, L = W.L;
    size_t C = (size_t)W.C;
    int o0 = W.n_out_off[n], nout = W.n_out_off[n + 1] - o0;
    int i0 = W.n_in_off[n], nin = W.n_in_off[n + 1] - i0;
    bool fresh = (t % W.itt) == 0 && !W.node_rl;  // (ensembles: k_link_pre<true> has measured the mean speeds already)
    const bool split = W.pre_split != 0;  // the scalar link update and the node's signal / capacity step ran before this kernel (k_link_pre)
    NLAP_INIT
    // ---- instantaneous travel time of the out-links (whole warp per link, only every `itt` steps)
    if (fresh) {
        for (int k = 0; k < nout; k++) {
            int l = W.n_out[o0 + k];
            int hd = LQ(W, l).head[cur], nq = LQ(W, l).tail - hd;
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
    NLAP(g_node_prof2, 0)
    // ---- the parallel phase.  Attempts: the vertical queue of an origin is sorted by departure step, and platoons with departure
    //      step <= t-1 are in it (HOME -> WAIT, traffi.cpp:836-841), so the attempts of this step are the released prefix of the next
    //      `sum of out-link lanes` entries; every lane tests its own entry.
    const int g0 = W.gen_off[n], glen = W.gen_off[n + 1] - g0;
    int gh = NS(W, n).gen_head;
    int amax = glen - gh;
    { int tl_ = W.n_out_lanes[n]; if (amax > tl_) amax = tl_; if (amax > MC) amax = MC; }
    if (nout == 0) amax = 0;
    const int rB = amax, rC = rB + nout, rD = rC + (split ? 0 : nin);  // first lane of the out-link / in-link / node parts
    WARP_FOR(ln) {
        for (int i = ln; i < amax; i += 32) {  // Node::generate, the draw (traffi.cpp:112-128)
            int vid = W.gen_list[g0 + gh + i], e = 0, ch = -1, q = 255, trav = 0;
            if (W.gen_ts[g0 + gh + i] <= t - 1) {
                trav = W.v_trav[vid];
                ch = route_choice(W, vid, n, t, (unsigned)n, UX_RNG_GENERATE, (unsigned)i, &e);
                if (ch >= 0) { q = 0; while (q < nout && W.n_out[o0 + q] != ch) q++; }
            } else vid = -1;
            G.g_vid[i] = vid; G.choice[i] = ch; G.g_err[i] = e; G.g_q[i] = (unsigned char)q; G.g_trav[i] = trav;
        }
        for (int q = (ln - rB) & 31; q < nout; q += 32) {  // out-links: Link::update of the inflow side (traffi.cpp:542-565), tail state, window
            int ol = W.n_out[o0 + q], woff = W.n_out_woff[o0 + q];
            int hd = LQ(W, ol).head[cur], tl = LQ(W, ol).tail, lanes = W.l_lanes[ol], cap = W.l_rcap[ol];
            long long o = W.l_roff[ol];
            size_t row = TLI(W, t, ol), prev = TLI(W, t - 1, ol);
            // loads first, in two rounds (the record, then the ring entries it points at); stores afterwards
            double ci_ = LR(W, ol).cap_in_rem, dpl_ = W.l_dpl_dn[ol], vm_ = W.l_vmax[ol], cin = W.l_cap_in[ol];
            double ca_ = (!split && t != 0) ? W.cum_arr[prev] : W.cum_arr[row];
            double ttp_ = (!split && !fresh && t > 0) ? W.tt_inst[prev] : 0.0;
            int sz = tl - hd, wn = sz < lanes ? sz : lanes;
            int ll_ = sz > 0 ? W.LANE[o + ring_slot(tl - 1, 0, cap)] : -1;
            if (fresh) W.tt_inst[row] = tt_s[q];
            else if (!split && t > 0) W.tt_inst[row] = ttp_;
            if (!split) { if (cin < 10e9) { if (ci_ < W.dn * lanes) ci_ += cin * W.dt; } else ci_ = 10e9; }
            G.o_hd[q] = hd; G.o_tail[q] = tl; G.o_tail0[q] = tl; G.o_lanes[q] = lanes; G.o_cap[q] = cap; G.o_roff[q] = o;
            G.o_ci[q] = ci_; G.o_dpl[q] = dpl_; G.o_vmax[q] = vm_; G.o_cumarr[q] = ca_; G.o_lastlane[q] = ll_;
            G.o_win_off[q] = woff; G.o_win_n[q] = wn; G.o_win_f[q] = 0;
        }
        {   // the window of every out-link (its last `lanes` platoons), one lane per entry
            int wtot = W.n_out_lanes[n]; if (wtot > MC) wtot = MC; if (amax == 0) wtot = 0;
            for (int e = (ln - rD - 1) & 31; e < wtot; e += 32) {
                int q = 0;
                while (q + 1 < nout && W.n_out_woff[o0 + q + 1] <= e) q++;
                int ol = W.n_out[o0 + q], k = e - W.n_out_woff[o0 + q];
                int hd = LQ(W, ol).head[cur], sz = LQ(W, ol).tail - hd, lanes = W.l_lanes[ol], wn = sz < lanes ? sz : lanes;
                if (k < wn) G.w_x[e] = W.X[(size_t)cur * C + W.l_roff[ol] + ring_slot(hd, sz - wn + k, W.l_rcap[ol])];
            }
        }
        if (!split) {
            for (int k = (ln - rC) & 31; k < nin; k += 32) {  // in-links: the outflow side (departure curve, capacity_out tokens, pops of this step)
                int l = W.n_in[i0 + k];
                size_t row = TLI(W, t, l);
                if (t != 0) W.cum_dep[row] = W.cum_dep[TLI(W, t - 1, l)];
                double co = W.l_cap_out[l];
                if (co < 10e9) { if (LR(W, l).cap_out_rem < W.dn * W.l_lanes[l]) LR(W, l).cap_out_rem += co * W.dt; } else LR(W, l).cap_out_rem = 10e9;
                LC(W, l).pop_k2 = 0;
            }
            if (ln == (rD & 31)) node_signal_capacity(W, n, t);
        }
    }
    WARP_SYNC();
    NLAP(g_node_prof2, 3)
    // ---- generate (traffi.cpp:105-189): lane 0 walks the attempts in order on shared memory; the first refusal ends them
    WARP_FOR(ln) if (ln == 0) {
        int tried = 0, accepted = 0;
        for (int i = 0; i < amax; i++) {
            int vid = G.g_vid[i];
            if (vid < 0) break;  // not released yet
            if (G.g_err[i]) { dev_error(W, G.g_err[i], vid); break; }
            tried = i + 1;
            if (G.choice[i] < 0) break;
            int q = G.g_q[i];
            int lanes = G.o_lanes[q], sz = G.o_tail[q] - G.o_hd[q], w0 = G.o_win_off[q];
            bool ok = false;  // traffi.cpp:130-142
            if (sz < lanes) ok = true;
            else if (G.w_x[w0 + G.o_win_f[q]] > G.o_dpl[q]) ok = true;
            if (G.o_ci[q] < W.dn) ok = false;
            if (!ok) break;
            if (sz >= G.o_cap[q]) { dev_error(W, DERR_RING_FULL, G.choice[i]); break; }
            accepted = i + 1;
            int lane_new = sz > 0 ? (G.o_lastlane[q] + 1) % lanes : 0;
            G.g_slot[i] = ring_slot(G.o_tail[q], 0, G.o_cap[q]); G.g_lane[i] = (unsigned char)lane_new;
            G.o_tail[q] += 1; G.o_lastlane[q] = lane_new;
            if (G.o_win_n[q] < lanes) { G.w_x[w0 + G.o_win_n[q]] = 0.0; G.o_win_n[q] += 1; }
            else { G.w_x[w0 + G.o_win_f[q]] = 0.0; G.o_win_f[q] = G.o_win_f[q] + 1 == lanes ? 0 : G.o_win_f[q] + 1; }
            G.o_cumarr[q] += W.dn;
            G.o_ci[q] -= W.dn;
        }
        G.tried = tried; G.accepted = accepted;
        if (accepted) NS(W, n).gen_head = gh + accepted;
        { int nv = 0; while (nv < amax && G.g_vid[nv] >= 0) nv++; if (gh + nv > NS(W, n).gen_avail) NS(W, n).gen_avail = gh + nv; }  // released so far (the replica-lane form keeps counting from here)
    }
    WARP_SYNC();
    NLAP(g_node_prof2, 5)
    // ---- apply: the platoons that were tried / generated, then the out-link scalars
    WARP_FOR(ln) {
        const int tried = G.tried, accepted = G.accepted;
        for (int i = ln; i < tried; i += 32) {
            int vid = G.g_vid[i], ol = G.choice[i];
            W.v_next[vid] = ol;
            if (i >= accepted) continue;
            int q = G.g_q[i], slot = G.g_slot[i];
            long long o = G.o_roff[q];
            W.X[(size_t)cur * C + o + slot] = 0.0;
            W.X[(size_t)(cur ^ 1) * C + o + slot] = 0.0;
            W.V[o + slot] = G.o_vmax[q];
            W.VID[o + slot] = vid;
            W.LANE[o + slot] = G.g_lane[i];
            W.v_state[vid] = UX_VS_RUN;
            if (W.log_mode) W.v_gen_ts[vid] = t;
            W.v_atl[vid] = (double)t * W.dt;
            W.v_entry_x[vid] = 0.0;
            // mark_entered (traffi.cpp:160-166)
            if (W.no_cyclic) { int sn = W.l_start[ol]; W.v_visited[(size_t)vid * W.vis_words + (sn >> 5)] |= 1u << (sn & 31); }
            W.v_trav[vid] = G.g_trav[i] + 1;
            W.v_link[vid] = ol;
            record_traversal(W, vid, G.g_trav[i], ol, t);
        }
        for (int q = (ln - rB) & 31; q < nout; q += 32) {
            int ol = W.n_out[o0 + q];
            const bool pushed = G.o_tail[q] != G.o_tail0[q];
            if (!split || pushed) { LR(W, ol).cap_in_rem = G.o_ci[q]; W.cum_arr[TLI(W, t, ol)] = G.o_cumarr[q]; }
            if (pushed) LQ(W, ol).tail = G.o_tail[q];
            // tail after generation (the link update tells transfer pushes from generated ones): k_link_pre already copied the unchanged tails
            if (!split || pushed) LQ(W, ol).tail_k1 = G.o_tail[q];
        }
    }
    WARP_SYNC();
    NLAP(g_node_prof2, 6)
    NLAP_TOTAL(g_node_prof2, 7)
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
    // head positions of the in-links (gathered once; the serial loop below reads and writes only these)
    int p_vid[MC], p_next[MC], p_dts[MC], p_trav[MC], p_es[MC];
    short p_in[MC], p_j[MC], c_idx[MC], expd[MC];
    unsigned char p_flag[MC], p_q[MC];  // p_q: index of the requested out-link (255: none)
    double p_xold[MC], p_v[MC], p_mr[MC], p_atl[MC], p_dist[MC];  // p_mr: residual move, rescaled to the wanted out-link's speed before the loop; p_dist: distance incl. this link
    // what the loop decided for a position, applied to global memory afterwards by all lanes
    unsigned char p_act[MC], p_olane[MC];  // 0 untouched, 1 moved onto out-link p_q, 2 trip ended
    int p_ord[MC], p_oslot[MC];            // exit-event number on the in-link; ring slot on the out-link (-1: ring full)
    double p_xn[MC];                       // position on the out-link
    int nc, npos, nol, any_wait, nexp;
    int omask[MD];             // per requested out-link: in-links whose front platoon may move onto it now (bit per in-link)
    unsigned char ogood[MD];   // per requested out-link: acceptance test result
    unsigned char c_first[MC];
    double fc_rem;
    short shuf_j[MC];
    double u_merge[MC];
    // requested out-links
    int ol[MD], o_hd[MD], o_tail[MD], o_lanes[MD], o_cap[MD], o_lastlane[MD], o_win_off[MD + 1], o_win_n[MD], o_win_f[MD];
    long long o_roff[MD];
    double o_ci[MD], o_dpl[MD], o_len[MD], o_vmax[MD], o_cumarr[MD];
    double w_x[MC], w_xold[MC];  // window of the last `lanes` platoons of each out-link; circular once full (o_win_f = oldest entry)
};

// Vehicle::end_trip (traffi.cpp:886-927) of the front platoon (position p) of an in-link, inside the serial loop: counters only,
// the per-platoon stores follow in the apply phase
template <int MC, int MD>
UX_DEV void xfer_end_trip(const DevWorld &W, XferSmem<MC, MD> &S, int p) {
    int ii = S.p_in[p];
    S.trips[ii] += 1;
    S.cumdep[ii] += W.dn;
    S.p_ord[p] = ++S.evc[ii];
    S.p_act[p] = 2;
    S.popped[ii] += 1;
}

// the out-link (index) that the front platoon of in-link ii may move onto now, 255 if none (traffi.cpp:305-330: front of its
// link, outflow tokens, green)
template <int MC, int MD>
UX_DEV int xfer_front_q(const DevWorld &W, const XferSmem<MC, MD> &S, int ii, int npos) {
    if (S.popped[ii] >= S.hc[ii]) return 255;
    int fp = S.pos_off[ii] + S.popped[ii];
    if (fp >= npos || !(S.p_flag[fp] & VF_CAND)) return 255;
    if (S.co_rem[ii] < W.dn || !S.sig_ok[ii]) return 255;
    return S.p_q[fp];
}

template <int MC, int MD>
UX_DEV bool xfer_accepts(const DevWorld &W, const XferSmem<MC, MD> &S, int q) {  // traffi.cpp:285-300 (the node capacity is tested by the caller)
    int lanes_ol = S.o_lanes[q], sz = S.o_tail[q] - S.o_hd[q];
    bool ok = false;
    if (sz < lanes_ol) ok = true;
    else if (S.w_x[S.o_win_off[q] + S.o_win_f[q]] > S.o_dpl[q]) ok = true;
    if (S.o_ci[q] < W.dn) ok = false;
    return ok;
}

#ifndef XFER_WARPS
#define XFER_WARPS 4
#endif
// Returns true when it has published the node's completion itself (`publish`: some node may be waiting for this one).
template <int MC, int MD>
UX_DEV bool xfer_node(const DevWorld &W, XferSmem<MC, MD> &S, int n, int t, bool has_deps, bool publish) {
    int cur = t & 1, L = W.L;
    size_t C = (size_t)W.C;
    NLAP_INIT
    if (NS(W, n).node_act != t) return false;  // no in-link of this node has a flagged head position (hcount == 0 everywhere): nothing to do
    int i0 = W.n_in_off[n], nin = W.n_in_off[n + 1] - i0;
    if (nin == 0) return false;
    int nsig = W.n_sig_off[n + 1] - W.n_sig_off[n], phase = NS(W, n).sig_phase;
    // ---- A1: in-link state
    WARP_FOR(lane) {
        for (int ii = lane; ii < nin; ii += 32) {
            // loads first, shared stores afterwards: the pointers are generic, so a shared store between two loads
            // would serialise them (the compiler must assume the load may target shared memory)
            int il = W.n_in[i0 + ii];
            int hd_ = LQ(W, il).head[cur], hc_ = LC(W, il).hcount, cap_ = W.l_rcap[il], evc_ = LC(W, il).ev_cnt, trips_ = LC(W, il).trips;
            long long roff_ = W.l_roff[il];
            double co_ = LR(W, il).cap_out_rem, mp_ = W.l_mp[il], vm_ = W.l_vmax[il], cd_ = W.cum_dep[TLI(W, t, il)];
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
    if (npos == 0) return false;
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
            double x_ = 0, xo_ = 0, v_ = 0, mr_ = 0, atl_ = 0, ex_ = 0, di_ = 0; int dts_ = 0, trav_ = 0;
            if (f & (VF_CAND | VF_WAIT_END)) {
                x_ = W.X[(size_t)cur * C + o + slot]; xo_ = W.X[(size_t)(cur ^ 1) * C + o + slot]; v_ = W.V[o + slot];
                mr_ = W.v_mr[vid]; atl_ = W.v_atl[vid]; ex_ = W.v_entry_x[vid]; di_ = W.v_dist[vid]; dts_ = W.v_depart_ts[vid];
                if (f & VF_CAND) trav_ = W.v_trav[vid];
            }
            int es = (int)(atl_ / W.dt);  // entry step of the link-exit event (record_tt_event, traffi.cpp:357-362)
            if (es < 0) es = 0;
            if (es >= W.T) es = W.T - 1;
            S.p_vid[p] = vid; S.p_in[p] = (short)ii; S.p_j[p] = (short)j; S.p_flag[p] = f; S.p_next[p] = nx; S.p_q[p] = 255; S.p_act[p] = 0;
            S.p_xold[p] = xo_; S.p_v[p] = v_; S.p_mr[p] = mr_; S.p_atl[p] = atl_; S.p_dist[p] = di_ + (x_ - ex_); S.p_dts[p] = dts_;
            S.p_trav[p] = trav_; S.p_es[p] = es;
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
    if (nc == 0 && !S.any_wait) return false;
    WARP_FOR(lane) {
        for (int k = lane; k < nc; k += 32) {
            int p = S.c_idx[k], nl = S.p_next[p], q = 255;
            for (int r = 0; r < nol; r++) if (S.ol[r] == nl) { q = r; break; }
            S.p_q[p] = (unsigned char)q;
        }
    }
    NLAP(g_node_prof, 0)
    // ---- dependency wait (only nodes with SEES_POPS out-links): everything above -- in-link heads, candidate sort -- is
    //      independent of the lower-level nodes, so it overlaps their work; only the out-link tail state needs their pops.
    //      The wait is per REQUESTED out-link: only a link some candidate wants to enter is looked at below, and its end node can
    //      pop from it only if the last link update flagged head positions there (hcount > 0, final before this launch).
    if (has_deps) {
        WARP_FOR(lane) {
            for (int q = lane; q < nol; q += 32) {
                int ol = S.ol[q];
                if (!(W.l_flags[ol] & LF_SEES_POPS)) continue;
                const int hc_ = LC(W, ol).hcount;
                if (hc_ == 0) continue;  // nothing can leave that link in this step
                // ... and even if all flagged platoons leave, a link that keeps `lanes` platoons looks the same from its tail: the
                // acceptance test, the carry-over bound and the lane rule read the platoon `lanes` positions from the tail
                if (LQ(W, ol).tail - LQ(W, ol).head[cur] - hc_ >= W.l_lanes[ol]) continue;
                volatile int *done = &NDONE(W, W.l_end[ol]);
#ifdef UX_EMU
                if (*done < t + 1) dev_error(W, DERR_DEP_TIMEOUT, n);
#else
                // bounded: the awaited warp has a smaller linear index and is normally resident or finished already; if the
                // launch order ever broke that assumption the step fails with an error code instead of hanging the device
                long long spins = 0;
                while (*done < t + 1) { __nanosleep(64); if (++spins > UX_DEP_SPIN_BUDGET) { dev_error(W, DERR_DEP_TIMEOUT, n); break; } }
#endif
            }
        }
#ifndef UX_EMU
        __threadfence();
#endif
        WARP_SYNC();
    }
    NLAP(g_node_prof, 1)
    // ---- A4: out-link tail state
    WARP_FOR(lane) {
        for (int q = lane; q < nol; q += 32) {
            int ol = S.ol[q];
            int hd = LQ(W, ol).head[cur], fl_ = W.l_flags[ol];
            // pops of this step by the link's end node (a lower level of this launch).  Only a node with flagged head positions on
            // this link can pop from it -- otherwise the end node may not even have reset last step's value yet (small worlds reset
            // it inside this launch, in that node's own pre_node), so the count is taken as 0 without reading it.
            int tl = LQ(W, ol).tail, lanes = W.l_lanes[ol], cap = W.l_rcap[ol];
            int pk_ = 0;
            if (fl_ & LF_SEES_POPS) { int hc_ = LC(W, ol).hcount; if (hc_ > 0 && tl - hd - hc_ < lanes) pk_ = ld_cg(&LC(W, ol).pop_k2); }  // (same test as the wait above)
            long long o = W.l_roff[ol];
            double ci_ = LR(W, ol).cap_in_rem, dpl_ = W.l_dpl_dn[ol], len_ = W.l_length[ol], vm_ = W.l_vmax[ol], ca_ = W.cum_arr[TLI(W, t, ol)];
            if (fl_ & LF_SEES_POPS) hd += pk_;
            int sz = tl - hd;
            int ll_ = sz > 0 ? W.LANE[o + ring_slot(tl - 1, 0, cap)] : -1;
            S.o_hd[q] = hd; S.o_tail[q] = tl; S.o_lanes[q] = lanes; S.o_cap[q] = cap; S.o_roff[q] = o;
            S.o_ci[q] = ci_; S.o_dpl[q] = dpl_; S.o_len[q] = len_; S.o_vmax[q] = vm_; S.o_cumarr[q] = ca_;
            S.o_lastlane[q] = ll_;
            S.o_win_n[q] = sz < lanes ? sz : lanes; S.o_win_f[q] = 0; S.omask[q] = 0;
        }
    }
    WARP_SYNC();
    WARP_FOR(lane) if (lane == 0) {
        int acc = 0, nexp = 0;
        for (int q = 0; q < nol; q++) { S.o_win_off[q] = acc; acc += S.o_lanes[q]; for (int r = 0; r < S.o_lanes[q] && nexp < MC; r++) S.expd[nexp++] = (short)q; }
        S.o_win_off[nol] = acc; S.nexp = nexp;
    }
    WARP_SYNC();
    int nexp = S.nexp;
    int nshuf = (!W.hard_det && nexp > 1) ? nexp - 1 : 0;
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
        // all random numbers of the node (shuffle indices, merge-lottery uniforms) are drawn up front: Philox is stateless
        if (!W.hard_det) {
            for (int i = 1 + lane; i < nexp; i += 32)  // Fisher-Yates step i uses draw number (nexp-1-i) (traffi.cpp:276-282)
                S.shuf_j[i] = (short)ux_rng_below(W.seed, (uint32_t)t, (uint32_t)n, UX_RNG_TRANSFER, (uint32_t)(nexp - 1 - i), i + 1);
            for (int d = lane; d < nexp; d += 32)
                S.u_merge[d] = ux_rng_uniform(W.seed, (uint32_t)t, (uint32_t)n, UX_RNG_TRANSFER, (uint32_t)(nshuf + d));
        }
        // the residual move of every candidate in the speed of the link it wants to enter (traffi.cpp:400-402): the division stays out of the loop
        for (int k = lane; k < nc; k += 32) {
            int p = S.c_idx[k], q = S.p_q[p];
            if (q != 255) S.p_mr[p] = S.p_mr[p] * S.o_vmax[q] / S.vmax[S.p_in[p]];
        }
    }
    WARP_SYNC();
    NLAP(g_node_prof, 2)
    WARP_FOR(lane) {
        for (int q = lane; q < nol; q += 32) S.ogood[q] = xfer_accepts(W, S, q) ? 1 : 0;
        for (int ii = lane; ii < nin; ii += 32) { int f = xfer_front_q(W, S, ii, npos); if (f != 255) atomicOr(&S.omask[f], 1 << ii); }
    }
    WARP_SYNC();
    NLAP(g_node_prof, 3)
    // ---- B: the slot loop (traffi.cpp:283-430), lane 0, on shared memory only.  Slots are inherently sequential -- each move
    //          changes what the next slot sees -- but at any moment an in-link offers at most ONE platoon (its front), so the state a
    //          slot looks at is one mask of in-links per out-link plus the out-link's acceptance flag, both refreshed only for the
    //          links a move touched.  Global memory sees nothing here: what the loop decides per head position (moved where / trip
    //          ended, event number, slot, lane, position) is applied afterwards by all lanes.
    double fc = W.n_fc[n];
    WARP_FOR(lane) if (lane == 0) {
        if (!W.hard_det) for (int i = nexp - 1; i > 0; i--) { int j = S.shuf_j[i]; short tmp = S.expd[i]; S.expd[i] = S.expd[j]; S.expd[j] = tmp; }
        double fc_rem = NR(W, n).fc_rem;
        int n_picks = 0;
        for (int s = 0; s < nexp; s++) {
            if (fc_rem < W.dn) break;  // no slot can move anything any more (traffi.cpp:298-300)
            const int q = S.expd[s];
            int m = S.omask[q];
            if (!m || !S.ogood[q]) continue;
            int ii;
            if (!(m & (m - 1))) {  // one in-link offers a platoon (the usual case): no lottery, but its draw is consumed (traffi.cpp:332-340)
                ii = 0; while (!((m >> ii) & 1)) ii++;
                n_picks++;
            } else {
                short mv[MD]; double mp[MD]; int nm = 0;
                while (m) {  // eligible fronts in ascending vehicle id
                    int i2 = 0; while (!((m >> i2) & 1)) i2++;
                    m &= m - 1;
                    int vid = S.p_vid[S.pos_off[i2] + S.popped[i2]], at = nm++;
                    while (at > 0 && S.p_vid[S.pos_off[mv[at - 1]] + S.popped[mv[at - 1]]] > vid) { mv[at] = mv[at - 1]; at--; }
                    mv[at] = (short)i2;
                }
                for (int i = 0; i < nm; i++) mp[i] = S.mp[mv[i]];
                int pick;
                if (W.hard_det) { pick = 0; for (int i = 1; i < nm; i++) if (mp[i] > mp[pick]) pick = i; }
                else pick = weighted_pick(mp, nm, S.u_merge[n_picks++]);
                ii = mv[pick];
            }
            const int p = S.pos_off[ii] + S.popped[ii];
            const int lanes_ol = S.o_lanes[q], sz = S.o_tail[q] - S.o_hd[q], w0 = S.o_win_off[q];
            S.co_rem[ii] -= W.dn; S.o_ci[q] -= W.dn;
            if (fc >= 0.0) fc_rem -= W.dn;
            S.cumdep[ii] += W.dn; S.o_cumarr[q] += W.dn;
            S.p_ord[p] = ++S.evc[ii];
            S.p_act[p] = 1;
            S.popped[ii] += 1;
            // carry the residual move onto the new link (traffi.cpp:400-419)
            double x_next = S.p_mr[p];
            if (sz >= lanes_ol) {
                double x_cong = S.w_xold[w0 + S.o_win_f[q]] - S.o_dpl[q];
                if (x_cong < 0.0) x_cong = 0.0;
                if (x_next > x_cong) x_next = x_cong;
            }
            if (x_next >= S.o_len[q]) x_next = S.o_len[q];
            if (x_next < 0.0) x_next = 0.0;
            S.p_xn[p] = x_next;
            if (sz >= S.o_cap[q]) { dev_error(W, DERR_RING_FULL, S.ol[q]); S.p_oslot[p] = -1; }
            else {  // push (lane rule traffi.cpp:385-390)
                int lane_new = sz > 0 ? S.o_lastlane[q] + 1 : 0;  // (lane + 1) % lanes: lane labels are < lanes
                if (lane_new >= lanes_ol) lane_new = 0;
                S.p_oslot[p] = ring_slot(S.o_tail[q], 0, S.o_cap[q]); S.p_olane[p] = (unsigned char)lane_new;
                S.o_tail[q] += 1; S.o_lastlane[q] = lane_new;
                if (S.o_win_n[q] < lanes_ol) { int e = w0 + S.o_win_n[q]; S.w_x[e] = x_next; S.w_xold[e] = S.p_xold[p]; S.o_win_n[q] += 1; }
                else { int e = w0 + S.o_win_f[q]; S.w_x[e] = x_next; S.w_xold[e] = S.p_xold[p]; S.o_win_f[q] = S.o_win_f[q] + 1 == lanes_ol ? 0 : S.o_win_f[q] + 1; }
            }
            // the new front of the in-link may be waiting for its trip end (traffi.cpp:421-424)
            if (S.popped[ii] < S.hc[ii]) {
                int fp = S.pos_off[ii] + S.popped[ii];
                if (fp < npos && (S.p_flag[fp] & VF_WAIT_END)) xfer_end_trip(W, S, fp);
            }
            // what this move changed for the slots to come: the in-link's front and tokens, the out-link's tail
            S.omask[q] &= ~(1 << ii);
            { int f = xfer_front_q(W, S, ii, npos); if (f != 255) S.omask[f] |= 1 << ii; }
            S.ogood[q] = xfer_accepts(W, S, q) ? 1 : 0;
        }
        // flush trip-end-waiting fronts (traffi.cpp:432-441)
        for (int ii = 0; ii < nin; ii++) {
            int lanes = S.hc[ii];  // hcount = head positions up to the last flagged one: positions beyond it cannot be waiting
            for (int r = 0; r < lanes; r++) {
                if (S.popped[ii] >= S.hc[ii]) break;
                int fp = S.pos_off[ii] + S.popped[ii];
                if (fp >= npos || !(S.p_flag[fp] & VF_WAIT_END)) break;
                xfer_end_trip(W, S, fp);
            }
        }
        NR(W, n).fc_rem = fc_rem;
        if (publish) {
            // A waiting node needs one thing of this node: how many platoons left each in-link.  That is final here, so the
            // completion flag goes out before the apply phase (nothing it writes is read by another node of this launch).
            for (int ii = 0; ii < nin; ii++) LC(W, S.il[ii]).pop_k2 = S.popped[ii];
#ifndef UX_EMU
            __threadfence();
#endif
            *((volatile int *)&NDONE(W, n)) = t + 1;
        }
    }
    WARP_SYNC();
    NLAP(g_node_prof, 4)
    // ---- C: apply.  One lane per head position the loop touched: the per-platoon fields, the ring slot on the new link, the
    //          link-exit travel-time event (of two events of one in-link with the same entry step the later one stays, as in the
    //          serial order); then the per-link accumulators.
    WARP_FOR(lane) {
        for (int p = lane; p < npos; p += 32) {
            const int act = S.p_act[p];
            if (!act) continue;
            const int ii = S.p_in[p], il = S.il[ii], vid = S.p_vid[p], ord = S.p_ord[p], es = S.p_es[p];
            bool last = true;
            { int pe = S.pos_off[ii + 1] < npos ? S.pos_off[ii + 1] : npos; for (int p2 = S.pos_off[ii]; p2 < pe; p2++) if (S.p_act[p2] && S.p_es[p2] == es && S.p_ord[p2] > ord) { last = false; break; } }
            const double now = (double)t * W.dt;
            if (last) { W.ev_ord[TLI(W, es, il)] = ord; W.ev_tt[TLI(W, es, il)] = (act == 2 ? ((double)t + 1.0) * W.dt : now) - S.p_atl[p]; }
            W.v_atl[vid] = now;
            W.v_dist[vid] = S.p_dist[p];
            W.v_flags[vid] = 0;
            if (act == 2) {  // Vehicle::end_trip (traffi.cpp:886-927)
                W.v_link[vid] = -1;
                if (S.p_flag[p] & VF_ABORT) { W.v_state[vid] = UX_VS_ABORT; W.v_arr_ts[vid] = -1; W.v_tt[vid] = -1.0; }
                else { W.v_state[vid] = UX_VS_END; W.v_arr_ts[vid] = t; W.v_tt[vid] = now - (double)S.p_dts[p] * W.dt; }
                if (W.log_mode) { W.v_end_ts[vid] = t; W.v_end_kind[vid] = 2; }
            } else {
                const int q = S.p_q[p], ol = S.ol[q], slot = S.p_oslot[p];
                const double x_next = S.p_xn[p];
                if (W.no_cyclic) { int sn = W.l_start[ol]; W.v_visited[(size_t)vid * W.vis_words + (sn >> 5)] |= 1u << (sn & 31); }
                W.v_trav[vid] = S.p_trav[p] + 1; record_traversal(W, vid, S.p_trav[p], ol, t);
                W.v_link[vid] = ol;
                if (slot >= 0) {
                    long long o = S.o_roff[q];
                    W.X[(size_t)cur * C + o + slot] = x_next;
                    W.X[(size_t)(cur ^ 1) * C + o + slot] = S.p_xold[p];
                    W.V[o + slot] = S.p_v[p] + x_next / W.dt;
                    W.VID[o + slot] = vid;
                    W.LANE[o + slot] = S.p_olane[p];
                }
                W.v_entry_x[vid] = x_next;
                W.v_mr[vid] = 0.0;
            }
        }
        for (int ii = lane; ii < nin; ii += 32) {
            int il = S.il[ii];
            LR(W, il).cap_out_rem = S.co_rem[ii]; W.cum_dep[TLI(W, t, il)] = S.cumdep[ii];
            if (!publish) LC(W, il).pop_k2 = S.popped[ii];
            LC(W, il).ev_cnt = S.evc[ii]; LC(W, il).trips = S.trips[ii];
        }
        for (int q = lane; q < nol; q += 32) {
            int ol = S.ol[q];
            LR(W, ol).cap_in_rem = S.o_ci[q]; W.cum_arr[TLI(W, t, ol)] = S.o_cumarr[q]; LQ(W, ol).tail = S.o_tail[q];
        }
    }
    NLAP(g_node_prof, 5)
    NLAP_TOTAL(g_node_prof, 6)
    return publish;
}

// K0: the per-link and per-node scalar bookkeeping of a timestep, one THREAD per (link | node, replica):
// Link::update (traffi.cpp:542-565: capacity tokens, carry of the cumulative curves and of the instantaneous travel time to
// row t), reset of the per-step pop count, Node::signal_update (194-207) and Node::flow_capacity_update (226-234).  Nothing here
// depends on the order of links or nodes, and everything k_node reads of it is final before k_node starts -- inside k_node the
// same updates cost a warp per node with 4-9 of 32 lanes busy and a quarter of that kernel's stall samples.
// RL = true (ensembles): a warp is ONE entity in 32 replicas -- with the replica-minor layout every load and store of the warp
// is one contiguous row -- and on the steps that measure the mean speed (every `itt`-th) the thread also sums its own queue, in
// the oracle's fixed order (32 interleaved partial sums + butterfly, strided_sum32): the 32 replicas of a link have queues of
// similar length, so the lanes stay balanced, which a thread per link of ONE replica would not be.
// RL = false (few replicas): thread per entity, blockIdx.y = replica; the mean speed stays in k_node (warp per node).
UX_DEV double queue_vsum_serial(const double *V, int hd, int n, int cap) {
    double part[32];
#pragma unroll
    for (int k = 0; k < 32; k++) part[k] = 0.0;
    if (n <= 32) {
#pragma unroll
        for (int k = 0; k < 32; k++) if (k < n) part[k] = V[ring_slot(hd, k, cap)];
    } else {
        for (int i0 = 0; i0 < n; i0 += 32) {
#pragma unroll
            for (int k = 0; k < 32; k++) if (i0 + k < n) part[k] += V[ring_slot(hd, i0 + k, cap)];
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 32; k++) if (!(k & off)) part[k] = part[k] + part[k | off];  // == part[k] + part[k ^ off] of the butterfly, for the entries that reach part[0]
    }
    return part[0];
}

// 6 blocks of 256 threads per SM (<= 40 registers; the 32 partial sums of the mean-speed pass spill to L1): a streaming kernel wants the
// residency, measured on the B200 at 1024 replicas: 157 us (no bound, 94 registers) -> 100 (4) -> 96 (6)
#ifndef LPRE_MINB
#define LPRE_MINB 6
#endif
template <bool RL>
__global__ void __launch_bounds__(256, LPRE_MINB) k_link_pre(const DevWorld *worlds, int step_off, int R) {
    int idx = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    int r = RL ? (int)blockIdx.y * 32 + (idx & 31) : (int)blockIdx.y, i = RL ? idx >> 5 : idx;
    if (r >= R) return;
    const DevWorld &W = worlds[RL ? (int)blockIdx.y * 32 : r];  // block-uniform field loads; RL: the group's first replica + lane offsets
    const int ln = RL ? (idx & 31) : 0;
    int t = *W.t_base + step_off, L = W.L;
    if (i < L) {
        int l = i;
        LinkQ *lq = &LQ(W, l) + ln; LinkC *lc = &LC(W, l) + ln; LinkR *lr = &LR(W, l) + ln;
        if (t > 0) {
            size_t row = TLI(W, t, l) + ln, prev = TLI(W, t - 1, l) + ln;
            W.tt_inst[row] = W.tt_inst[prev];  // overwritten below / by k_node on the steps that measure the mean speed
            W.cum_arr[row] = W.cum_arr[prev];
            W.cum_dep[row] = W.cum_dep[prev];
        }
        double thr = W.dn * W.l_lanes[l];
        double ci = W.l_cap_in[l];
        if (ci < 10e9) { if (lr->cap_in_rem < thr) lr->cap_in_rem += ci * W.dt; } else lr->cap_in_rem = 10e9;
        double co = W.l_cap_out[l];
        if (co < 10e9) { if (lr->cap_out_rem < thr) lr->cap_out_rem += co * W.dt; } else lr->cap_out_rem = 10e9;
        lc->pop_k2 = 0;
        int tl = lq->tail;
        lq->tail_k1 = tl;  // k_node rewrites it for the out-links of nodes that generate
        if (RL && W.node_rl && (t % W.itt) == 0) {  // Link::set_travel_time (traffi.cpp:599-622); with the warp-per-node kernel it stays there
            int hd = lq->head[t & 1], nq = tl - hd;
            double tt;
            if (nq > 0) {
                const double *V = (const double *)((const char *)W.V + (size_t)ln * (size_t)W.rep_stride) + W.l_roff[l];
                double avg = queue_vsum_serial(V, hd, nq, W.l_rcap[l]) / (double)nq;
                tt = avg > 0.0 ? W.l_length[l] / avg : W.l_length[l] / (W.l_vmax[l] / 100.0);
            } else tt = W.l_length[l] / W.l_vmax[l];
            W.tt_inst[TLI(W, t, l) + ln] = tt;
        }
    } else if (i < L + W.N) {
        node_signal_capacity(W, i - L, t, ln);
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
// MINB = resident blocks per SM the compiler has to leave room for: 8 (64 registers) where residency pays -- ensembles, large worlds -- and 2
// (~100 registers, no spills) when the whole launch is one wave of blocks anyway and only the chain of the busiest node counts (B200, Chicago
// sketch, one scenario: 18.1 -> 17.0 ms; at 8 replicas the same setting loses)
template <int MC, int MD, int MINB = 32 / XFER_WARPS>
__global__ void __launch_bounds__(32 * XFER_WARPS, MINB) k_node(const DevWorld *worlds, int step_off) {
    // replica = blockIdx.x (fastest in launch order): the blocks resident together are the same nodes of different replicas
    // -- same code path, same static tables -- and every replica's dependency chains start in the first wave
#ifdef UX_PROFILE_NODE
    const long long kstart_ = clock64();
#endif
    LOAD_WORLD_N(W, worlds, (XFER_WARPS > 1 ? 64 : 32), blockIdx.x)
    WARP_BEGIN
    __shared__ NodeSmem<MC, MD> sm_all[XFER_WARPS];
    __shared__ double tt_all[XFER_WARPS][MD];
    int warp = (int)((blockIdx.y * blockDim.x + threadIdx.x) >> 5), wib = (int)((threadIdx.x >> 5) % XFER_WARPS);
    if (warp >= W.N) return;
    int n = W.lvl_nodes[warp], t = *W.t_base + step_off;
    int lvl_raw = W.n_levels > 1 ? W.lvl_level[warp] : 0, lvl = lvl_raw & 0xffff;  // bit 16: some node waits for this one
    pre_node(W, sm_all[wib].pre, tt_all[wib], n, t, (int)(threadIdx.x & 31));
#ifndef UX_EMU
    __threadfence_block();
#endif
    XferSmem<MC, MD> &S = sm_all[wib].xf;
    const bool publish = (lvl_raw >> 16) != 0;  // publish completion only where somebody may be waiting: the fence is not free
    if (!xfer_node(W, S, n, t, lvl > 0, publish) && publish && NS(W, n).node_act == t) {  // (waiters skip nodes without flagged head positions)
        WARP_SYNC();
        WARP_FOR(lane) if (lane == 0) { __threadfence(); *((volatile int *)&NDONE(W, n)) = t + 1; }
    }
#ifdef UX_PROFILE_NODE
    if ((threadIdx.x & 31) == 0) atomicMax(&g_node_prof[t & 4095][7], (unsigned long long)(clock64() - kstart_));
#endif
}

// ---------------------------------------------------------------------------------------------------
// k_node_rl -- the node step in REPLICA-LANE form (ensembles): one THREAD per (node, replica); a warp is one node in 32 seeded
// replicas.  k_node spends a warp on a node whose transfer has one to three candidates (9-12 of 32 lanes busy, ~2 000 warp
// instructions per node and step, shared-memory staging that caps residency); here every lane walks the reference's serial
// algorithm (Node::generate traffi.cpp:105-189, Node::transfer 239-445) for its own replica.  The static tables of the node are
// warp-uniform loads, the per-link / per-node scalars of the 32 replicas are contiguous rows (replica-minor layout), and the
// ring / per-platoon arrays of replica r are replica 0's + r * rep_stride -- no pointer table per lane.  Order-free bookkeeping
// and the mean-speed travel times are done by k_link_pre<true> before this kernel.  Nothing is staged: a thread re-reads what
// it wrote itself (ring slots it pushed, counters it advanced) from global memory, which is coherent for a single thread.
// Dependency levels (links short enough that the id-ordered visibility of pops matters) work as in k_node: warps are ordered by
// (level, id), a lane waits for the completion flag of exactly the lower-level nodes it depends on, in its own replica.
// Random draws are the same counters as in k_node / the oracle, so the results are bit-identical.
// ---------------------------------------------------------------------------------------------------
#ifndef NRL_WARPS
#define NRL_WARPS 4
#endif
#define NRL_HEAVY_WARPS (NRL_WARPS < 4 ? NRL_WARPS : 4)
// 4 blocks per SM: more residency LOSES here (B200, x1024: 878 us at 3-4 blocks, 945 at 5, 1025 at 6, 1069 at 8) -- the threads' local
// arrays (the gathered neighbourhood of the node) already fill the L1, and more threads push them out to the L2
#ifndef NRL_MINB
#define NRL_MINB 4
#endif
#define RLP(T_, ptr_) ((T_ *)((char *)(ptr_) + ro))  // the lane's replica-major array

// W = the staged world of the lane group's first replica (the lanes add their offsets), Wr = the lane's own pointer table, for
// the rare helpers only (route choice, error word, traversal log)
template <int MC, int MD>
UX_DEV void node_step_rl(const DevWorld &W, const DevWorld &Wr, const int ln, const int n, const int t, const int lvl_raw) {
    const size_t ro = (size_t)ln * (size_t)W.rep_stride;
    const int cur = t & 1;
    const size_t C = (size_t)W.C;
    double *Xc = RLP(double, W.X) + (size_t)cur * C, *Xo = RLP(double, W.X) + (size_t)(cur ^ 1) * C, *V = RLP(double, W.V);
    int *VID = RLP(int, W.VID), *LANE = RLP(int, W.LANE);
    const int o0 = W.n_out_off[n], nout = W.n_out_off[n + 1] - o0;
    const int i0 = W.n_in_off[n], nin = W.n_in_off[n + 1] - i0;
    NodeS *ns = &NS(W, n) + ln; NodeR *nr = &NR(W, n) + ln;

    NPROF_T(pt0)
    // ---- release HOME -> WAIT (traffi.cpp:836-841): the per-origin list is sorted by departure step
    const int g0 = W.gen_off[n], glen = W.gen_off[n + 1] - g0;
    int av = ns->gen_avail, gh = ns->gen_head;
    if (av < glen) {
        int av0 = av;
        while (av < glen && W.gen_ts[g0 + av] <= t - 1) av++;
        if (av != av0) ns->gen_avail = av;
    }
    // ---- generate (traffi.cpp:105-189): attempts in order, the first refusal ends them
    int natt = av - gh;
    { int tl_ = W.n_out_lanes[n]; if (natt > tl_) natt = tl_; }
    if (nout == 0) natt = 0;
    if (natt > 0) {
        const int gh0 = gh;
        for (int i = 0; i < natt; i++) {
            int vid = W.gen_list[g0 + gh], e = 0;
            int ol = route_choice(W, Wr, vid, n, t, (unsigned)n, UX_RNG_GENERATE, (unsigned)i, &e);
            if (e) { dev_error(Wr, e, vid); break; }
            RLP(int, W.v_next)[vid] = ol;
            if (ol < 0) break;
            LinkQ *lq = &LQ(W, ol) + ln; LinkR *lr = &LR(W, ol) + ln;
            int hd = lq->head[cur], tl = lq->tail, lanes = W.l_lanes[ol], cap = W.l_rcap[ol], sz = tl - hd;
            long long o = W.l_roff[ol];
            bool ok = false;  // traffi.cpp:130-142
            if (sz < lanes) ok = true;
            else if (Xc[o + ring_slot(hd, sz - lanes, cap)] > W.l_dpl_dn[ol]) ok = true;
            double ci = lr->cap_in_rem;
            if (ci < W.dn) ok = false;
            if (!ok) break;
            if (sz >= cap) { dev_error(Wr, DERR_RING_FULL, ol); break; }
            gh++;
            int lane_new = sz > 0 ? (LANE[o + ring_slot(tl - 1, 0, cap)] + 1) % lanes : 0;
            int slot = ring_slot(tl, 0, cap);
            Xc[o + slot] = 0.0; Xo[o + slot] = 0.0; V[o + slot] = W.l_vmax[ol]; VID[o + slot] = vid; LANE[o + slot] = lane_new;
            lq->tail = tl + 1;
            RLP(int, W.v_state)[vid] = UX_VS_RUN;
            if (W.log_mode) RLP(int, W.v_gen_ts)[vid] = t;
            RLP(double, W.v_atl)[vid] = (double)t * W.dt;
            RLP(double, W.v_entry_x)[vid] = 0.0;
            if (W.no_cyclic) { int sn = W.l_start[ol]; RLP(unsigned, W.v_visited)[(size_t)vid * W.vis_words + (sn >> 5)] |= 1u << (sn & 31); }  // mark_entered (160-166)
            { int *v_trav = RLP(int, W.v_trav); int k = v_trav[vid]; v_trav[vid] = k + 1; RLP(int, W.v_link)[vid] = ol; if (W.tv_count) record_traversal(Wr, vid, k, ol, t); }
            W.cum_arr[TLI(W, t, ol) + ln] += W.dn;
            lr->cap_in_rem = ci - W.dn;
        }
        if (gh != gh0) {
            ns->gen_head = gh;
            for (int k = 0; k < nout; k++) { LinkQ *lq = &LQ(W, W.n_out[o0 + k]) + ln; lq->tail_k1 = lq->tail; }  // tail after generation (k_link_pre copied the unchanged ones)
        }
    }

    // ---- transfer (traffi.cpp:239-445).  What bounds this kernel is the chain of dependent memory round trips of the busiest
    //      node, so everything the serial slot loop reads is fetched first, in batches of independent loads (registers, then
    //      thread-local arrays), and the loop itself runs on local data: global memory only sees its stores.
    NPROF_T(pt1) NPROF_MAX(0, pt0, pt1)
    const bool act = ns->node_act == t && nin > 0;
    if (act) {
        const int nsig = W.n_sig_off[n + 1] - W.n_sig_off[n], phase = ns->sig_phase;
        int i_hd[MD], i_hc[MD], i_evc[MD], i_trips[MD], i_poff[MD + 1];
        unsigned long long pops0 = 0, pops1 = 0;  // platoons popped from in-link ii so far: 8 bits each, in registers (read by every eligibility test)
#define I_POP(ii_) ((int)(((ii_) < 8 ? pops0 >> ((ii_) * 8) : pops1 >> (((ii_) - 8) * 8)) & 255ull))
#define I_POP_INC(ii_) { if ((ii_) < 8) pops0 += 1ull << ((ii_) * 8); else pops1 += 1ull << (((ii_) - 8) * 8); }
        double i_co[MD], i_cumdep[MD], i_mp[MD];
        unsigned sig_ok = 0, i_dirty = 0, co_ok = 0;  // co_ok: in-links with outflow tokens left
        int npos = 0;
        // A: in-link records, four links per batch (two 128-bit records + tokens + curve value each)
        for (int b = 0; b < nin; b += 4) {
            int il_[4]; int4 q_[4], c_[4]; double co_[4], cd_[4];
#pragma unroll
            for (int u = 0; u < 4; u++) il_[u] = b + u < nin ? W.n_in[i0 + b + u] : -1;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int il = il_[u] < 0 ? il_[0] : il_[u];
                q_[u] = *reinterpret_cast<const int4 *>(&LQ(W, il) + ln); c_[u] = *reinterpret_cast<const int4 *>(&LC(W, il) + ln);
                co_[u] = (&LR(W, il) + ln)->cap_out_rem; cd_[u] = W.cum_dep[TLI(W, t, il) + ln];
            }
#pragma unroll
            for (int u = 0; u < 4; u++) if (il_[u] >= 0) {
                int ii = b + u, hc = c_[u].y;  // LinkQ {head[0], head[1], tail, tail_k1}, LinkC {pop_k2, hcount, trips, ev_cnt}
                i_hd[ii] = cur ? q_[u].y : q_[u].x; i_hc[ii] = hc; i_trips[ii] = c_[u].z; i_evc[ii] = c_[u].w;
                i_co[ii] = co_[u]; i_cumdep[ii] = cd_[u]; i_poff[ii] = npos; i_mp[ii] = W.l_mp[il_[u]];
                if (co_[u] >= W.dn) co_ok |= 1u << ii;
                if (npos + hc > MC) hc = MC - npos;
                npos += hc;
            }
        }
        i_poff[nin] = npos;
        if (npos > 0) {
            // B: the platoons at the flagged head positions: ids (eight per batch), then flags and next links (four per batch)
            int p_vid[MC], p_next[MC]; unsigned char p_flag[MC], p_ii[MC];
            { int ii = 0; for (int p = 0; p < npos; p++) { while (i_poff[ii + 1] <= p) ii++; p_ii[p] = (unsigned char)ii; } }
            for (int b = 0; b < npos; b += 8) {
                int v_[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    int p = b + u < npos ? b + u : b, ii = p_ii[p], il = W.n_in[i0 + ii];
                    v_[u] = VID[W.l_roff[il] + ring_slot(i_hd[ii], p - i_poff[ii], W.l_rcap[il])];
                }
#pragma unroll
                for (int u = 0; u < 8; u++) if (b + u < npos) p_vid[b + u] = v_[u];
            }
            const unsigned char *v_flags = RLP(unsigned char, W.v_flags); const int *v_next = RLP(int, W.v_next);
            bool any_wait = false;
            unsigned char c_idx[MC]; int nc = 0;  // candidate positions sorted by platoon id (250-253)
            for (int b = 0; b < npos; b += 4) {
                unsigned char f_[4]; int x_[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { int vid = p_vid[b + u < npos ? b + u : b]; f_[u] = v_flags[vid]; x_[u] = v_next[vid]; }
#pragma unroll
                for (int u = 0; u < 4; u++) if (b + u < npos) {
                    int p = b + u;
                    p_flag[p] = f_[u]; p_next[p] = x_[u];
                    if (f_[u] & VF_WAIT_END) any_wait = true;
                    if (f_[u] & VF_CAND) {
                        int k = nc++, vid = p_vid[p];
                        while (k > 0 && p_vid[c_idx[k - 1]] > vid) { c_idx[k] = c_idx[k - 1]; k--; }
                        c_idx[k] = (unsigned char)p;
                    }
                }
            }
            if (nc > 0 || any_wait) {
                if (nsig > 1) {
                    for (int ii = 0; ii < nin; ii++) if (i_hc[ii] > 0) { int il = W.n_in[i0 + ii], s0 = W.l_sg_off[il], gl = W.l_sg_off[il + 1] - s0; if (list_has(W.l_sg + s0, gl, phase)) sig_ok |= 1u << ii; }
                } else sig_ok = 0xffffffffu;
                // requested out-links in order of first appearance, each `lanes` times (260-273)
                int o_id[MD], o_hd[MD], o_tail[MD], o_tail0[MD], o_lastlane[MD], o_lanes[MD], o_woff[MD + 1]; double o_ci[MD], o_cumarr[MD];
                unsigned char expd[MC], p_q[MC];  // p_q: per head position, the index of the out-link it wants to enter (255: none / not a candidate)
                int nol = 0, nexp = 0;
                for (int p = 0; p < npos; p++) p_q[p] = 255;
                for (int k = 0; k < nc; k++) {
                    int p = c_idx[k], nl = p_next[p], q = 255;
                    if (nl >= 0) {
                        for (int x = 0; x < nol; x++) if (o_id[x] == nl) { q = x; break; }
                        if (q == 255 && nol < MD) { q = nol; o_id[nol++] = nl; }
                    }
                    p_q[p] = (unsigned char)q;
                }
                { int acc = 0; for (int q = 0; q < nol; q++) { int lanes = W.l_lanes[o_id[q]]; o_lanes[q] = lanes; o_woff[q] = acc; acc += lanes; for (int x = 0; x < lanes && nexp < MC; x++) expd[nexp++] = (unsigned char)q; } o_woff[nol] = acc; }
                const int nshuf = (!W.hard_det && nexp > 1) ? nexp - 1 : 0;
                if (!W.hard_det) {  // Fisher-Yates (276-282); step i uses draw number nexp-1-i
                    for (int i = nexp - 1; i > 0; i--) {
                        int j = ux_rng_below(W.seed + ln, (uint32_t)t, (uint32_t)n, UX_RNG_TRANSFER, (uint32_t)(nexp - 1 - i), i + 1);
                        unsigned char tmp = expd[i]; expd[i] = expd[j]; expd[j] = tmp;
                    }
                }
                NPROF_T(pt2) NPROF_MAX(1, pt1, pt2)
                // ---- dependency wait: only now are the pops of lower-level end nodes needed
                if ((lvl_raw & 0xffff) > 0) {
                    for (int q = 0; q < nol; q++) {  // per requested out-link: its end node can pop only flagged head positions (hcount > 0)
                        const int ol = o_id[q];
                        if (!(W.l_flags[ol] & LF_SEES_POPS)) continue;
                        const int hc_ = (&LC(W, ol) + ln)->hcount;
                        if (hc_ == 0) continue;
                        { const LinkQ *lq = &LQ(W, ol) + ln; if (lq->tail - lq->head[cur] - hc_ >= W.l_lanes[ol]) continue; }  // keeps `lanes` platoons whatever leaves: looks the same from its tail
                        volatile int *done = &NDONE(W, W.l_end[ol]) + ln;
#ifdef UX_EMU
                        if (*done < t + 1) dev_error(Wr, DERR_DEP_TIMEOUT, n);
#else
                        long long spins = 0;
                        while (*done < t + 1) { __nanosleep(64); if (++spins > UX_DEP_SPIN_BUDGET) { dev_error(Wr, DERR_DEP_TIMEOUT, n); break; } }
#endif
                    }
#ifndef UX_EMU
                    __threadfence();
#endif
                }
                NPROF_T(pt3) NPROF_MAX(2, pt2, pt3)
                // C: out-link records (four per batch), then the lane of the last platoon and the window of the last `lanes` ones
                for (int b = 0; b < nol; b += 4) {
                    int4 q_[4]; double ci_[4], ca_[4]; int pk_[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        int ol = o_id[b + u < nol ? b + u : b];
                        q_[u] = *reinterpret_cast<const int4 *>(&LQ(W, ol) + ln); ci_[u] = (&LR(W, ol) + ln)->cap_in_rem; ca_[u] = W.cum_arr[TLI(W, t, ol) + ln];
                        pk_[u] = 0;
                        if (W.l_flags[ol] & LF_SEES_POPS) { const int hc_ = (&LC(W, ol) + ln)->hcount; if (hc_ > 0 && q_[u].z - (cur ? q_[u].y : q_[u].x) - hc_ < W.l_lanes[ol]) pk_[u] = ld_cg(&(&LC(W, ol) + ln)->pop_k2); }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) if (b + u < nol) {
                        int q = b + u;
                        o_hd[q] = (cur ? q_[u].y : q_[u].x) + pk_[u]; o_tail[q] = o_tail0[q] = q_[u].z; o_ci[q] = ci_[u]; o_cumarr[q] = ca_[u];
                    }
                }
                double w_x[MC], w_xo[MC];  // window of the last `lanes` platoons of each out-link, oldest first (as many as there are)
                {
                    const int wtot = o_woff[nol] < MC ? o_woff[nol] : MC;
                    for (int b = 0; b < wtot; b += 8) {
                        double a_[8], o_[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            int e = b + u < wtot ? b + u : b, q = 0;
                            while (o_woff[q + 1] <= e) q++;
                            int k = e - o_woff[q], sz = o_tail[q] - o_hd[q], wn = sz < o_lanes[q] ? sz : o_lanes[q];
                            a_[u] = 0.0; o_[u] = 0.0;
                            if (k < wn) {
                                size_t at = (size_t)W.l_roff[o_id[q]] + ring_slot(o_hd[q], sz - wn + k, W.l_rcap[o_id[q]]);
                                a_[u] = Xc[at]; o_[u] = Xo[at];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++) if (b + u < wtot) { w_x[b + u] = a_[u]; w_xo[b + u] = o_[u]; }
                    }
                    for (int b = 0; b < nol; b += 4) {
                        int l_[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            int q = b + u < nol ? b + u : b, sz = o_tail[q] - o_hd[q];
                            l_[u] = sz > 0 ? LANE[W.l_roff[o_id[q]] + ring_slot(o_tail[q] - 1, 0, W.l_rcap[o_id[q]])] : -1;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) if (b + u < nol) o_lastlane[b + u] = l_[u];
                    }
                }
                unsigned char w_n[MD], w_f[MD];  // platoons in the window; once it is full (`lanes` entries) it is circular and w_f is its oldest entry
                for (int q = 0; q < nol; q++) w_f[q] = 0;
                for (int q = 0; q < nol; q++) { int sz = o_tail[q] - o_hd[q]; w_n[q] = (unsigned char)(sz < o_lanes[q] ? sz : o_lanes[q]); }
                NPROF_T(pt4) NPROF_MAX(3, pt3, pt4)
                const double fc = W.n_fc[n];
                double fc_rem = nr->fc_rem;
                int n_picks = 0;
                // per in-link: the out-link its front platoon may move onto now (255: none) -- a byte each, in two registers
                unsigned long long fq0 = ~0ull, fq1 = ~0ull;
#define F_SET(ii_, v_) { if ((ii_) < 8) fq0 = (fq0 & ~(255ull << ((ii_) * 8))) | ((unsigned long long)(v_) << ((ii_) * 8)); else fq1 = (fq1 & ~(255ull << (((ii_) - 8) * 8))) | ((unsigned long long)(v_) << (((ii_) - 8) * 8)); }
                auto front_q = [&](int ii) -> int {  // front of its link, outflow tokens, green (305-330)
                    const int pop = I_POP(ii);
                    if (pop >= i_hc[ii] || !(((sig_ok & co_ok) >> ii) & 1u)) return 255;
                    const int fp = i_poff[ii] + pop;
                    return fp < npos ? (int)p_q[fp] : 255;
                };
                for (int ii = 0; ii < nin; ii++) if (i_hc[ii] > 0) { const int f = front_q(ii); F_SET(ii, f) }
                unsigned good = 0;  // out-links that accept a platoon now (285-297; the node capacity is tested once per slot)
                auto accepts = [&](int q) -> bool {
                    const int sz = o_tail[q] - o_hd[q];
                    bool ok = false;
                    if (sz < o_lanes[q]) ok = true;
                    else if (w_x[o_woff[q] + w_f[q]] > W.l_dpl_dn[o_id[q]]) ok = true;
                    if (o_ci[q] < W.dn) ok = false;
                    return ok;
                };
                for (int q = 0; q < nol; q++) if (accepts(q)) good |= 1u << q;
                // trip end of the front platoon (position p) of in-link ii: Vehicle::end_trip (886-927)
                auto end_front = [&](int ii, int p) {
                    const int il = W.n_in[i0 + ii], vid = p_vid[p];
                    const double atl = RLP(double, W.v_atl)[vid], ex = RLP(double, W.v_entry_x)[vid], di = RLP(double, W.v_dist)[vid];
                    const double x = Xc[W.l_roff[il] + ring_slot(i_hd[ii], p - i_poff[ii], W.l_rcap[il])];
                    i_trips[ii] += 1; i_cumdep[ii] += W.dn; i_dirty |= 1u << ii;
                    int s = (int)(atl / W.dt);
                    if (s < 0) s = 0;
                    if (s >= W.T) s = W.T - 1;
                    int ord = ++i_evc[ii];
                    W.ev_ord[TLI(W, s, il) + ln] = ord;
                    W.ev_tt[TLI(W, s, il) + ln] = ((double)t + 1.0) * W.dt - atl;
                    RLP(double, W.v_atl)[vid] = (double)t * W.dt;
                    RLP(double, W.v_dist)[vid] = di + (x - ex);
                    RLP(int, W.v_link)[vid] = -1;
                    if (p_flag[p] & VF_ABORT) { RLP(int, W.v_state)[vid] = UX_VS_ABORT; RLP(int, W.v_arr_ts)[vid] = -1; RLP(double, W.v_tt)[vid] = -1.0; }
                    else { RLP(int, W.v_state)[vid] = UX_VS_END; RLP(int, W.v_arr_ts)[vid] = t; RLP(double, W.v_tt)[vid] = (double)t * W.dt - (double)W.v_depart_ts[vid] * W.dt; }
                    if (W.log_mode) { RLP(int, W.v_end_ts)[vid] = t; RLP(unsigned char, W.v_end_kind)[vid] = 2; }
                    RLP(unsigned char, W.v_flags)[vid] = 0;
                    I_POP_INC(ii)
                };
                for (int s = 0; s < nexp; s++) {
                    if (fc_rem < W.dn) break;  // no slot can move anything any more (298-300)
                    const int q = expd[s];
                    if (!((good >> q) & 1u)) continue;
                    // in-links whose front wants this out-link: byte-wise compare of the packed fronts with q
                    const unsigned long long pat = 0x0101010101010101ull * (unsigned long long)q;
                    const unsigned long long x0 = fq0 ^ pat, x1 = fq1 ^ pat;
                    const unsigned long long z0 = (x0 - 0x0101010101010101ull) & ~x0 & 0x8080808080808080ull, z1 = (x1 - 0x0101010101010101ull) & ~x1 & 0x8080808080808080ull;
                    if (!(z0 | z1)) continue;
                    int ii;
                    if (!(z0 & (z0 - 1)) && !z1) { ii = (__ffsll((long long)z0) - 1) >> 3; n_picks++; }  // one offer (the usual case): no lottery, its draw is consumed
                    else if (!z0 && !(z1 & (z1 - 1))) { ii = 8 + ((__ffsll((long long)z1) - 1) >> 3); n_picks++; }
                    else {
                        // several in-links offer a platoon: candidates in ascending platoon id, random_choice over their merge priorities (utils.h:62-87)
                        // (the borrow of the zero-byte test can flag a byte above a true match: every flagged byte is checked)
                        int cand[MD]; int nm = 0;
                        for (int i2 = 0; i2 < nin; i2++) {
                            const int f = (int)(((i2 < 8 ? fq0 >> (i2 * 8) : fq1 >> ((i2 - 8) * 8))) & 255ull);
                            if (f != q) continue;
                            const int vid2 = p_vid[i_poff[i2] + I_POP(i2)];
                            int at = nm++;
                            while (at > 0 && p_vid[i_poff[cand[at - 1]] + I_POP(cand[at - 1])] > vid2) { cand[at] = cand[at - 1]; at--; }
                            cand[at] = i2;
                        }
                        int pick = 0;
                        if (W.hard_det) { for (int i = 1; i < nm; i++) if (i_mp[cand[i]] > i_mp[cand[pick]]) pick = i; }
                        else {
                            const double u = ux_rng_uniform(W.seed + ln, (uint32_t)t, (uint32_t)n, UX_RNG_TRANSFER, (uint32_t)(nshuf + n_picks));
                            n_picks++;
                            double wsum = 0.0;
                            for (int i = 0; i < nm; i++) wsum += i_mp[cand[i]];
                            if (wsum <= 0.0) { pick = (int)(u * (double)nm); if (pick >= nm) pick = nm - 1; }
                            else { const double rr = u * wsum; double accum = 0.0; pick = nm - 1; for (int i = 0; i < nm; i++) { accum += i_mp[cand[i]]; if (rr <= accum) { pick = i; break; } } }
                        }
                        ii = cand[pick];
                    }
                    // ---- move the front platoon of in-link ii onto out-link q: everything it needs in one batch of loads, then stores only
                    const int p = i_poff[ii] + I_POP(ii), vid = p_vid[p], il = W.n_in[i0 + ii], ol = o_id[q];
                    const int lanes_ol = o_lanes[q], sz = o_tail[q] - o_hd[q], w0 = o_woff[q];
                    const size_t in_at = (size_t)W.l_roff[il] + ring_slot(i_hd[ii], p - i_poff[ii], W.l_rcap[il]);
                    const double atl = RLP(double, W.v_atl)[vid], px = Xc[in_at], pxold = Xo[in_at], pv = V[in_at], ex = RLP(double, W.v_entry_x)[vid],
                                 di = RLP(double, W.v_dist)[vid], mr = RLP(double, W.v_mr)[vid];
                    const int trav = RLP(int, W.v_trav)[vid];
                    const double vm_ol = W.l_vmax[ol], vm_il = W.l_vmax[il], len_ol = W.l_length[ol], dpl = W.l_dpl_dn[ol];
                    i_co[ii] -= W.dn; o_ci[q] -= W.dn; i_dirty |= 1u << ii;
                    if (i_co[ii] < W.dn) co_ok &= ~(1u << ii);
                    if (fc >= 0.0) fc_rem -= W.dn;
                    i_cumdep[ii] += W.dn; o_cumarr[q] += W.dn;
                    {   // record_tt_event (357-362)
                        int es = (int)(atl / W.dt);
                        if (es < 0) es = 0;
                        if (es >= W.T) es = W.T - 1;
                        int ord = ++i_evc[ii];
                        W.ev_ord[TLI(W, es, il) + ln] = ord;
                        W.ev_tt[TLI(W, es, il) + ln] = (double)t * W.dt - atl;
                    }
                    RLP(double, W.v_atl)[vid] = (double)t * W.dt;
                    RLP(double, W.v_dist)[vid] = di + (px - ex);
                    I_POP_INC(ii)
                    if (W.no_cyclic) { int sn = W.l_start[ol]; RLP(unsigned, W.v_visited)[(size_t)vid * W.vis_words + (sn >> 5)] |= 1u << (sn & 31); }
                    RLP(int, W.v_trav)[vid] = trav + 1; if (W.tv_count) record_traversal(Wr, vid, trav, ol, t);
                    RLP(int, W.v_link)[vid] = ol;
                    RLP(unsigned char, W.v_flags)[vid] = 0;
                    // carry the residual move onto the new link (400-419)
                    double x_next = mr * vm_ol / vm_il;
                    if (sz >= lanes_ol) {
                        double x_cong = w_xo[w0 + w_f[q]] - dpl;
                        if (x_cong < 0.0) x_cong = 0.0;
                        if (x_next > x_cong) x_next = x_cong;
                    }
                    if (x_next >= len_ol) x_next = len_ol;
                    if (x_next < 0.0) x_next = 0.0;
                    if (sz >= W.l_rcap[ol]) dev_error(Wr, DERR_RING_FULL, ol);
                    else {  // push (lane rule 385-390)
                        int lane_new = sz > 0 ? o_lastlane[q] + 1 : 0;  // (lane + 1) % lanes: lane labels are < lanes
                        if (lane_new >= lanes_ol) lane_new = 0;
                        size_t at = (size_t)W.l_roff[ol] + ring_slot(o_tail[q], 0, W.l_rcap[ol]);
                        Xc[at] = x_next; Xo[at] = pxold; V[at] = pv + x_next / W.dt; VID[at] = vid; LANE[at] = lane_new;
                        o_tail[q] += 1; o_lastlane[q] = lane_new;
                        if (w_n[q] < lanes_ol) { w_x[w0 + w_n[q]] = x_next; w_xo[w0 + w_n[q]] = pxold; w_n[q] += 1; }
                        else { int e = w0 + w_f[q]; w_x[e] = x_next; w_xo[e] = pxold; w_f[q] = (unsigned char)(w_f[q] + 1 == lanes_ol ? 0 : w_f[q] + 1); }
                    }
                    RLP(double, W.v_entry_x)[vid] = x_next;
                    RLP(double, W.v_mr)[vid] = 0.0;
                    // the new front of the in-link may be waiting for its trip end (421-424)
                    if (I_POP(ii) < i_hc[ii]) { int fp = i_poff[ii] + I_POP(ii); if (fp < npos && (p_flag[fp] & VF_WAIT_END)) end_front(ii, fp); }
                    // what this move changed for the slots to come: the in-link's front and tokens, the out-link's tail
                    { const int f = front_q(ii); F_SET(ii, f) }
                    if (accepts(q)) good |= 1u << q; else good &= ~(1u << q);
                }
#undef F_SET
                NPROF_T(pt5) NPROF_MAX(4, pt4, pt5)
                // flush trip-end-waiting fronts (432-441)
                for (int ii = 0; ii < nin; ii++) {
                    int hc = i_hc[ii];
                    for (int x = 0; x < hc; x++) {
                        if (I_POP(ii) >= hc) break;
                        int fp = i_poff[ii] + I_POP(ii);
                        if (fp >= npos || !(p_flag[fp] & VF_WAIT_END)) break;
                        end_front(ii, fp);
                    }
                }
                // write the per-link accumulators back
                for (int ii = 0; ii < nin; ii++) if ((i_dirty >> ii) & 1u) {
                    int il = W.n_in[i0 + ii];
                    LinkC *lc = &LC(W, il) + ln;
                    (&LR(W, il) + ln)->cap_out_rem = i_co[ii]; W.cum_dep[TLI(W, t, il) + ln] = i_cumdep[ii];
                    lc->pop_k2 = I_POP(ii); lc->ev_cnt = i_evc[ii]; lc->trips = i_trips[ii];
                }
                for (int q = 0; q < nol; q++) if (o_tail[q] != o_tail0[q]) {
                    int ol = o_id[q];
                    (&LR(W, ol) + ln)->cap_in_rem = o_ci[q]; W.cum_arr[TLI(W, t, ol) + ln] = o_cumarr[q]; (&LQ(W, ol) + ln)->tail = o_tail[q];
                }
                nr->fc_rem = fc_rem;
                NPROF_T(pt6) NPROF_MAX(5, pt5, pt6) NPROF_MAX(6, pt0, pt6)
            }
        }
    }
    if ((lvl_raw & 0x10000) && ns->node_act == t) {  // publish completion only where somebody may be waiting (waiters skip idle nodes)
#ifndef UX_EMU
        __threadfence();
#endif
        *((volatile int *)(&NDONE(W, n) + ln)) = t + 1;
    }
}

#undef I_POP
#undef I_POP_INC
// k_node_h -- the node step of an ensemble as ONE launch over a table of blocks, each in one of two forms:
//   form 1 (replica lanes): four nodes x 32 replicas, node_step_rl above -- all light nodes;
//   form 0 (warp per node): four nodes of one replica, pre_node + xfer_node of k_node -- the heavy nodes (many lanes in or out:
//   a single thread walking 30 slots x 20 candidates is slower than a warp that scans them in parallel on shared memory).
// The table is ordered by dependency level (heavy blocks of a level, then its light blocks), so a waiting lane or warp only
// ever waits for blocks with a smaller index, as in k_node.
// HEAVY = false: the table holds form-1 blocks only (no shared-memory staging is declared, the L1 keeps the threads' local arrays).
template <int MC, int MD, bool HEAVY>
__global__ void __launch_bounds__(32 * NRL_WARPS, NRL_MINB) k_node_h(const DevWorld *worlds, int step_off, int R, const int4 *blks) {
    const int4 b = blks[blockIdx.x];  // {form, first position in lvl_nodes, nodes in this block, replica (form 0) | lane group (form 1)}
    LOAD_WORLD_N(W, worlds, (NRL_WARPS > 1 ? 64 : 32), b.x ? b.w * 32 : b.w)
    const int wib = (int)(threadIdx.x >> 5), ln = (int)(threadIdx.x & 31);
    if (wib >= b.z) return;
    const int pos = b.y + wib, n = W.lvl_nodes[pos], t = *W.t_base + step_off;
    const int lvl_raw = W.lvl_level[pos];
    if (b.x) {
        const int r = b.w * 32 + ln;
        if (r < R) node_step_rl<MC, MD>(W, worlds[r], ln, n, t, lvl_raw);
        return;
    }
    if constexpr (HEAVY) {
        __shared__ NodeSmem<MC, MD> sm_all[NRL_HEAVY_WARPS];  // (the host puts at most NRL_HEAVY_WARPS nodes into a warp-form block)
        __shared__ double tt_all[NRL_HEAVY_WARPS][MD];
        WARP_BEGIN
        pre_node(W, sm_all[wib].pre, tt_all[wib], n, t, ln);
#ifndef UX_EMU
        __threadfence_block();
#endif
        const bool publish = (lvl_raw & 0x10000) != 0;
        if (!xfer_node(W, sm_all[wib].xf, n, t, (lvl_raw & 0xffff) > 0, publish) && publish && NS(W, n).node_act == t) {
            WARP_SYNC();
            WARP_FOR(lane) if (lane == 0) { __threadfence(); *((volatile int *)&NDONE(W, n)) = t + 1; }
        }
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
// A work item gets UPD_GROUP lanes (template parameter).  With 16 a warp carries two (link, part) items side by side (a link head has 1-13
// positions and most queues a handful of platoons); the halves never exchange anything and every synchronisation is a
// __syncwarp over the half's own lane mask.  Measured on the B200 (parts of 16*UPD_GROUP ring slots): C4 69.2 -> 58.1 us and
// Chicago x32 56.7 -> 54.7 us, but long queues and single scenarios lose (Chicago dn=5 x16 47.8 -> 54.5, x1 16.5 -> 22.9; Sioux
// Falls x64 15.0 -> 18.3; Chicago x1 12.6 -> 16.0), so the host picks 16 only for large, lightly loaded worlds (ux_world::upd_group).
#define UPD_GROUPS_PER_BLOCK(G) (UPD_WARPS * 32 / (G))
#ifdef UX_EMU
#define UPD_BEGIN if (threadIdx.x & (UPD_GROUP - 1)) return;
#define UPD_FOR(lane) for (int lane = 0; lane < UPD_GROUP; lane++)
#define UPD_SYNC()
#else
#define UPD_BEGIN const unsigned upd_mask_ = UPD_GROUP == 32 ? 0xffffffffu : ((0xffffffffu >> (32 - UPD_GROUP)) << ((threadIdx.x & 31u) & ~(unsigned)(UPD_GROUP - 1)));
#define UPD_FOR(lane) for (int lane = (int)(threadIdx.x & (UPD_GROUP - 1)), once_ = 1; once_; once_ = 0)
#define UPD_SYNC() __syncwarp(upd_mask_)
#endif
#ifndef K_UPDATE_MINB
#define K_UPDATE_MINB 10
#endif
template <int UPD_GROUP>
__global__ void __launch_bounds__(32 * UPD_WARPS, K_UPDATE_MINB) k_update(const DevWorld *worlds, int step_off) {
    LOAD_WORLD(W, worlds)
    UPD_BEGIN
    __shared__ HeadSmem hs_all[UPD_GROUPS_PER_BLOCK(UPD_GROUP)];
    int gw = (int)((blockIdx.x * blockDim.x + threadIdx.x) / UPD_GROUP);
    if (gw >= W.nseg) return;
    // work item = (link, part): the warp sweeps the 32-position windows part, part+nparts, ... of the link's queue, so the
    // work is proportional to the platoons present, and long queues are shared by several warps
    int l = W.seg_link[gw], part = W.seg_start[gw] & 0xffff, nparts = W.seg_start[gw] >> 16, L = W.L;
    int t = *W.t_base + step_off, cur = t & 1;
    size_t C = (size_t)W.C;
    int cap = W.l_rcap[l];
    long long o = W.l_roff[l];
    int hd0 = LQ(W, l).head[cur], pk2 = LC(W, l).pop_k2, tl = LQ(W, l).tail;
    int lanes = W.l_lanes[l];
    double len = W.l_length[l], udt = W.l_udt[l], dpl = W.l_dpl_dn[l];
    int hd = hd0 + pk2, n = tl - hd;
    int s0 = part;
    if (n == 0) {  // empty link: nothing moves, nothing is counted; only the double-buffered head is carried over
        if (s0 == 0 && (threadIdx.x & (UPD_GROUP - 1)) == 0) { LQ(W, l).head[cur ^ 1] = hd; LC(W, l).hcount = 0; }
        return;
    }
    const double *Xc = W.X + (size_t)cur * C + o;
    if (s0 == 0) {
        // ---- link head
        HeadSmem &H = hs_all[(threadIdx.x / UPD_GROUP) % UPD_GROUPS_PER_BLOCK(UPD_GROUP)];
        int hc = n < lanes ? n : lanes;
        if (hc > UX_MAXLANE) hc = UX_MAXLANE;
        int end_node = W.l_end[l];
        bool dead_end = W.n_out_off[end_node + 1] == W.n_out_off[end_node];
        UPD_FOR(lane) {
            for (int j = lane; j < hc; j += UPD_GROUP) {
                int slot = ring_slot(hd, j, cap);
                int vid = W.VID[o + slot];
                double x = Xc[slot];
                double atl_ = 0.0, ex_ = 0.0, di_ = 0.0; int dts_ = 0;
                double xn = x + udt;  // positions < lanes have no leader
                if (xn < x) xn = x;
                bool over = xn > len;
                double mr_ = xn - len;
                if (over) xn = len;
                unsigned char kind = 0;
                int nx = -1;
                if (fabs(xn - len) < 1e-9) {  // only a platoon that reaches the link end needs its per-platoon fields (most do not)
                    if (end_node == W.v_dest[vid]) kind = 1;
                    else if (dead_end) kind = 2;  // traffi.cpp:864-871 (trip_abort is always 1 in this engine)
                    else { kind = 3; nx = route_choice(W, vid, end_node, t, (unsigned)vid, UX_RNG_VEHROUTE, 0u); }
                    if (kind != 3) { atl_ = W.v_atl[vid]; ex_ = W.v_entry_x[vid]; di_ = W.v_dist[vid]; dts_ = W.v_depart_ts[vid]; }
                }
                if (over) W.v_mr[vid] = mr_;
                H.vid[j] = vid; H.kind[j] = kind; H.xn[j] = xn; H.next[j] = nx;
                H.atl[j] = atl_; H.ex[j] = ex_; H.dist[j] = di_; H.dts[j] = dts_;
            }
        }
        UPD_SYNC();
        UPD_FOR(lane) if (lane == 0) {
            // lane relabel when the end node (smaller id) emptied the link before this step's pushes
            if ((W.l_flags[l] & LF_END_LT_START) && !(W.l_flags[l] & LF_SEES_POPS) && pk2 > 0) {
                int n0 = LQ(W, l).tail_k1 - hd0, pushes = tl - LQ(W, l).tail_k1;
                if (pk2 == n0 && pushes > 0) for (int j = 0; j < pushes; j++) W.LANE[o + ring_slot(hd, j, cap)] = j % lanes;
            }
            LSTAT(W, l) += n;
            int k = 0, evc = 0, trips = 0, last_f = -1;
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
                        if (!touched) { evc = LC(W, l).ev_cnt; trips = LC(W, l).trips; cumdep = W.cum_dep[TLI(W, t, l)]; touched = true; }
                        trips += 1; cumdep += W.dn;
                        double atl = H.atl[j];
                        int es = (int)(atl / W.dt);
                        if (es < 0) es = 0;
                        if (es >= W.T) es = W.T - 1;
                        evc += 1;
                        W.ev_ord[TLI(W, es, l)] = evc;
                        W.ev_tt[TLI(W, es, l)] = ((double)t + 1.0) * W.dt - atl;
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
                if (f) last_f = j;
            }
            if (touched) { LC(W, l).ev_cnt = evc; LC(W, l).trips = trips; W.cum_dep[TLI(W, t, l)] = cumdep; }
            LQ(W, l).head[cur ^ 1] = hd + k;
            // head positions the next transfer has to look at: up to the last one that is a candidate or waits for its trip
            // end (everything behind it is mid-link), counted from the new head -- 0 for most links in most steps
            LC(W, l).hcount = last_f >= 0 ? last_f - k + 1 : 0;
            if (last_f >= 0) NS(W, end_node).node_act = t + 1;  // the end node has something to transfer next step (same value from every in-link)
            H.k_ended = k;
        }

        if (W.log_mode) {  // positions that may have a just-ended leader are logged here, once the ended count is known
            UPD_SYNC();
            int nlog = n < 2 * lanes ? n : 2 * lanes, k_ended = H.k_ended;
            UPD_FOR(lane) { for (int pos = lane; pos < nlog; pos += UPD_GROUP) log_run_record(W, l, t, cur, hd, n, pos, cap, o, lanes, len, udt, dpl, k_ended, 0); }
            UPD_SYNC();
        }
    }
    UPD_FOR(lane) {
        for (int pos = part * UPD_GROUP + lane; pos < n; pos += nparts * UPD_GROUP) {
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
// K3, ensemble form ("replica lanes"): the same link update, but one THREAD per (link, replica) -- a warp is one link in
// 32 seeded replicas.  k_update spends a whole warp on a link whose head has 1-3 positions and whose queue holds two
// platoons on average (Chicago sketch: 343 warp instructions per link and replica, 13 lanes active); here the 32 replicas
// of the link walk the same code together, the static link constants are one broadcast load per warp, and the kernel is one
// wave of L*R/32 warps instead of L*R.  Each thread runs the head positions and the first UPD_RL_SERIAL queue positions
// serially; longer queues are finished warp-cooperatively, one replica at a time (32 positions per pass, as k_update does).
// The replicas' DevWorld structs are staged in shared memory once per block (row stride = odd number of 8-byte words: the
// per-lane pointer reads are bank-conflict free), so every helper (route_choice, log_run_record, dev_error) is shared with
// k_update.  Bit-identical to k_update: the arithmetic per link is the same and links do not interact within this kernel.
// ---------------------------------------------------------------------------------------------------
#ifndef UPD_RL_WARPS
#define UPD_RL_WARPS 8
#endif
// 4 blocks per SM (64 registers, 32 warps resident): the kernel is bound by load latency, not by registers -- B200, Chicago sketch x1024:
// 649 us (95 registers, 2 blocks) -> 506 (80, 3 blocks) -> 488 (64, 4 blocks) -> 554 (5 blocks: spills)
#ifndef UPD_RL_MINB
#define UPD_RL_MINB 4
#endif
#define UPD_RL_SERIAL 32
#define UPD_RL_STRIDE (((sizeof(DevWorld) + 7) / 8) | 1)

UX_DEV void upd_newell(const DevWorld &W, int l, int t, int cur, int hd, int n, int pos, int cap, long long o, int lanes,
                       double len, double udt, double dpl, const double *Xc) {
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
    W.X[(size_t)(cur ^ 1) * (size_t)W.C + o + p] = xn;
    W.V[o + p] = (xn - x) / W.dt;
}

#define UPD_RL_HITEMS 64
struct HeadItems { int vid[UPD_RL_HITEMS], next[UPD_RL_HITEMS]; double xn[UPD_RL_HITEMS]; unsigned char kind[UPD_RL_HITEMS]; };

// The front logic of one head position (what lane 0 of k_update does per position): flags, and the trip end when everything
// ahead of it has ended in this pass (Vehicle::end_trip, traffi.cpp:886-927).
struct HeadState { int k, evc, trips, last_f; double cumdep; bool touched; };
UX_DEV void upd_head_position(const DevWorld &W, HeadState &H, int l, int t, int j, int vid, int kind, int nx, double xn) {
    unsigned char f = 0;
    if (kind == 3) { W.v_next[vid] = nx; f = VF_CAND; }
    else if (kind == 1 || kind == 2) {
        f = kind == 1 ? VF_WAIT_END : (VF_WAIT_END | VF_ABORT);
        if (kind == 2) W.v_next[vid] = -1;
        if (H.k == j) {
            if (!H.touched) { H.evc = LC(W, l).ev_cnt; H.trips = LC(W, l).trips; H.cumdep = W.cum_dep[TLI(W, t, l)]; H.touched = true; }
            H.trips += 1; H.cumdep += W.dn;
            double atl = W.v_atl[vid];
            int es = (int)(atl / W.dt);
            if (es < 0) es = 0;
            if (es >= W.T) es = W.T - 1;
            H.evc += 1;
            W.ev_ord[TLI(W, es, l)] = H.evc;
            W.ev_tt[TLI(W, es, l)] = ((double)t + 1.0) * W.dt - atl;
            W.v_atl[vid] = (double)t * W.dt;
            W.v_dist[vid] = W.v_dist[vid] + (xn - W.v_entry_x[vid]);
            W.v_link[vid] = -1;
            if (kind == 2) { W.v_state[vid] = UX_VS_ABORT; W.v_arr_ts[vid] = -1; W.v_tt[vid] = -1.0; }
            else { W.v_state[vid] = UX_VS_END; W.v_arr_ts[vid] = t; W.v_tt[vid] = (double)t * W.dt - (double)W.v_depart_ts[vid] * W.dt; }
            if (W.log_mode) { W.v_end_ts[vid] = t; W.v_end_kind[vid] = 1; }
            f = 0;
            H.k++;
        }
    }
    W.v_flags[vid] = f;
    if (f) H.last_f = j;
}

// The decision of one head position (the parallel phase of k_update's link head): where it gets to, and whether that is
// the link end -- then trip end (1), abort at a dead end (2) or a next-link draw (3).
UX_DEV int upd_head_decide(const DevWorld &W, int t, int vid, double x, double udt, double len, int end_node, bool dead_end, double *xn_out, int *nx_out) {
    double xn = x + udt;  // positions < lanes have no leader
    if (xn < x) xn = x;
    bool over = xn > len;
    double mr_ = xn - len;
    if (over) xn = len;
    int kind = 0, nx = -1;
    if (fabs(xn - len) < 1e-9) {
        if (end_node == W.v_dest[vid]) kind = 1;
        else if (dead_end) kind = 2;  // traffi.cpp:864-871 (trip_abort is always 1 in this engine)
        else { kind = 3; nx = route_choice(W, vid, end_node, t, (unsigned)vid, UX_RNG_VEHROUTE, 0u); }
    }
    if (over) W.v_mr[vid] = mr_;
    *xn_out = xn; *nx_out = nx;
    return kind;
}

__global__ void __launch_bounds__(32 * UPD_RL_WARPS, UPD_RL_MINB) k_update_rl(const DevWorld *worlds, int step_off, int R, int L_all) {
    int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
    int r = (int)blockIdx.y * 32 + lane, l = (int)blockIdx.x * UPD_RL_WARPS + warp;
    // blockIdx.z > 0: helper blocks for long queues -- they take every gridDim.z-th pass of the car-following list and leave
    // at once (before staging anything) when none of their eight links has that much queued
    const int part = (int)blockIdx.z, nparts = (int)gridDim.z;
#ifdef UX_EMU
    if (r >= R || part > 0) return;
    static DevWorld W_sh; W_sh = worlds[r]; const DevWorld &W = W_sh;
    if (l >= W.L) return;
    const bool active = true;
#else
    if (part > 0) {
        int n_q = 0;
        if (r < R && l < L_all) {  // the queue counters are replica-minor: the lanes' loads are one contiguous row
            const DevWorld &W0 = worlds[0];
            int cur_ = (*W0.t_base + step_off) & 1;
            size_t e = (size_t)l * (size_t)W0.RS + (size_t)r;
            n_q = W0.lq[e].tail - (W0.lq[e].head[cur_] + W0.lc[e].pop_k2);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) n_q += __shfl_xor_sync(0xffffffffu, n_q, d);
        if (!__syncthreads_or(n_q > part * 128)) return;
    }
    extern __shared__ unsigned long long upd_rl_smem[];
    unsigned long long (*sw)[UPD_RL_STRIDE] = reinterpret_cast<unsigned long long (*)[UPD_RL_STRIDE]>(upd_rl_smem);
    {
        const int nw = (int)(sizeof(DevWorld) / 8);
        int nrep = R - (int)blockIdx.y * 32; if (nrep > 32) nrep = 32;
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(worlds + (size_t)blockIdx.y * 32);
        // flat copy (a row-per-warp loop without the index division measured 10 % slower: fewer loads in flight)
        for (int q = (int)threadIdx.x; q < nrep * nw; q += 32 * UPD_RL_WARPS) sw[q / nw][q % nw] = src[q];
    }
    __syncthreads();
    const bool active = r < R;
    const DevWorld &W = *reinterpret_cast<const DevWorld *>(&sw[active ? lane : 0][0]);
    if (l >= W.L) return;  // warp-uniform
#endif
    int L = W.L, t = *W.t_base + step_off, cur = t & 1;
    size_t C = (size_t)W.C;
    int cap = W.l_rcap[l], lanes = W.l_lanes[l];
    long long o = W.l_roff[l];
    double len = W.l_length[l], udt = W.l_udt[l], dpl = W.l_dpl_dn[l];
    int end_node = W.l_end[l];
    bool dead_end = W.n_out_off[end_node + 1] == W.n_out_off[end_node];
    int hd = 0, n = 0, hc = 0;
    if (active) {
        int hd0 = LQ(W, l).head[cur], pk2 = LC(W, l).pop_k2, tl = LQ(W, l).tail;
        hd = hd0 + pk2; n = tl - hd;
        if (part == 0) { hc = n < lanes ? n : lanes; if (hc > UX_MAXLANE) hc = UX_MAXLANE; }
        if (part == 0 && pk2 > 0) {  // lane relabel when the end node (smaller id) emptied the link before this step's pushes
            int fl = W.l_flags[l];
            if ((fl & LF_END_LT_START) && !(fl & LF_SEES_POPS)) {
                int n0 = LQ(W, l).tail_k1 - hd0, pushes = tl - LQ(W, l).tail_k1;
                if (pk2 == n0 && pushes > 0) for (int j = 0; j < pushes; j++) W.LANE[o + ring_slot(hd, j, cap)] = j % lanes;
            }
        }
        if (part == 0 && n > 0) LSTAT(W, l) += n;
    }
    HeadState H = {0, 0, 0, -1, 0.0, false};
#ifdef UX_EMU
    const double *Xc = W.X + (size_t)cur * C + o;
    for (int j = 0; j < hc; j++) {
        int vid = W.VID[o + ring_slot(hd, j, cap)], nx; double xn;
        int kind = upd_head_decide(W, t, vid, Xc[ring_slot(hd, j, cap)], udt, len, end_node, dead_end, &xn, &nx);
        upd_head_position(W, H, l, t, j, vid, kind, nx, xn);
    }
#else
    // ---- link head.  The head positions of the link in all 32 replicas form one item list (prefix sum of min(lanes, n));
    //      decisions are taken one lane per item -- the route draws of different replicas run side by side -- then every lane
    //      walks the items of its own replica in order for the "is it the front" logic.
    __shared__ int pre_all[UPD_RL_WARPS][33], hd_all[UPD_RL_WARPS][32], hpre_all[UPD_RL_WARPS][33];
    __shared__ HeadItems hi_all[UPD_RL_WARPS];
    int *pre = pre_all[warp], *hdv = hd_all[warp], *hpre = hpre_all[warp];
    HeadItems &HI = hi_all[warp];
    int incl = n, hincl = hc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, d), hv = __shfl_up_sync(0xffffffffu, hincl, d);
        if (lane >= d) { incl += v; hincl += hv; }
    }
    int total = __shfl_sync(0xffffffffu, incl, 31), htotal = __shfl_sync(0xffffffffu, hincl, 31);
    pre[lane] = incl - n; hpre[lane] = hincl - hc; hdv[lane] = hd;
    if (lane == 31) { pre[32] = total; hpre[32] = htotal; }
    __syncwarp();
    for (int hbase = 0; hbase < htotal; hbase += UPD_RL_HITEMS) {
        int hend = hbase + UPD_RL_HITEMS < htotal ? hbase + UPD_RL_HITEMS : htotal;
        for (int i = hbase + lane; i < hend; i += 32) {
            int lo = 0;
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) if (hpre[lo + st] <= i) lo += st;
            const DevWorld &Ws = *reinterpret_cast<const DevWorld *>(&sw[lo][0]);
            int slot = ring_slot(hdv[lo], i - hpre[lo], cap);
            int vid = Ws.VID[o + slot], nx; double xn;
            int kind = upd_head_decide(Ws, t, vid, Ws.X[(size_t)cur * C + o + slot], udt, len, end_node, dead_end, &xn, &nx);
            HI.vid[i - hbase] = vid; HI.next[i - hbase] = nx; HI.xn[i - hbase] = xn; HI.kind[i - hbase] = (unsigned char)kind;
        }
        __syncwarp();
        if (active) {
            int i0 = hpre[lane] > hbase ? hpre[lane] : hbase, i1 = hpre[lane] + hc < hend ? hpre[lane] + hc : hend;
            for (int i = i0; i < i1; i++) upd_head_position(W, H, l, t, i - hpre[lane], HI.vid[i - hbase], HI.kind[i - hbase], HI.next[i - hbase], HI.xn[i - hbase]);
        }
        __syncwarp();
    }
#endif
    if (active && part == 0) {
        if (H.touched) { LC(W, l).ev_cnt = H.evc; LC(W, l).trips = H.trips; W.cum_dep[TLI(W, t, l)] = H.cumdep; }
        LQ(W, l).head[cur ^ 1] = hd + H.k;
        LC(W, l).hcount = H.last_f >= 0 ? H.last_f - H.k + 1 : 0;  // as in k_update
        if (H.last_f >= 0) NS(W, end_node).node_act = t + 1;
        if (W.log_mode) {  // positions that may have a just-ended leader, now that the ended count is known
            int nlog = n < 2 * lanes ? n : 2 * lanes;
            for (int pos = 0; pos < nlog; pos++) log_run_record(W, l, t, cur, hd, n, pos, cap, o, lanes, len, udt, dpl, H.k, 0);
        }
#ifdef UX_EMU
        for (int pos = 0; pos < n; pos++) upd_newell(W, l, t, cur, hd, n, pos, cap, o, lanes, len, udt, dpl, Xc);
#endif
    }
#ifndef UX_EMU
    // ---- car following: the queue positions of the link in all 32 replicas form one list of work items (prefix sum of the
    //      queue lengths over the lanes); every lane takes four items per pass and issues their loads before the first store,
    //      so a pass costs one memory round trip for 128 platoons whatever replica they belong to
    __syncwarp();  // the log records above read V before this step's move
    for (int base = part * 128; base < total; base += nparts * 128) {
        int rr[4], ps[4], pp[4];
        double xx[4], gg[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int i = base + q * 32 + lane;
            rr[q] = -1;
            if (i < total) {
                int lo = 0;  // largest replica lane with pre[lo] <= i (empty queues share their successor's offset and lose)
#pragma unroll
                for (int st = 16; st >= 1; st >>= 1) if (pre[lo + st] <= i) lo += st;
                int pos = i - pre[lo], hs = hdv[lo];
                const DevWorld &Ws = *reinterpret_cast<const DevWorld *>(&sw[lo][0]);
                const double *Xs = Ws.X + (size_t)cur * C + o;
                int p = ring_slot(hs, pos, cap);
                rr[q] = lo; ps[q] = pos; pp[q] = p;
                xx[q] = Xs[p];
                gg[q] = pos >= lanes ? Xs[ring_slot(hs, pos - lanes, cap)] : 0.0;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (rr[q] < 0) continue;
            const DevWorld &Ws = *reinterpret_cast<const DevWorld *>(&sw[rr[q]][0]);
            int pos = ps[q];
            if (Ws.log_mode && pos >= 2 * lanes) log_run_record(Ws, l, t, cur, hdv[rr[q]], pre[rr[q] + 1] - pre[rr[q]], pos, cap, o, lanes, len, udt, dpl, 0, 0);
            double x = xx[q], xn = x + udt;
            if (pos >= lanes) {
                double gap = gg[q] - dpl;
                if (gap < x) gap = x;
                if (xn > gap) xn = gap;
            }
            if (xn < x) xn = x;
            if (xn > len) xn = len;
            Ws.X[(size_t)(cur ^ 1) * C + o + pp[q]] = xn;
            Ws.V[o + pp[q]] = (xn - x) / Ws.dt;
        }
    }
#endif
}
