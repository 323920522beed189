// ux_host_world.inl -- part of engine.cu (one translation unit; included in the order engine.cu lists).
// Host side, C++ linkage: scenario structs, the world handle, error helper, slab allocator, upload helpers, host-thread helpers.

// =====================================================================================================
// host side
// =====================================================================================================
struct HNode {
    double x, y, sig_offset, fc; std::vector<double> sig; int lanes;
    int sig_phase0; double sig_t0, fc_rem0;
    std::vector<int> in, out;
};
struct HLink {
    int start, end, lanes; double length, vmax, kappa, delta, dpl, tau, w, capacity, mp, cap_out, cap_in, cap_out_rem0, cap_in_rem0, penalty, tt0;
    std::vector<int> sg; std::vector<double> toll;
};
struct HVeh { int orig, dest, ts; bool aborted = false; int ts0 = 0; };  // aborted: taken out before the run (UX_OPT_ABORT_VEHICLE); ts0 = its departure step as given
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
    bool initialized = false, uploaded = false, net_dirty = false, sig_tables_dirty = false;
    std::string err;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<void *> allocs;  // cudaMalloc'ed scratch (freed with the world)
    std::vector<Slab> slabs;     // state memory (returned to the slab cache)
    std::vector<DevWorld> hw;  // host copies of the per-replica device worlds
    DevWorld *dworlds = nullptr;
    const double *d_init_co = nullptr, *d_init_ci = nullptr, *d_init_st = nullptr, *d_init_fc = nullptr; const int *d_init_ph = nullptr;  // initial tokens / signal state (k_init_dyn)
    bool uniform = false;  // the replica-major arrays of replica r sit at replica 0's + r * rep_stride (what the replica-lane kernels need)
    int n_levels = 1, n_active = 0;
    long long launches = 0, C = 0; int nseg = 0, n_edges = 0;
    int route_tables_at = -1;  // UX_OPT_ROUTE_TABLES_AT
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
    int upd_group = 32;      // lanes per (link, part) item of k_update: 16 for large, lightly loaded worlds
    std::vector<int4> h_node_blks; int4 *d_node_blks = nullptr; size_t node_blks_cap = 0; int n_node_blks = 0, node_heavy_lanes = 12;  // block table of k_node_h
    bool node_rl = false;    // node step in replica-lane form (k_link_pre<true> + k_node_rl): ensembles of >= 32 replicas
    bool pre_split = false;  // k_link_pre runs the scalar link / node bookkeeping (worlds with >= 40k link+node states per step)
    int swap_last = 0;             // replica of the most recent result (scalar queries without a replica index)
    double swap_ms = 0;
    // Edie states of the last ux_edie_state call
    struct EdieResult { int replica = -1; std::vector<long long> off; std::vector<int> nx; std::vector<double> tn, dn; } edie;
    // OD / link summary of the last ux_analysis_summary call
    struct SummaryResult { int replica = -1; std::vector<int> o, d; std::vector<double> od, link; } summary;
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

// a replica's copy of an array that replica 0 already holds: device-to-device, nothing staged on the host
template <typename T>
static T *dclone(ux_world *w, const T *src, size_t n) {
    T *p = dalloc<T>(w, n, false);
    if (p && n) cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyDeviceToDevice, w->stream);
    return p;
}

template <typename T>
static T *dgrow(ux_world *w, T *old, size_t oldn, size_t newn, int fill_byte) {
    T *p = dalloc<T>(w, newn, false);
    if (!p) return nullptr;
    if (fill_byte) cudaMemsetAsync(p, fill_byte, (newn ? newn : 1) * sizeof(T), w->stream);
    if (old && oldn) cudaMemcpyAsync(p, old, oldn * sizeof(T), cudaMemcpyDeviceToDevice, w->stream);
    return p;
}

static int ensure_scratch(ux_world *w, size_t bytes) {
    if (bytes <= w->scratch_bytes) return 0;
    size_t want = std::max(bytes, std::max<size_t>(2 * w->scratch_bytes, (size_t)4 << 20));  // from the slabs: no driver allocation in steady state
    char *p = dalloc<char>(w, want, false);
    if (!p) return fail(w, UX_ERR_CUDA, "device allocation failed (%zu bytes of scratch)", want);
    w->d_scratch = (double *)p; w->scratch_bytes = want;
    return 0;
}

// One member of a replica-minor record array, for one replica, as a contiguous DEVICE array in scratch: `first` = address of the
// member in entity 0 (e.g. &LQ(d, 0).tail), consecutive entities are RS records apart.
template <typename T>
static T *gather_device(ux_world *w, const void *first, size_t rec_bytes, long long n, size_t scratch_off = 0) {
    if (ensure_scratch(w, scratch_off + sizeof(T) * (size_t)(n > 0 ? n : 1) + 16)) return nullptr;
    T *out = (T *)((char *)w->d_scratch + scratch_off);
    if (n > 0) UX_LAUNCH(k_gather<T>, dim3((unsigned)((n + 255) / 256)), dim3(256), w->stream, (const char *)first, rec_bytes * (size_t)w->hw[0].RS, n, out);
    return out;
}
template <typename T>
static bool fetch_field(ux_world *w, const void *first, size_t rec_bytes, long long n, std::vector<T> &h) {
    h.resize((size_t)n);
    T *d = gather_device<T>(w, first, rec_bytes, n);
    if (!d) return false;
    if (n > 0 && cudaMemcpyAsync(h.data(), d, sizeof(T) * (size_t)n, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) return false;
    return cudaStreamSynchronize(w->stream) == cudaSuccess;
}

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

