/*
 * uxsim_b200.h -- C-ABI of the B200-native UXsim simulation-loop backend.
 *
 * This is the drop-in boundary for ONE path of toruseo/UXsim: World.exec_simulation (the per-timestep
 * loop).  Every entry point below replaces one function of the reference's native interface, the
 * nanobind module `uxsim_cpp` (reference: uxsim/trafficpp/bindings.cpp), which the reference's Python
 * wrapper (uxsim/uxsim_cpp_wrapper.py) calls.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions
 *   - Nodes, links and vehicles (platoons) are addressed by their integer id = creation order, which is
 *     what the reference uses internally (traffi.cpp:49,479,730); names stay on the Python side.
 *   - All functions returning `int` return 0 on success and a negative UX_ERR_* code on failure;
 *     `ux_last_error()` then holds a message.  The error class mirrors the C++ exception the reference
 *     would have thrown (bindings: std::out_of_range -> IndexError, std::invalid_argument -> ValueError,
 *     std::runtime_error -> RuntimeError).
 *   - The world owns all host and device memory.  Device pointers handed out by
 *     `ux_get_device_array` are borrowed and valid until the next `ux_main_loop` / `ux_destroy_world`.
 *   - Single caller thread per world (the reference holds the GIL across main_loop,
 *     uxsim_cpp_wrapper.py:1192).
 *   - A world may carry R independent seeded replicas of one scenario (same network and demand, seeds
 *     random_seed + r): the natural data-parallel shard of this path.
 *
 * Three libraries export this same ABI with different symbol prefixes:
 *   ux_*   libuxsim_b200.so           the product: hand-written sm_100a CUDA (uxsim_b200/csrc/)
 *   uxo_*  oracle/liboracle.so        TEST INFRASTRUCTURE: serial C restatement of the reference
 *   uxr_*  oracle/_ref/libuxsim_ref.so TEST INFRASTRUCTURE: the reference's own traffi.cpp behind a harness
 * Only the first one is shipped; the other two exist for tests/ and bench.py's CPU baseline.
 */
#ifndef UXSIM_B200_H
#define UXSIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef UX_PREFIX
#define UX_PREFIX ux_
#endif
#define UX_CAT2(a, b) a##b
#define UX_CAT(a, b) UX_CAT2(a, b)
#define UX_FN(name) UX_CAT(UX_PREFIX, name)

#if defined(__GNUC__)
#define UX_API __attribute__((visibility("default")))
#else
#define UX_API
#endif

typedef struct ux_world ux_world; /* opaque */

/* error codes (negative return values) */
enum {
    UX_OK = 0,
    UX_ERR_RUNTIME = -1,      /* std::runtime_error      -> RuntimeError  */
    UX_ERR_INVALID = -2,      /* std::invalid_argument   -> ValueError    */
    UX_ERR_OUT_OF_RANGE = -3, /* std::out_of_range       -> IndexError    */
    UX_ERR_CUDA = -4,         /* CUDA runtime failure (no reference analogue) */
    UX_ERR_CAPACITY = -5      /* a device ring / log buffer overflowed (no reference analogue) */
};

/* vehicle states: traffi.h:32-38 (enum VehicleState) */
enum { UX_VS_HOME = 0, UX_VS_WAIT = 1, UX_VS_RUN = 2, UX_VS_END = 3, UX_VS_ABORT = 4 };

/* World construction parameters == arguments of create_world (bindings.cpp:150-181, 277-290) plus the
 * read-write World fields the wrapper sets right after (uxsim_cpp_wrapper.py:894-898). */
typedef struct ux_world_params {
    double t_max;                    /* simulation duration [s]                     (t_max)            */
    double delta_n;                  /* platoon size [veh]                          (delta_n)          */
    double tau;                      /* reaction time [s]; delta_t = tau * delta_n  (tau)              */
    double duo_update_time;          /* route-choice update interval [s]                                */
    double duo_update_weight;        /* DUO smoothing weight                                            */
    double route_choice_uncertainty; /* duo_noise: edge cost *= U(1, 1+noise)                           */
    int64_t random_seed;             /* replica r uses random_seed + r                                  */
    int32_t print_mode;              /* kept for signature parity; the engine never prints              */
    int32_t vehicle_log_mode;        /* 1: keep per-platoon per-step logs (vehicle_logging_timestep_interval=1) */
    int32_t hard_deterministic_mode; /* no random draws: argmax choices, no shuffle, no noise           */
    int32_t route_choice_update_gradual;
    int32_t no_cyclic_routing;
    int32_t adjust_node_capacity;
    int32_t instantaneous_TT_timestep_interval; /* already clamped to the route-update interval by the caller */
    int32_t num_replicas;            /* >= 1 (ignored by the oracle / reference harness: they run 1)    */
    int32_t device;                  /* CUDA device ordinal; -1 = current device                        */
    int32_t num_threads;             /* OpenMP threads for the reference harness only (World.num_threads) */
    int32_t reserved0;
} ux_world_params;

/* ---- scenario definition ------------------------------------------------------------------------ */

/* create_world (bindings.cpp:150-181).  Returns NULL on failure (message via ux_last_error(NULL)). */
UX_API ux_world *UX_FN(create_world)(const ux_world_params *params);
UX_API void UX_FN(destroy_world)(ux_world *w);

/* add_node (bindings.cpp:186-190; Node::Node traffi.cpp:47-100).  `flow_capacity` < 0 = unlimited,
 * `number_of_lanes` <= 0 = derive.  Returns the node id (>= 0) or an error code. */
UX_API int UX_FN(add_node)(ux_world *w, double x, double y, const double *signal_intervals, int n_signal,
                           double signal_offset, double flow_capacity, int number_of_lanes);

/* add_link (bindings.cpp:192-207; Link::Link traffi.cpp:465-537).  capacity_out/in < 0 = default.
 * Returns the link id (>= 0) or an error code. */
UX_API int UX_FN(add_link)(ux_world *w, int start_node, int end_node, double vmax, double kappa, double length,
                           int number_of_lanes, double merge_priority, double capacity_out, double capacity_in,
                           const int *signal_group, int n_signal_group);

/* add_demand (bindings.cpp:209-216; traffi.cpp:1910-1941): FP64 accumulation `demand += flow*delta_t;
 * while demand >= delta_n: new platoon`.  Returns the number of platoons created (>= 0) or an error. */
UX_API int64_t UX_FN(add_demand)(ux_world *w, int orig, int dest, double start_t, double end_t, double flow,
                                 const int *links_preferred, int n_links_preferred);

/* Bulk form of CppWorld.addVehicle (uxsim_cpp_wrapper.py:927-993, which calls add_demand once per
 * platoon): n platoons with departure timestep depart_ts[i] (time = depart_ts * delta_t). */
UX_API int64_t UX_FN(add_vehicles)(ux_world *w, int64_t n, const int *orig, const int *dest, const int *depart_ts);

/* Vectorised scenario ingest: the same three calls over arrays, so that a whole scenario crosses
 * the FFI in three calls instead of one call per node / link / OD row (uxsim_cpp_wrapper.py pays one nanobind
 * call each).  Semantics are exactly add_node / add_link / add_demand applied in index order; CSR offsets have
 * n+1 entries.  Return the number of entities (platoons for add_demands) created, or an error code. */
UX_API int64_t UX_FN(add_nodes)(ux_world *w, int64_t n, const double *x, const double *y, const int *signal_off,
                                const double *signal_intervals, const double *signal_offset, const double *flow_capacity,
                                const int *number_of_lanes);
UX_API int64_t UX_FN(add_links)(ux_world *w, int64_t n, const int *start_node, const int *end_node, const double *vmax,
                                const double *kappa, const double *length, const int *number_of_lanes,
                                const double *merge_priority, const double *capacity_out, const double *capacity_in,
                                const int *signal_group_off, const int *signal_group);
UX_API int64_t UX_FN(add_demands)(ux_world *w, int64_t n, const int *orig, const int *dest, const double *start_t,
                                  const double *end_t, const double *flow);

/* Vehicle.enforce_route / links_preferred / links_avoid (bindings.cpp Vehicle class; traffi.cpp:1035-1039).
 * kind: 0 = specified_route (also sets links_preferred, as enforce_route does), 1 = links_preferred,
 * 2 = links_avoid.  n == 0 clears. */
UX_API int UX_FN(set_vehicle_links)(ux_world *w, int64_t vehicle, int kind, const int *links, int n);

/* dta_batch_enforce_routes (batch_enforce_routes_csr, bindings.cpp; dta_solver.cpp:109-124): CSR of link
 * ids per vehicle; empty rows are skipped. */
UX_API int UX_FN(batch_enforce_routes_csr)(ux_world *w, const int *link_ids, const int64_t *offsets, int64_t n_vehicles);

/* World::set_t_max (traffi.cpp:1234-1243). */
UX_API int UX_FN(set_t_max)(ux_world *w, double t_max);

/* Link.set_toll_timeseries (uxsim_cpp_wrapper.py:1109-1112): n = number of timesteps. */
UX_API int UX_FN(set_toll_timeseries)(ux_world *w, int link, const double *toll, int n);

/* World::initialize_adj_matrix (traffi.cpp:1245-1265): freezes the network (adjust_node_capacity, route
 * preference allocation).  In the CUDA build this is where device state is allocated and uploaded. */
UX_API int UX_FN(initialize_adj_matrix)(ux_world *w);

/* ---- the hot path --------------------------------------------------------------------------------- */

/* World::main_loop(duration_t, until_t) (traffi.cpp:1642-1866): runs timesteps [timestep, end_ts), with
 * end_ts = floor((duration_t + time)/dt) + 1 | floor(until_t/dt) + 1 | total_timesteps (both < 0), clamped
 * to total_timesteps.  Both >= 0 is UX_ERR_RUNTIME. */
UX_API int UX_FN(main_loop)(ux_world *w, double duration_t, double until_t);

/* World::check_simulation_ongoing (traffi.cpp:1869-1875). */
UX_API int UX_FN(check_simulation_ongoing)(ux_world *w);

/* ---- mid-run interventions (setters of the nanobind Node/Link classes) ------------------------------ */
enum {
    UX_LINK_VMAX = 1,           /* Link::change_free_flow_speed (traffi.cpp:570-577) */
    UX_LINK_KAPPA = 2,          /* Link::change_jam_density     (traffi.cpp:579-587) */
    UX_LINK_CAPACITY_OUT = 3,
    UX_LINK_CAPACITY_IN = 4,
    UX_LINK_CAPACITY_OUT_REMAIN = 5,
    UX_LINK_CAPACITY_IN_REMAIN = 6,
    UX_LINK_MERGE_PRIORITY = 7,
    UX_LINK_ROUTE_CHOICE_PENALTY = 8
};
UX_API int UX_FN(set_link_param)(ux_world *w, int link, int field, double value);
UX_API int UX_FN(set_link_signal_group)(ux_world *w, int link, const int *signal_group, int n);
/* Node.signal_intervals / signal_phase / signal_t setters (uxsim_cpp_wrapper.py:89-124).
 * phase < 0 / signal_t < 0 = keep the current value. */
UX_API int UX_FN(set_node_signal)(ux_world *w, int node, const double *signal_intervals, int n_signal, int phase,
                                  double signal_t);

/* ---- results -------------------------------------------------------------------------------------- */

/* scalar queries */
enum {
    UX_I_NUM_NODES = 1,
    UX_I_NUM_LINKS = 2,
    UX_I_NUM_VEHICLES = 3,
    UX_I_TIMESTEP = 4,         /* World.timestep */
    UX_I_TOTAL_TIMESTEPS = 5,  /* World.total_timesteps */
    UX_I_NUM_REPLICAS = 6,
    UX_I_ROUTE_UPDATE_STEPS = 7, /* timestep_for_route_update */
    UX_I_PLATOON_UPDATES = 8,  /* sum over steps of RUN platoons == sum(link_stat_count), traffi.cpp:1063-1069 (replica 0) */
    UX_I_TRIPS_COMPLETED = 9,  /* trips_completed_count()  (platoons, replica 0) */
    UX_I_LOG_ENTRIES = 10,     /* total RUN log records (replica 0) */
    UX_I_KERNEL_LAUNCHES = 11, /* CUDA kernels launched so far (product only; 0 elsewhere) */
    UX_I_NUM_ACTIVE_DESTS = 12,
    UX_I_TRANSFER_LEVELS = 13, /* number of dependency levels of the node-transfer stage */
    UX_I_LINK_EVENTS = 14,     /* link-exit events recorded (replica 0) */
    UX_I_H2D_BYTES = 15,       /* bytes copied host->device so far (product only) */
    UX_I_D2H_BYTES = 16,       /* bytes copied device->host so far (product only) */
    UX_I_RING_SLOTS = 17,      /* total ring-buffer slots per replica */
    UX_I_MAX_DEPART_TS = 18,   /* largest departure timestep over all platoons (for the automatic TMAX, uxsim.py:2325-2333) */
    UX_I_KERNEL_LAUNCHES_OF = 100 /* + kernel id (UX_K_*): launches of that kernel in timing mode */
};
/* kernel ids for the timing mode */
enum { UX_K_PRE = 0, UX_K_TRANSFER = 1, UX_K_UPDATE = 2, UX_K_EDGE_COST = 3, UX_K_SSSP = 4, UX_K_DUO = 5, UX_K_MISC = 6, UX_K_COUNT = 7 };
UX_API int64_t UX_FN(get_int)(ux_world *w, int what);
UX_API double UX_FN(get_double)(ux_world *w, int what);
enum {
    UX_D_DELTA_T = 1, UX_D_TIME = 2, UX_D_T_MAX = 3, UX_D_AVE_V_SUM = 4, UX_D_STAT_SAMPLE_COUNT = 5,
    UX_D_LAST_LOOP_MS = 6,      /* device time of the last main_loop, CUDA events on the launching stream (product only) */
    UX_D_UPLOAD_MS = 7,         /* host wall time of building + uploading the device state [ms] (product only) */
    UX_D_GRAPH_BUILD_MS = 8,    /* host wall time spent capturing/instantiating CUDA graphs in the last main_loop [ms] */
    UX_D_LAST_LOOP_WALL_MS = 9, /* host wall time of the last main_loop call [ms] */
    UX_D_DTA_N_SWAP = 10,       /* RouteSwapResult::n_swap of the last dta_route_swap */
    UX_D_DTA_POTENTIAL_N_SWAP = 11,
    UX_D_DTA_TOTAL_T_GAP = 12,
    UX_D_DTA_SWAP_MS = 13,      /* device time of the last dta_route_swap's kernels (product only) */
    UX_D_DTA_N_SWAP_OF = 1000,  /* + replica: n_swap of that replica after a batched dta_route_swap (replica == -1); +1000: potential_n_swap; +2000: total_t_gap */
    UX_D_KERNEL_MS_OF = 100     /* + kernel id: summed device time of that kernel in timing mode [ms] */
};
/* engine options (product only; the test libraries accept and ignore them) */
enum {
    UX_OPT_KERNEL_TIMING = 1,   /* 1: launch kernels eagerly with a CUDA-event pair around each (for roofline numbers) */
    UX_OPT_RECORD_TRAVERSALS = 3, /* 1 (before the first main_loop): record every link entry (platoon, link, step) on the device so that
                                   dta_route_swap can read the traversed routes without the per-step vehicle logs (CUDA build; the CPU
                                   engines read their logs, as dta_get_traveled_route does, dta_solver.cpp:30-57) */
    UX_OPT_ABORT_VEHICLE = 5,   /* value = platoon id (before the first main_loop): the platoon never departs and reports state ABORT -- what setting
                                   Vehicle.state = "abort" before the run does in the reference (uxsim_cpp_wrapper.py:517-519, traffi.cpp:833-881) */
    UX_OPT_ROUTE_TABLES_AT = 4, /* timestep >= 0: the next UX_A_ROUTE_DIST / UX_A_ROUTE_NEXT reads return the tables of the route update that was
                                   current at that timestep (World::route_dist_record, traffi.cpp:1840-1851; the engine keeps the edge costs of every
                                   update -- 8 B per node pair -- and searches on demand); -1: the latest update (default) */
    UX_OPT_ALL_DESTINATIONS = 2 /* 1 (before the first main_loop): keep route preferences for every node, as the reference does,
                                   instead of only for the destinations present in the demand.  Needed only if demand towards
                                   NEW destinations will be added between main_loop calls. */
};
UX_API int UX_FN(set_option)(ux_world *w, int option, double value);

/* array results; layouts are row-major.  T = total_timesteps, L = links, N = nodes, P = vehicles. */
enum {
    /* Link.get_*_np (bindings.cpp:819-835): one row per link */
    UX_A_CUM_ARRIVAL = 1,       /* f64 [L][T]  arrival_curve                         */
    UX_A_CUM_DEPARTURE = 2,     /* f64 [L][T]  departure_curve                       */
    UX_A_TRAVELTIME_INSTANT = 3,/* f64 [L][T]                                        */
    UX_A_TRAVELTIME_ACTUAL = 4, /* f64 [L][T]  traveltime_real after ensure_traveltime_real (traffi.cpp:631-648) */
    /* per vehicle (bindings.cpp Vehicle class; get_all_vehicle_states) */
    UX_A_VEH_STATE = 10,        /* i32 [P]  effective state (abort folded in, traffi.cpp:1947-1958) */
    UX_A_VEH_ARRIVAL_TS = 11,   /* i32 [P]  arrival timestep, -1 if not arrived / aborted */
    UX_A_VEH_DEPART_TS = 12,    /* i32 [P]                                                */
    UX_A_VEH_TRAVEL_TIME = 13,  /* f64 [P]  arrival_time - departure_time, -1 if none     */
    UX_A_VEH_DISTANCE = 14,     /* f64 [P]  distance_traveled                             */
    UX_A_VEH_ORIG = 15,         /* i32 [P] */
    UX_A_VEH_DEST = 16,         /* i32 [P] */
    UX_A_VEH_LINK = 17,         /* i32 [P]  current link id or -1                         */
    UX_A_VEH_X = 18,            /* f64 [P]  current position on link                      */
    UX_A_VEH_V = 19,            /* f64 [P]  current speed                                 */
    UX_A_VEH_LANE = 20,         /* i32 [P] */
    UX_A_VEH_LINK_ARRIVAL_TIME = 21, /* f64 [P]  Vehicle.arrival_time_link: time the platoon entered its current link (bindings.cpp Vehicle fields) */
    /* per node */
    UX_A_SIGNAL_LOG = 30,       /* i32 [N][T]  Node.signal_log                            */
    UX_A_NODE_SIGNAL_PHASE = 31,/* i32 [N] */
    UX_A_NODE_SIGNAL_T = 32,    /* f64 [N] */
    UX_A_NODE_FLOW_CAPACITY = 33, /* f64 [N]  Node.flow_capacity (-1: none), after adjust_node_capacity (bindings.cpp: Node rw fields) */
    UX_A_NODE_NUM_LANES = 34,   /* i32 [N]  Node.number_of_lanes */
    /* per link, current state */
    UX_A_LINK_NUM_VEHICLES = 40,/* i32 [L]  len(link.vehicles)                            */
    UX_A_LINK_CAPACITY_IN_REMAIN = 41,  /* f64 [L] */
    UX_A_LINK_CAPACITY_OUT_REMAIN = 42, /* f64 [L] */
    UX_A_LINK_STAT_COUNT = 43,  /* f64 [L]  link_stat_count (RUN samples per link)        */
    UX_A_LINK_TRIPS_COMPLETED = 44, /* f64 [L] link_trips_completed                       */
    UX_A_LINK_AVE_V_SUM = 45,   /* f64 [L]  link_ave_v_sum                                */
    /* route choice (World.route_preference / route_dist / route_next) */
    UX_A_ROUTE_PREF = 50,       /* f64 [N][L]  rows of inactive destinations are 0        */
    UX_A_ROUTE_DIST = 51,       /* f64 [N][N]  dist[i][k] from i to k (inf if unreachable / k inactive) */
    UX_A_ROUTE_NEXT = 52,       /* i32 [N][N]  next node from i towards k, -1 if none     */
    /* World::build_all_vehicle_logs_flat_compact (traffi.cpp:1990-2081); E = total entries */
    UX_A_LOG_OFFSETS = 60,      /* i64 [P+1] */
    UX_A_LOG_T = 61,            /* f64 [E] */
    UX_A_LOG_X = 62,            /* f64 [E] */
    UX_A_LOG_V = 63,            /* f64 [E] */
    UX_A_LOG_STATE = 64,        /* i32 [E] */
    UX_A_LOG_S = 65,            /* f64 [E] */
    UX_A_LOG_LANE = 66,         /* i32 [E] */
    UX_A_LOG_LINK = 67,         /* i32 [E] */
    UX_A_LOG_N_MISSING = 68,    /* i32 [P] */
    UX_A_LTL_OFFSETS = 69,      /* i64 [P+1]  log_t_link */
    UX_A_LTL_T = 70,            /* f64 [E2] */
    UX_A_LTL_ID = 71,           /* i32 [E2]  link id, -1 home, -2 end */
    /* World::build_enter_log_data (traffi.cpp:1965-1983) */
    UX_A_ENTER_LINK = 80,       /* i32 [E3] */
    UX_A_ENTER_TIME = 81,       /* f64 [E3] */
    UX_A_ENTER_VEH = 82,        /* i32 [E3] */
    /* DTA solver batch operations.  OD route-set store (ODRouteSetStore::to_csr, bindings.cpp), pairs
     * sorted by (orig, dest): */
    UX_A_DTA_OD_ORIG = 100,     /* i32 [n_od] */
    UX_A_DTA_OD_DEST = 101,     /* i32 [n_od] */
    UX_A_DTA_OD_OFFSETS = 102,  /* i64 [n_od+1]  routes of pair p = [off[p], off[p+1]) */
    UX_A_DTA_ROUTE_OFFSETS = 103, /* i64 [n_routes+1] */
    UX_A_DTA_ROUTE_LINKS = 104, /* i32 */
    /* result of the last dta_route_swap (RouteSwapResult dta_solver.h:104-114, as the flat CSR of bindings.cpp:661-706) */
    UX_A_DTA_RS_OFFSETS = 110,  /* i64 [P+1]  routes_specified: route to enforce in the next iteration (empty = none) */
    UX_A_DTA_RS_LINKS = 111,    /* i32 */
    UX_A_DTA_RA_OFFSETS = 112,  /* i64 [P+1]  route_actual: links traversed in this run */
    UX_A_DTA_RA_LINKS = 113,    /* i32 */
    UX_A_DTA_COST_ACTUAL = 114, /* f64 [P]    arrival_time - first link entry time */
    UX_A_DTA_T_GAP = 115,       /* f64 [P]    per-platoon cost gap; total_t_gap is its sum in index order */
    UX_A_DTA_CHOICE = 116,      /* i32 [P]    -2 nothing enforced (trip unfinished), -1 keeps the traversed route, r >= 0 = index into the store's routes */
    UX_A_DTA_RA_ENTRY_T = 117,  /* f64        entry time of each link of route_actual (TraveledRouteInfo::entry_times) */
    /* result of the last edie_state: link l owns cells [off[l], off[l+1]) = a row-major [nt_l][nx_l] grid */
    UX_A_EDIE_OFFSETS = 120,    /* i64 [L+1] */
    UX_A_EDIE_NX = 121,         /* i32 [L]   cells along the link; nt_l = (off[l+1] - off[l]) / nx_l */
    UX_A_EDIE_TN = 122,         /* f64 [cells]  vehicle-seconds spent in the cell  (Link.tn_mat) */
    UX_A_EDIE_DN = 123,         /* f64 [cells]  vehicle-metres travelled in the cell (Link.dn_mat) */
    /* result of the last analysis_summary: Analyzer.od_analysis / link_analysis_coarse (analyzer.py:121-203) */
    UX_A_OD_ORIG = 130,         /* i32 [K]    OD pairs that occur in the demand, ascending in (orig * N + dest) */
    UX_A_OD_DEST = 131,         /* i32 [K] */
    UX_A_OD_STATS = 132,        /* f64 [K][6] platoons, completed platoons, mean and std of their travel time (-1 if none completed),
                                              mean and std of the distance travelled by all platoons of the pair */
    UX_A_LINKC_STATS = 133      /* f64 [L][4] traffic volume, vehicles remaining, mean and std of traveltime_actual over the steps where it
                                              is > 0 (-1, -1 for a link without volume) */
};
enum { UX_DT_F64 = 1, UX_DT_I32 = 2, UX_DT_I64 = 3 };

/* Size query: number of ELEMENTS and dtype of an array result for one replica. */
UX_API int64_t UX_FN(array_size)(ux_world *w, int what, int *dtype_out);
/* Copy an array result of `replica` into host memory (the reference's "owned numpy copy" getters,
 * bindings.cpp:26-42).  `capacity` in elements.  Returns elements written or an error code. */
UX_API int64_t UX_FN(get_array)(ux_world *w, int what, int replica, void *dst, int64_t capacity);
/* CUDA build: replica == -1 copies the per-replica state arrays UX_A_VEH_ARRIVAL_TS / VEH_TRAVEL_TIME / VEH_LINK /
 * LINK_STAT_COUNT / NODE_SIGNAL_* / LINK_CAPACITY_*_REMAIN of ALL replicas, replica-major [R][n], in one transfer. */
/* Borrowed DEVICE pointer to an array result (CUDA build only; used to wrap results as torch tensors
 * without a host copy).  Returns elements or an error code. */
UX_API int64_t UX_FN(get_device_array)(ux_world *w, int what, int replica, void **dptr_out, int *dtype_out);

/* Ensemble statistics for the multi-GPU shard: per-link sums over this world's replicas of
 * {traffic_volume, sum of actual travel time over steps, sum of squares, platoons still on link},
 * written to DEVICE memory as f64 [L][4] ready for one ncclAllReduce(sum). */
UX_API int64_t UX_FN(ensemble_link_stats_device)(ux_world *w, void **dptr_out);
UX_API int64_t UX_FN(ensemble_link_stats)(ux_world *w, double *dst, int64_t capacity);

/* ---- DTA solver batch operations (dta_solver.h, bindings.cpp:461-706) ------------------------ */

/* World.enumerate_od_route_sets_ids_cpp(k, seed) (bindings.cpp:461-465; World::enumerate_k_random_routes_cpp
 * traffi.cpp:1387-1481): up to k distinct shortest routes per connected node pair under free-flow, then randomly
 * re-weighted (mt19937(seed), U(0,100)) link costs, until the mean number of routes per pair is stable for 11 rounds.
 * The store is kept by the world (UX_A_DTA_OD_* / ROUTE_*).  Returns the number of routes or an error code. */
UX_API int64_t UX_FN(dta_enumerate_route_sets)(ux_world *w, int k, unsigned int seed);
/* World.make_od_route_set_store (bindings.cpp:467-489): an external store in CSR form. */
UX_API int UX_FN(dta_set_route_sets)(ux_world *w, int64_t n_od, const int *orig, const int *dest, const int64_t *od_offsets,
                                     const int64_t *route_offsets, const int *route_links);
/* The reference keeps ONE ODRouteSetStore object for a whole solve and hands it to every day's world
 * (DTAsolvers_cpp_adapter.py:280-291, 330): share `from`'s store with `w` without copying it (in the CUDA build the device
 * image is uploaded once per solve). */
UX_API int UX_FN(dta_share_route_sets)(ux_world *w, ux_world *from);
/* Enforced routes of ONE replica, CSR as in batch_enforce_routes_csr (the CPU engines have a single replica: replica must
 * be 0 and the call is batch_enforce_routes_csr).  A replica-batched day-to-day solve runs one independent chain of
 * route sets per replica.  Before the first main_loop. */
UX_API int UX_FN(dta_enforce_routes)(ux_world *w, int replica, const int *link_ids, const int64_t *offsets, int64_t n_vehicles);
/* World.route_swap_due / route_swap_dso (bindings.cpp:661-706; dta_route_swap_due / _dso dta_solver.cpp:130-373) on the
 * finished run of `replica`.  mode 0 = DUE (travel time + tolls), 1 = DSO (travel time + congestion externality,
 * Link::estimate_congestion_externality traffi.cpp:650-689); swap_num < 0 = probabilistic candidates (swap_prob),
 * otherwise exactly swap_num candidates (DSO only).  Results: UX_A_DTA_RS_* .. UX_A_DTA_CHOICE, UX_D_DTA_*.
 * replica == -1 (CUDA build): all replicas at once, replica r drawing its candidates from rng_seed + r; array_size of the
 * variable-length results is then an upper bound over the replicas and get_array returns the replica's actual length. */
UX_API int UX_FN(dta_route_swap)(ux_world *w, int replica, int mode, double swap_prob, int swap_num,
                                 int has_external_route_sets, unsigned int rng_seed);

/* Rewind a finished world to t = 0 for the next day of a day-to-day solve (CUDA build): dynamic state re-initialised in
 * place, each replica's routes_specified of the last dta_route_swap enforced straight from device memory, network / platoon
 * tables / captured graphs kept.  Same results as: new world + batch_enforce_routes_csr(rs) + main_loop.  The reference
 * builds a new World per day (DTAsolvers_cpp_adapter.py:318-330), so the CPU engines answer UX_ERR_RUNTIME. */
UX_API int UX_FN(dta_next_day)(ux_world *w);

/* Analyzer.compute_edie_state (analyzer.py:205-330, with compute_accurate_traj 205-247): Edie's generalised traffic state
 * per link on an (edie_dt[l] x edie_dx[l]) time-space grid, from the per-step platoon trajectories (needs vehicle_log_mode):
 * every trajectory segment adds its travel time and distance to the cells it crosses (time cell of the segment's start), the
 * ends of a link traversal are extrapolated at free-flow speed.  k = tn / (dt dx), q = dn / (dt dx), v = q / k.
 * Contributions are accumulated in 2^-24 fixed point, so the sums do not depend on the order of accumulation (and are
 * identical in the three engines); they differ from the reference's float sums by < 1e-7 relative. */
UX_API int UX_FN(edie_state)(ux_world *w, int replica, const double *edie_dt, const double *edie_dx);

/* Analyzer.od_analysis (analyzer.py:121-181) and Analyzer.link_analysis_coarse (183-203) of the finished run of `replica` as
 * reductions over the per-platoon and per-link results (the reference loops over every vehicle and link in Python): per OD pair
 * the trip counts and the mean / population standard deviation of travel time (completed trips) and of distance travelled (all
 * platoons of the pair); per link the volume, the remaining vehicles and the mean / std of the positive entries of
 * traveltime_actual.  Sums run in platoon-id / time order, so the three engines agree bit for bit; free-flow times and minimum
 * distances (static shortest paths) stay with the caller.  Results: UX_A_OD_* and UX_A_LINKC_STATS. */
UX_API int UX_FN(analysis_summary)(ux_world *w, int replica);

/* Message of the last failure on `w` (or of the last failed create_world when w == NULL). */
UX_API const char *UX_FN(last_error)(ux_world *w);

#ifdef __cplusplus
}
#endif
#endif /* UXSIM_B200_H */
