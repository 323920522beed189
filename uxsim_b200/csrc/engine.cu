// engine.cu -- B200-native engine for UXsim's per-timestep simulation loop (World.exec_simulation).
//
// One translation unit: sm_100a CUDA kernels + the C++ host driver + the C-ABI of include/uxsim_b200.h.
// Reference behaviour: toruseo/UXsim uxsim/trafficpp/traffi.cpp (C++ engine) == uxsim/uxsim.py (Python
// engine); each kernel cites the reference functions it implements.  Design notes: DESIGN.md.
//
// Data layout (all FP64 where the reference is FP64; FMA contraction disabled with -fmad=false so every
// compare that decides an integer outcome sees the reference's values):
//   * every link owns a ring of `cap` slots inside one big SoA arena: X[2][C] (positions, double-buffered by
//     timestep parity: X[t&1] = position at the start of step t, X[(t+1)&1] = position before the last move,
//     i.e. the reference's x_old), V[C], VID[C], LANE[C].  A platoon's leader is the slot `lanes` positions
//     ahead in the same ring (traffi.cpp:175-181), so car-following is a coalesced read of X[p] and X[p-lanes].
//   * per-link queue = (head[2], tail) monotone counters; head is double-buffered like X because the update
//     kernel both reads it (all slots) and advances it (trip ends).
//   * time series are time-major [T][L] so one step writes one contiguous row.
//   * R replicas (seeds) = R DevWorld structs sharing the static network arrays; blockIdx.y = replica.
//
// Timestep = 2 kernels (+ route update every DELTAT_ROUTE steps), replayed from CUDA graphs:
//   k_node     warp per node : Link::update of the link sides the node owns, Node::generate, signal, node capacity,
//              then Node::transfer; all dependency levels in one launch (see DESIGN.md "transfer order")
//   k_update   warp per (link, part) : car_follow_newell + Vehicle::update for the occupied positions; the part-0
//              warp also handles the link head (trip end, abort, next-link choice) and the per-step log records
//   k_edge_cost / k_sssp / k_duo(+finish) : update_adj_time_matrix, route_search_all, route_choice_duo
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <mutex>
#include <chrono>
#include <algorithm>
#include <random>
#include <queue>
#include <thread>
#include <atomic>
#include <memory>

#include "../../include/uxsim_b200.h"
#include "../../include/ux_rng.h"

#ifdef UX_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define UX_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

// One translation unit, several files: device code first, then the host driver.
#include "ux_device.cuh"
#include "ux_kernels_step.cuh"
#include "ux_kernels_route.cuh"
#include "ux_kernels_results.cuh"
#include "ux_kernels_dta.cuh"
#include "ux_host_world.inl"

extern "C" {
#include "ux_host_build.inl"
#include "ux_host_run.inl"
#include "ux_host_results.inl"
#include "ux_host_dta.inl"
#include "ux_host_arrays.inl"
}  // extern "C"
