"""TEST INFRASTRUCTURE (dev container only).  Build a copy of the reference's C++ engine whose four random-draw
sites use the Philox stream of include/ux_rng.h instead of std::mt19937, WITHOUT changing anything else:

    python tests/golden/philox_patch_reference.py  ->  oracle/_ref/libuxsim_ref_philox.so  (git-ignored)

Why: the reference's engines are not stream-compatible with each other nor with a counter-based generator, so on
stochastic networks parity with the unmodified reference is statistical.  Swapping ONLY the source of the uniform
numbers (same call sites, same mapping from a uniform number to a decision: Fisher-Yates index, weighted pick,
noise factor) turns the reference into a bit-exact judge of the oracle on stochastic networks as well:
tests/golden/make_golden.py records its outputs, tests/test_oracle_golden.py replays them against the oracle
(CPU) and tests/test_gpu_golden.py against the CUDA engine.

The patched copy of traffi.cpp lives in a temporary directory and is deleted; no reference source enters the repo.
Sites (reference file:line):
  traffi.cpp:278-279  Node::transfer shuffle          -> ux_rng_below(seed, t, node, TRANSFER, draw++)
  traffi.cpp:334-337  Node::transfer merge lottery    -> pick with ux_rng_uniform(seed, t, node, TRANSFER, draw++)
  traffi.cpp:1027-1030 route_next_link_choice         -> pick with ux_rng_uniform(seed, t, entity, purpose, draw)
                       (entity/purpose/draw set by the two callers: generate traffi.cpp:123 -> (node, GENERATE, attempt),
                        update traffi.cpp:875 -> (vehicle, VEHROUTE, 0))
  traffi.cpp:1294     update_adj_time_matrix noise    -> 1 + noise * ux_rng_uniform(seed, t, link, LINKNOISE, 0)
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("UXSIM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref", "libuxsim_ref_philox.so")

PRELUDE = '''#include "traffi.h"
#include "%s/include/ux_rng.h"
static thread_local uint32_t g_entity = 0, g_purpose = 0, g_draw = 0;
template <typename T> static T *pick_u(const vector<T *> &items, const vector<double> &weights, double u) {
    int n = (int)items.size(); if (n == 0) return nullptr;
    double wsum = 0; for (auto w : weights) wsum += w;
    if (wsum <= 0.0) { int k = (int)(u * (double)n); if (k >= n) k = n - 1; return items[k]; }
    double r = u * wsum, accum = 0;
    for (int i = 0; i < n; i++) { accum += weights[i]; if (r <= accum) return items[i]; }
    return items.back();
}
''' % ROOT

SUBS = [
    ('#include "traffi.h"', PRELUDE),
    ('            std::uniform_int_distribution<int> dist(0, i);\n            int j = dist(w->node_rngs[id]);',
     '            int j = ux_rng_below(w->random_seed, (uint32_t)w->timestep, (uint32_t)id, UX_RNG_TRANSFER, xdraw++, i + 1);'),
    ('    // Shuffle for fairness (skip if hard_deterministic_mode)', '    uint32_t xdraw = 0;\n    // Shuffle for fairness (skip if hard_deterministic_mode)'),
    ('            chosen_veh = random_choice<Vehicle>(\n                merging_vehs,\n                merge_priorities,\n                w->node_rngs[id]);',
     '            chosen_veh = pick_u<Vehicle>(merging_vehs, merge_priorities, ux_rng_uniform(w->random_seed, (uint32_t)w->timestep, (uint32_t)id, UX_RNG_TRANSFER, xdraw++));'),
    ('        route_next_link = random_choice<Link>(\n            outlinks,\n            outlink_pref,\n            rng);',
     '        route_next_link = pick_u<Link>(outlinks, outlink_pref, ux_rng_uniform(w->random_seed, (uint32_t)w->timestep, g_entity, g_purpose, g_draw));'),
    ('        veh->route_next_link_choice(out_links, w->node_rngs[id]);',
     '        g_entity = (uint32_t)id; g_purpose = UX_RNG_GENERATE; g_draw = (uint32_t)i;\n        veh->route_next_link_choice(out_links, w->node_rngs[id]);'),
    ('                route_next_link_choice(link->end_node->out_links, w->link_rngs[link->id]);',
     '                g_entity = (uint32_t)id; g_purpose = UX_RNG_VEHROUTE; g_draw = 0;\n                route_next_link_choice(link->end_node->out_links, w->link_rngs[link->id]);'),
    ('            double noise_factor = random_range_float64(1.0, 1.0 + noise, link_rngs[ln->id]);',
     '            double noise_factor = 1.0 + noise * ux_rng_uniform(random_seed, (uint32_t)timestep, (uint32_t)ln->id, UX_RNG_LINKNOISE, 0u);'),
]


def build(out=OUT):
    src_dir = os.path.join(REF, "uxsim", "trafficpp")
    if not os.path.isdir(src_dir):
        raise SystemExit(f"reference not present at {REF}")
    tmp = tempfile.mkdtemp(prefix="uxref_philox_")
    try:
        for f in ("traffi.cpp", "traffi.h", "utils.h"):
            shutil.copy(os.path.join(src_dir, f), tmp)
        p = os.path.join(tmp, "traffi.cpp")
        s = open(p).read()
        for old, new in SUBS:
            if s.count(old) != 1:
                raise SystemExit(f"patch site not found exactly once (reference changed?):\n{old}")
            s = s.replace(old, new)
        open(p, "w").write(s)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O3", "-std=c++17", "-fopenmp", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-w",
                        f"-I{tmp}", "-shared", "-o", out, os.path.join(ROOT, "oracle", "ref_harness.cpp")], check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


if __name__ == "__main__":
    print("built", build())
