"""The golden-vector cases: (name, scenario builder, kind).
kind "python"  : produced by the reference's pure-Python engine (uxsim/uxsim.py) -- deterministic scenarios only;
kind "philox"  : produced by the reference's C++ engine with its random draws swapped for the Philox stream
                 (tests/golden/philox_patch_reference.py) -- stochastic networks, bit-exact target."""
from uxsim_b200 import scenario as S

PYTHON_CASES = [
    ("1link", S.straight_1link),
    ("2link_capout", S.straight_2link_bottleneck_capacity_out),
    ("2link_speed", S.straight_2link_bottleneck_speed),
    ("3link_spillback", S.straight_3link_spillback),
    ("2link_signal", S.straight_2link_signal),
    ("node_capacity", S.straight_node_capacity),
    ("lane_drop", S.lane_drop_2to1),
    ("merge_equal", lambda: S.merge_y(priority1=1, priority2=1)),
    ("merge_unequal", lambda: S.merge_y(priority1=1, priority2=2)),
    ("diverge", S.diverge_2route),
]
PHILOX_CASES = [
    ("merge_stochastic", lambda: S.merge_y(hard_deterministic_mode=False, random_seed=3)),
    ("3lane", lambda: S.multilane_3lane(random_seed=1)),
    ("short_chain", lambda: S.short_links_chain(random_seed=2)),
    ("siouxfalls_s1", lambda: S.sioux_falls(seed=1)),
    ("grid5_s2", lambda: S.grid(n=5, seed=2, tmax=3000, demand_t=1000)),
    ("grid6_signals_s5", lambda: S.grid(n=6, seed=5, tmax=3000, demand_t=1200, signals=True)),
    ("grid11_s2", lambda: S.grid(seed=2)),
    ("dta_grid9_s1", lambda: S.dta_grid9(seed=1)),
    ("chicago_dn30_s0", lambda: S.chicago_sketch(seed=0)),
]
