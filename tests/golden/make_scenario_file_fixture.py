"""Generate tests/golden/small_grid.uxsim_scenario with the REFERENCE's own writer (dev container only; needs /root/reference):

    python tests/golden/make_scenario_file_fixture.py

A 4x4 grid with a signalised centre and every demand form the file format records (adddemand, adddemand_point2point,
adddemand_nodes2nodes2, adddemand_area2area2), written by `World.save_scenario` (scenario_reader_writer.py:99-150).  The file is the
fixture of tests/test_load_scenario.py, which reads it back through `CudaWorld.load_scenario`."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from ref_python_shim import import_reference  # noqa: E402

ux = import_reference()
W = ux.World(name="small_grid", deltan=5, tmax=2400, print_mode=0, save_mode=0, show_mode=0, random_seed=0, meta_data={"SOURCE": "synthetic 4x4 grid (tests/golden/make_scenario_file_fixture.py)"})
n = 4
for i in range(n):
    for j in range(n):
        W.addNode(f"n{i}{j}", i, j, signal=[30, 30] if (i, j) in ((1, 1), (2, 2)) else [0])
for i in range(n):
    for j in range(n):
        for di, dj in ((1, 0), (0, 1)):
            a, b = i + di, j + dj
            if a < n and b < n:
                W.addLink(f"l{i}{j}_{a}{b}", f"n{i}{j}", f"n{a}{b}", length=800, free_flow_speed=15, number_of_lanes=2 if i == j else 1, signal_group=[0] if di else [1], merge_priority=1 + di)
                W.addLink(f"l{a}{b}_{i}{j}", f"n{a}{b}", f"n{i}{j}", length=800, free_flow_speed=15, signal_group=[0] if di else [1], capacity_out=0.6 if (a, b) == (3, 3) else None)
W.adddemand("n00", "n33", 0, 1200, 0.3)
W.adddemand("n30", "n03", 100, 900, volume=150)
W.adddemand_point2point(0.1, 2.9, 2.9, 0.2, 0, 1000, 0.2)
W.adddemand_nodes2nodes2(["n00", "n01", "n10"], ["n32", "n23", "n33"], 200, 1100, volume=200)
W.adddemand_area2area2(0.5, 2.5, 0.8, 2.5, 0.5, 0.8, 0, 800, flow=0.25)
W.save_scenario(os.path.join(HERE, "small_grid.uxsim_scenario"))
print("written", os.path.join(HERE, "small_grid.uxsim_scenario"), "platoons:", len(W.VEHICLES))
