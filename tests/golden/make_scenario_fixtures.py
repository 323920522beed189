"""Convert the reference's scenario DATA files (not source) into the flat .npz tables that
uxsim_b200.scenario.sioux_falls() / chicago_sketch() load.  Run in the dev container only (it reads
/root/reference/dat, which does not exist on the GPU box):

    python tests/golden/make_scenario_fixtures.py

Sources:
  dat/siouxfalls_{nodes,links,demand}.csv           (C1, tests/test_verification_sioux_falls.py:26-34)
  dat/chicago_sketch.uxsim_scenario (dill pickle)   (C3, example_23en_huge_scale_simlation_at_Chicago.py)
    Chicago-Sketch network by the Transportation Networks for Research Core Team,
    https://github.com/bstabler/TransportationNetworks/tree/master/Chicago-Sketch -- academic research
    use; the source must be cited in publications.
"""
import csv
import os
import sys

import numpy as np

REF = os.environ.get("UXSIM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "uxsim_b200", "data")


def sioux():
    nodes = [r for r in csv.reader(open(f"{REF}/dat/siouxfalls_nodes.csv")) if r[1] != "x"]
    links = [r for r in csv.reader(open(f"{REF}/dat/siouxfalls_links.csv")) if r[3] != "length"]
    dem = [r for r in csv.reader(open(f"{REF}/dat/siouxfalls_demand.csv")) if r[2] != "start_t"]
    np.savez_compressed(
        os.path.join(OUT, "siouxfalls.npz"),
        node_name=np.array([r[0] for r in nodes]), node_x=np.array([float(r[1]) for r in nodes]),
        node_y=np.array([float(r[2]) for r in nodes]),
        link_name=np.array([r[0] for r in links]), link_start=np.array([r[1] for r in links]),
        link_end=np.array([r[2] for r in links]), link_length=np.array([float(r[3]) for r in links]),
        link_u=np.array([float(r[4]) for r in links]), link_kappa=np.array([float(r[5]) for r in links]),
        link_merge_priority=np.array([float(r[6]) for r in links]),
        dem_orig=np.array([r[0] for r in dem]), dem_dest=np.array([r[1] for r in dem]),
        dem_t0=np.array([float(r[2]) for r in dem]), dem_t1=np.array([float(r[3]) for r in dem]),
        dem_flow=np.array([float(r[4]) for r in dem]),
        dem_volume=np.array([float(r[5]) if len(r) > 5 and r[5] != "" else -1.0 for r in dem]),
    )


def chicago():
    import dill
    d = dill.load(open(f"{REF}/dat/chicago_sketch.uxsim_scenario", "rb"))
    N, L, D = d["Nodes"], d["Links"], d["adddemand"]
    np.savez_compressed(
        os.path.join(OUT, "chicago_sketch.npz"),
        meta=np.array(str(d["meta_data"])),
        node_name=np.array([str(n["name"]) for n in N]), node_x=np.array([float(n["x"]) for n in N]),
        node_y=np.array([float(n["y"]) for n in N]),
        link_name=np.array([str(l["name"]) for l in L]), link_start=np.array([str(l["start_node"]) for l in L]),
        link_end=np.array([str(l["end_node"]) for l in L]), link_length=np.array([float(l["length"]) for l in L]),
        link_u=np.array([float(l.get("free_flow_speed", 20)) for l in L]),
        link_kappa=np.array([float(l.get("jam_density", 0.2)) for l in L]),
        link_lanes=np.array([int(l.get("number_of_lanes", 1)) for l in L]),
        link_merge_priority=np.array([float(l.get("merge_priority", 1)) for l in L]),
        link_capacity_out=np.array([float(l["capacity_out"]) if l.get("capacity_out") is not None else -1.0 for l in L]),
        link_capacity_in=np.array([float(l["capacity_in"]) if l.get("capacity_in") is not None else -1.0 for l in L]),
        dem_orig=np.array([str(x["orig"]) for x in D]), dem_dest=np.array([str(x["dest"]) for x in D]),
        dem_t0=np.array([float(x["t_start"]) for x in D]), dem_t1=np.array([float(x["t_end"]) for x in D]),
        dem_flow=np.array([float(x["flow"]) for x in D]), dem_volume=np.array([float(x["volume"]) for x in D]),
    )


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    sioux()
    chicago()
    print("wrote", os.listdir(OUT))
