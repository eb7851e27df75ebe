#!/usr/bin/env python
"""Microseconds per timestep of the reference's shipped inputs under each small-grid strategy (run on a GPU box):
cluster-resident (default), L2-resident persistent cooperative kernel (LBM_B200_CLUSTER=0), one launch per step
(LBM_B200_PERSISTENT=0), and forced cluster sizes.  Prints one line per (case, variant)."""
import dataclasses
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {"128x128": 20000, "128x256": 20000, "256x256": 20000, "1024x1024": 5000}
VARIANTS = [("cluster", {})]
VARIANTS += [(f"cluster_{cells}x{cpt}", {"LBM_B200_CLUSTER_CELLS": str(cells), "LBM_B200_CLUSTER_CPT": str(cpt)})
             for cells, cpt in ((1024, 2), (512, 2), (512, 1), (256, 1))]
VARIANTS += [("cluster_cs8", {"LBM_B200_CLUSTER_SIZE": "8"}),
             ("persistent", {"LBM_B200_CLUSTER": "0"}), ("perstep", {"LBM_B200_PERSISTENT": "0", "LBM_B200_GRAPH": "0"}),
             ("perstep_graph", {"LBM_B200_PERSISTENT": "0", "LBM_B200_GRAPH": "1"})]

def child(case, steps):
    lbm = importlib.import_module("lbm-graphcore_b200")
    gold = os.path.join(ROOT, "tests", "golden", "params")
    p = lbm.Params.from_file(os.path.join(gold, f"input_{case}.params"))
    p = dataclasses.replace(p, max_iters=500 + 3 * steps)
    obst = lbm.load_obstacles(p.nx, p.ny, os.path.join(gold, f"obstacles_{case}.dat"))
    with lbm.Solver(p) as s:
        s.upload(lbm.initial_cells(p), obst)
        s.run(500)  # warm-up
        best = 1e30
        for _ in range(3):
            s.run(steps)
            best = min(best, s.elapsed_ms)
        print(f"{s.kernel_variant} {best * 1e3 / steps:.3f} us/step {p.nx * p.ny * steps / (best * 1e-3) / 1e6:.1f} MLUPS", flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    only = sys.argv[1:] or list(CASES)
    for case in only:
        for name, env in VARIANTS:
            if "cluster" in name and case == "1024x1024":
                continue
            e = dict(os.environ)
            e.update(env)
            r = subprocess.run([sys.executable, __file__, "--child", case, str(CASES[case])], capture_output=True, text=True, env=e, timeout=600)
            print(f"{case:10s} {name:14s} {r.stdout.strip() or ('FAILED: ' + r.stderr.strip()[-300:])}", flush=True)
