#!/usr/bin/env python
"""Microseconds per timestep of L2-resident single-slab grids (synthetic channel) under the persistent cooperative kernel and
under per-step launches replayed from the cached CUDA graph (run on a GPU box)."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = [("persistent", {"LBM_B200_PERSISTENT": "1"}), ("perstep_graph", {"LBM_B200_PERSISTENT": "0", "LBM_B200_GRAPH": "1"}),
            ("perstep", {"LBM_B200_PERSISTENT": "0", "LBM_B200_GRAPH": "0"}), ("default", {})]


def child(n, steps):
    lbm = importlib.import_module("lbm-graphcore_b200")
    p = lbm.Params(n, n, 200 + 3 * steps, 10, 0.1, 0.005, 1.85)
    with lbm.Solver(p) as s:
        s.init_synthetic(1234, 0.01)
        s.run(200)
        best = 1e30
        for _ in range(3):
            s.run(steps)
            best = min(best, s.elapsed_ms)
        print(f"{s.kernel_variant} {best * 1e3 / steps:.3f} us/step {n * n * steps / (best * 1e-3) / 1e6:.1f} MLUPS", flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    for n in (384, 512, 768, 1024, 1536, 2048, 3072):
        for name, env in VARIANTS:
            e = dict(os.environ)
            e.update(env)
            r = subprocess.run([sys.executable, __file__, "--child", str(n), str(max(200, 4000 * 512 * 512 // (n * n)))], capture_output=True,
                               text=True, env=e, timeout=600)
            print(f"{n:5d}^2 {name:14s} {r.stdout.strip() or ('FAILED: ' + r.stderr.strip()[-300:])}", flush=True)
