#!/usr/bin/env python
"""Does cutting an L2-resident grid into several row slabs ON THE SAME GPU help?  No: slabs that share a GPU share one stream
(lbm_create_on_devices), so each step becomes several back-to-back launches without the PDL overlap between timesteps.
Prints microseconds per timestep for 1 / 2 / 3 / 4 slabs on device 0 and checks the lattices against the 1-slab run (run on a GPU box)."""
import hashlib
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    lbm = importlib.import_module("lbm-graphcore_b200")
    for n in [int(a) for a in sys.argv[1:]] or [1024, 768, 2048]:
        steps = max(200, 2000 * 1024 * 1024 // (n * n))
        want = None
        for slabs in (1, 2, 3, 4):
            p = lbm.Params(n, n, 200 + 3 * steps, 10, 0.1, 0.005, 1.85)
            with lbm.Solver(p, devices=[0] * slabs) as s:
                s.init_synthetic(1234, 0.01)
                s.run(200)
                best = 1e30
                for _ in range(3):
                    s.run(steps)
                    best = min(best, s.elapsed_ms)
                cells, _ = s.download()
                digest = hashlib.sha256(np.ascontiguousarray(cells).tobytes()).hexdigest()[:16]
                want = want or digest
                print(f"{n:5d}^2 {slabs} slab(s) {s.kernel_variant:22s} {best * 1e3 / steps:7.3f} us/step "
                      f"{n * n * steps / (best * 1e-3) / 1e9:6.1f} GLUPS  lattice {'same' if digest == want else 'DIFFERENT'}", flush=True)


if __name__ == "__main__":
    main()
