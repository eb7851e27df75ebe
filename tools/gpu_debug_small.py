"""development: one small per-step (non-persistent) run with per-launch synchronisation, printing av_vels vs oracle"""
import importlib
import os
import sys

import numpy as np

os.environ.setdefault("LBM_B200_PERSISTENT", "0")
os.environ.setdefault("LBM_B200_DEBUG_SYNC", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lbm = importlib.import_module("lbm-graphcore_b200")
from oracle import oracle  # noqa: E402

nx, ny, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(0)
p = lbm.Params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
obst = (rng.random((ny, nx)) < 0.05).astype(np.uint8)
cells = (lbm.initial_cells(p) * (1 + 0.1 * rng.random((ny, nx, 9)))).astype(np.float32)
want, wav = oracle.run(op, cells, obst, iters)
with lbm.Solver(p) as s:
    print("variant", s.kernel_variant, flush=True)
    s.upload(cells, obst)
    s.run(iters)
    got, av = s.download()
print("lattice mismatches", int((got.view(np.uint32) != want.view(np.uint32)).sum()))
print("av ratio", np.round(av / wav, 4))
