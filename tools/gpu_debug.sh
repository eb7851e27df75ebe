#!/bin/bash
cat > /tmp/tiny.py <<'PY'
import importlib, sys, numpy as np
sys.path.insert(0, '.')
lbm = importlib.import_module("lbm-graphcore_b200")
from oracle import oracle
nx, ny, it = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
p = lbm.Params(nx, ny, it, 10, 0.1, 0.005, 1.85)
op = oracle.make_params(nx, ny, it, 10, 0.1, 0.005, 1.85)
rng = np.random.default_rng(0)
obst = (rng.random((ny, nx)) < 0.05).astype(np.uint8)
cells = (lbm.initial_cells(p) * (1 + 0.1 * rng.random((ny, nx, 9)))).astype(np.float32)
with lbm.Solver(p) as s:
    print("variant", s.kernel_variant, flush=True)
    s.upload(cells, obst); s.run(it)
    got, av = s.download()
want, wav = oracle.run(op, cells, obst, it)
print("mismatches", int((got.view(np.uint32) != want.view(np.uint32)).sum()), "of", got.size, flush=True)
PY
echo "== tma tiny"; LBM_B200_DEBUG_SYNC=1 timeout 120 python /tmp/tiny.py 256 16 3 2>&1 | grep -v "^  File\|^    " | tail -3
