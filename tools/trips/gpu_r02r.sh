#!/bin/bash
# round 2, trip R (1 GPU): cluster kernel for nx = 512 -- parity, timing against the per-step route, then the whole suite again
set -u
OUT=gpurun_out/r02r; mkdir -p $OUT
echo "== cluster tests"; timeout 900 python -m pytest tests/test_cluster_gpu.py -x -q 2>&1 | tail -3 | tee $OUT/pytest_cluster.txt
for n in 512; do for v in "default" "LBM_B200_CLUSTER=0" "LBM_B200_CLUSTER_CELLS=1024"; do
  if [ "$v" = default ]; then e=X=1; else e=$v; fi
  env $e timeout 300 python tools/mid_grids.py --child $n 4000 2>&1 | tail -1 | sed "s/^/$n^2 [$v] /" | tee -a $OUT/cluster_512.txt
done; done
for shape in "512 256"; do set -- $shape
python - <<PY | tee -a $OUT/cluster_512.txt
import importlib, sys
sys.path.insert(0, '.')
lbm = importlib.import_module("lbm-graphcore_b200")
import os
for env in ({}, {"LBM_B200_CLUSTER": "0"}):
    os.environ.pop("LBM_B200_CLUSTER", None); os.environ.update(env)
    p = lbm.Params($1, $2, 200 + 3 * 8000, 10, 0.1, 0.005, 1.85)
    with lbm.Solver(p) as s:
        s.init_synthetic(1234, 0.01); s.run(200); best = 1e9
        for _ in range(3):
            s.run(8000); best = min(best, s.elapsed_ms)
        print("$1x$2", s.kernel_variant, round(best * 1e3 / 8000, 3), "us/step", round($1 * $2 * 8000 / (best * 1e-3) / 1e6, 1), "MLUPS")
PY
done
echo "== full GPU suite"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
