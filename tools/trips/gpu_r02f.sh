#!/bin/bash
# round 2, trip F: lean collide path of the cluster kernel -- parity, then A/B against the build without it
set -u
OUT=gpurun_out/r02f; mkdir -p $OUT
echo "== cluster tests (lean)"; timeout 900 python -m pytest tests/test_cluster_gpu.py tests/test_full_runs_gpu.py -x -q 2>&1 | tail -5 | tee $OUT/pytest_cluster.txt
for lib in liblbm_b200.so liblbm_b200_nolean.so; do
  for case in 128x128 128x256 256x256; do
    for rep in 1 2; do
      LBM_B200_LIB=$PWD/lbm-graphcore_b200/$lib timeout 120 python tools/small_grids.py --child $case 20000 2>&1 | tail -1 | sed "s/^/$lib $case /" | tee -a $OUT/lean_ab.txt
    done
  done
done
