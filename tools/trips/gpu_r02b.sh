#!/bin/bash
# round 2, trip B: the cluster-resident kernel -- parity, then microseconds per step of every small-grid strategy
set -u
OUT=gpurun_out/r02b; mkdir -p $OUT
echo "== cluster tests"; timeout 900 python -m pytest tests/test_cluster_gpu.py -x -q 2>&1 | tail -25 | tee $OUT/pytest_cluster.txt
echo "== small grids"; timeout 900 python tools/small_grids.py 128x128 128x256 256x256 2>&1 | tee $OUT/small_grids.txt
echo "== phase profile"
for shape in "1024 4" "1024 2" "512 2" "512 1"; do set -- $shape
  for case in 128x128 256x256; do
    LBM_B200_LIB=$PWD/lbm-graphcore_b200/liblbm_b200_prof.so LBM_B200_CLUSTER_CELLS=$1 LBM_B200_CLUSTER_CPT=$2 timeout 120 python tools/small_grids.py --child $case 20000 2>&1 | grep "cluster profile" | tail -1 | sed "s/^/$case /" | tee -a $OUT/cluster_phases.txt
  done
done
