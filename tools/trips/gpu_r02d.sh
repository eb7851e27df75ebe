#!/bin/bash
# round 2, trip D (1 GPU): 2-D block partition parity (all blocks on one GPU), cluster kernel incl. 256-cell CTAs, small-grid timings,
# A/B of a 16384x2048 slab alone vs as one of several slabs (what costs the 2 % at N=8), then the whole GPU suite
set -u
OUT=gpurun_out/r02d; mkdir -p $OUT
echo "== block tests"; timeout 900 python -m pytest tests/test_blocks_gpu.py -x -q 2>&1 | tail -25 | tee $OUT/pytest_blocks.txt
echo "== cluster tests"; timeout 900 python -m pytest tests/test_cluster_gpu.py -x -q 2>&1 | tail -8 | tee $OUT/pytest_cluster.txt
echo "== small grids"; timeout 900 python tools/small_grids.py 128x128 128x256 256x256 2>&1 | tee $OUT/small_grids.txt
echo "== slab A/B"
for args in "--ny 2048" "--ny 4096 --local-slabs 2" "--ny 16384 --local-slabs 8" ""; do
  timeout 600 python bench.py --nx 16384 $args --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-parity --no-weak --no-shipped 2>&1 | tail -1 | python -c '
import sys, json
j = json.loads(sys.stdin.read()); print(sys.argv[1:], j["ms_per_step"], "ms/step", j["value"], "MLUPS", j["roofline"]["achieved"], "GB/s", j["roofline"]["kernel"])' "$args" | tee -a $OUT/slab_ab.txt
done
echo "== full GPU suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
