#!/bin/bash
# round 2, trip O (2 GPUs, one process): block partition v3 (ghost columns / corners synchronised at launch granularity)
set -u
OUT=gpurun_out/r02o; mkdir -p $OUT
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_multi_gpu.py tests/test_prepared_gpu.py -x -q 2>&1 | tail -3 | tee $OUT/pytest_blocks_v3.txt
LBM_B200_LIB=$PWD/lbm-graphcore_b200/liblbm_b200_boundscheck.so timeout 600 python -m pytest tests/test_blocks_gpu.py -x -q 2>&1 | tail -2
for args in "--in-process 2" "--in-process 2 --grid 2x1" "--in-process 2 --grid 1x2"; do
  timeout 600 python bench.py --steps 200 --warmup 20 $args 2>&1 | tail -1 | python -c '
import sys, json
j = json.loads(sys.stdin.read()); print(sys.argv[1], "|", j["detail"]["partition"], "|", j["ms_per_step"], "ms/step", j["value"], "MLUPS")' "$args" | tee -a $OUT/compare_v3.txt
done
