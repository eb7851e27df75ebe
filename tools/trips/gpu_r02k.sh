#!/bin/bash
# round 2, trip K (1 GPU): bounds-checked build over the parity suites; then the whole GPU suite and the driver-style bench on the final code
set -u
OUT=gpurun_out/r02k; mkdir -p $OUT
echo "== bounds-checked build"; bash tools/gpu_boundscheck.sh r02k
echo "== full GPU suite"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench (driver style)"; timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo rc=$?; tail -c 5000 $OUT/bench.json; tail -3 $OUT/bench.err
