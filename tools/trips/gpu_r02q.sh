#!/bin/bash
# round 2, trip Q (1 GPU): band height of the host pipeline; then the whole GPU suite, smoke and the driver-style bench on the final code
set -u
OUT=gpurun_out/r02q; mkdir -p $OUT
for rows in 512 768 1024 1536 2048; do
  echo "== band rows $rows"; LBM_B200_BAND_ROWS=$rows timeout 300 python tools/solve_host_probe.py 16384 20 2>&1 | tail -2 | tee -a $OUT/band_rows.txt
done
echo "== full GPU suite"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench (driver style)"; timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo rc=$?; tail -c 1500 $OUT/bench.json; tail -3 $OUT/bench.err
