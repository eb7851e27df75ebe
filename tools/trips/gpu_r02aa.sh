#!/bin/bash
# round 2, trip AA (1 GPU): staging-chunk size of the upload / download pipelines under the band pipeline
set -u
OUT=gpurun_out/r02aa; mkdir -p $OUT
for mb in 16 32 64 128 256; do
  echo "== stage $mb MB"; LBM_B200_STAGE_MB=$mb timeout 300 python tools/solve_host_probe.py 16384 20 2>&1 | tail -2 | tee -a $OUT/stage_mb.txt
done
