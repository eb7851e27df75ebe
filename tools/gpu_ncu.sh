#!/bin/bash
# ncu --set full of the sweep kernel (one capture per change).  Usage: bash tools/gpu_ncu.sh <tag> [ENV=VAL ...]
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
env "$@" timeout 300 $CMD > $OUT/plain.log 2>&1 && \
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep -s 3 -c 1 -f -o $OUT/sweep $CMD > $OUT/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/plain.log
