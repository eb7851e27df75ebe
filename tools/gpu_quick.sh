#!/bin/bash
# Quick GPU trip: parity tests + kernel-only bench.  Usage: bash tools/gpu_quick.sh [tag] [extra bench args]
TAG=${1:-quick}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
timeout 600 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -2 | tee $OUT/bench_quick.json
timeout 600 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --local-slabs 2 "$@" 2>&1 | tail -1 | cut -c1-200 | tee $OUT/bench_slabs2.json
free -g | head -2 | tee $OUT/host_mem.txt; nproc | tee -a $OUT/host_mem.txt
