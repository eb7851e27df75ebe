#!/bin/bash
# Fixed per-step overhead hunt on one GPU: slab heights and the MULTI kernel with both slabs on GPU 0.
OUT=gpurun_out/${1:-sizes}; mkdir -p $OUT
run() { # label, args...
  local label=$1; shift
  local r=$(timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1)
  echo "$label $(echo "$r" | python -c 'import sys,json
try:
    j=json.loads(sys.stdin.read()); print(j["value"], j["ms_per_step"], j["roofline"]["achieved"], j["detail"]["kernel_variant"], j["clocks"]["reasons"])
except Exception as e: print("FAILED", e)')" | tee -a $OUT/sizes.txt
}
for ny in 16384 8192 4096 2048 1024; do run "single_ny$ny" --ny $ny; done
for n in 2 4 8; do run "local_slabs$n" --local-slabs $n; done
LBM_B200_KERNEL=ldg run "ldg_single_ny2048" --ny 2048
LBM_B200_KERNEL=ldg run "ldg_local_slabs2" --local-slabs 2
