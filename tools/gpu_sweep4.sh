#!/bin/bash
OUT=gpurun_out/${1:-sweep}; mkdir -p $OUT
run() { # label, env... -- args
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs+=("$1"); shift; done; shift
  local r=$(env "${envs[@]}" timeout 300 python bench.py --steps 60 --warmup 10 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1)
  echo "$label $(echo "$r" | python -c 'import sys,json
try:
    j=json.loads(sys.stdin.read()); print(j["value"], j["ms_per_step"], j["roofline"]["achieved"], j["config"]["kernel_variant"], j["clocks"]["reasons"])
except Exception as e: print("FAILED", e)')" | tee -a $OUT/sweep.txt
}
for pr in 0 1 2 3; do
  run "promo${pr}_b1" LBM_B200_L2PROMO=$pr --
  run "promo${pr}_b2_s3" LBM_B200_L2PROMO=$pr LBM_B200_BLOCKS_PER_SM=2 LBM_B200_STAGES=3 --
  run "promo${pr}_b2_s2" LBM_B200_L2PROMO=$pr LBM_B200_BLOCKS_PER_SM=2 LBM_B200_STAGES=2 --
  run "promo${pr}_multi2" LBM_B200_L2PROMO=$pr -- --local-slabs 2
done
