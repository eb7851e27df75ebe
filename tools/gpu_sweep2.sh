#!/bin/bash
OUT=gpurun_out/${1:-sweep2}; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
run() { # label, env...
  local label=$1; shift
  local r=$(env "$@" timeout 300 python bench.py --steps 60 --warmup 10 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$label $(echo "$r" | python -c 'import sys,json
try:
    j=json.loads(sys.stdin.read()); print(j["value"], j["roofline"]["achieved"], j["config"]["kernel_variant"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e: print("FAILED", e)')" | tee -a $OUT/sweep.txt
}
run tma_default X=1
for st in 2 3 4 5 6; do for b in 1 2 3; do run "tma_s${st}_b${b}" LBM_B200_STAGES=$st LBM_B200_BLOCKS_PER_SM=$b; done; done
run ldg LBM_B200_KERNEL=ldg
