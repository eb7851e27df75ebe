#!/bin/bash
# Kernel tuning sweep on one B200: each line = one kernel-only bench run of a build under lbm-graphcore_b200/exp/
OUT=gpurun_out/${1:-sweep}; mkdir -p $OUT
run() { # label, env...
  local label=$1; shift
  local r=$(env "$@" timeout 300 python bench.py --steps 100 --warmup 20 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$label $(echo "$r" | python -c 'import sys,json
try:
    j=json.loads(sys.stdin.read()); print(j["value"], j["ms_per_step"], j["roofline"]["achieved"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e: print("FAILED", e)')" | tee -a $OUT/sweep.txt
}
for v in lbm-graphcore_b200/exp/lib_*.so; do run "$(basename $v)" LBM_B200_LIB=$PWD/$v; done
