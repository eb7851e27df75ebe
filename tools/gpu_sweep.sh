#!/bin/bash
# Kernel tuning sweep on one B200: each line = one kernel-only bench run (MLUPS, GB/s).
OUT=gpurun_out/${1:-sweep}; mkdir -p $OUT
run() { # label, env...
  local label=$1; shift
  local r=$(env "$@" timeout 300 python bench.py --steps 60 --warmup 10 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$label $(echo "$r" | python -c 'import sys,json
try:
    j=json.loads(sys.stdin.read()); print(j["value"], j["roofline"]["achieved"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e: print("FAILED", e)')" | tee -a $OUT/sweep.txt
}
for v in lbm-graphcore_b200/exp/lib_*.so; do run "$(basename $v)" LBM_B200_LIB=$PWD/$v; done
L=$PWD/lbm-graphcore_b200/exp/lib_h0b3.so
for ri in 1 2 4 8 16 32; do run "h0b3_rows$ri" LBM_B200_LIB=$L LBM_B200_ROWS_ITER=$ri; done
for b in 1 2; do run "h0b3_blocks$b" LBM_B200_LIB=$L LBM_B200_BLOCKS_PER_SM=$b; done
run "h0b2_blocks1" LBM_B200_LIB=$PWD/lbm-graphcore_b200/exp/lib_h0b2.so LBM_B200_BLOCKS_PER_SM=1
