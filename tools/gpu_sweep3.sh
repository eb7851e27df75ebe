#!/bin/bash
OUT=gpurun_out/${1:-sweep}; mkdir -p $OUT
run() { # label, env...
  local label=$1; shift
  local r=$(env "$@" timeout 300 python bench.py --steps 60 --warmup 10 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$label $(echo "$r" | python -c 'import sys,json
try:
    j=json.loads(sys.stdin.read()); print(j["value"], j["roofline"]["achieved"], j["config"]["kernel_variant"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e: print("FAILED", e)')" | tee -a $OUT/sweep.txt
}
nvidia-smi --query-gpu=name,clocks.mem,clocks.max.mem,temperature.gpu,power.draw,serial --format=csv | tee -a $OUT/sweep.txt
OLD=$PWD/lbm-graphcore_b200/exp/lib_libdiv.so
for rep in 1 2; do
run ldg_new LBM_B200_KERNEL=ldg
run ldg_old LBM_B200_KERNEL=ldg LBM_B200_LIB=$OLD
run ldg_new_b3 LBM_B200_KERNEL=ldg LBM_B200_BLOCKS_PER_SM=3
run ldg_new_b2 LBM_B200_KERNEL=ldg LBM_B200_BLOCKS_PER_SM=2
run tma_new_s3b1 LBM_B200_STAGES=3 LBM_B200_BLOCKS_PER_SM=1
run tma_old_s3b1 LBM_B200_STAGES=3 LBM_B200_BLOCKS_PER_SM=1 LBM_B200_LIB=$OLD
run tma_new_s3b2 LBM_B200_STAGES=3 LBM_B200_BLOCKS_PER_SM=2
run tma_old_s3b2 LBM_B200_STAGES=3 LBM_B200_BLOCKS_PER_SM=2 LBM_B200_LIB=$OLD
done
