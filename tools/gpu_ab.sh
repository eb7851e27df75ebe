#!/bin/bash
# A/B of builds under lbm-graphcore_b200/exp/ on the headline grid and on two L2-resident grids
OUT=gpurun_out/${1:-ab}; mkdir -p $OUT
run() { # label lib args...
  local label=$1 lib=$2; shift; shift
  local r=$(LBM_B200_LIB=$lib timeout 300 python bench.py --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1)
  echo "$label $(echo "$r" | python -c 'import sys,json
try:
    j=json.loads(sys.stdin.read()); print(j["detail"]["kernel_variant"], j["value"], "MLUPS", round(j["ms_per_step"]*1000,2), "us/step", j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e: print("FAILED", e)')" | tee -a $OUT/ab.txt
}
for rep in 1 2; do
  for v in lbm-graphcore_b200/exp/lib_*.so; do
    run "$(basename $v .so)_16384" $PWD/$v --steps 100 --warmup 20
  done
done
for v in lbm-graphcore_b200/exp/lib_*.so; do
  run "$(basename $v .so)_1024" $PWD/$v --nx 1024 --ny 1024 --steps 20000 --warmup 1000
  run "$(basename $v .so)_256" $PWD/$v --nx 256 --ny 256 --steps 20000 --warmup 1000
done
