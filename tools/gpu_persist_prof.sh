#!/bin/bash
# development: per-step latency of the persistent small-grid kernel for builds under lbm-graphcore_b200/exp/
# (LBM_PERSIST_MINB variants) against the one-launch-per-step path
OUT=gpurun_out/${1:-persist_prof}; mkdir -p $OUT
one() { # label env... -- n
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  local n=$1
  env "${envs[@]}" python bench.py --no-e2e --no-cpu-baseline --nx $n --ny $n --steps 20000 --warmup 1000 > $OUT/tmp.log 2>&1
  local us=$(tail -1 $OUT/tmp.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["detail"]["kernel_variant"], round(j["ms_per_step"]*1000,2), "us/step", j["value"], "MLUPS")' 2>/dev/null)
  echo "$label ${n}x${n}: $us | $(grep 'persistent profile' $OUT/tmp.log | tail -1)" | tee -a $OUT/prof.txt
}
for n in 128 256 512 1024; do
  one default_2_per_sm X=1 -- $n
  for v in lbm-graphcore_b200/exp/lib_pm*.so; do [ -f $v ] && one "$(basename $v .so)" LBM_B200_LIB=$PWD/$v -- $n; done
  one perstep LBM_B200_PERSISTENT=0 -- $n
done
