#!/bin/bash
# per-step latency on the reference's small grid sizes (synthetic channel of that size)
OUT=gpurun_out/${1:-small}; mkdir -p $OUT
run() { # label, env... -- args
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs+=("$1"); shift; done; shift
  local r=$(env "${envs[@]}" timeout 300 python bench.py --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1)
  echo "$label $(echo "$r" | python -c 'import sys,json
try:
    j=json.loads(sys.stdin.read()); print(j["detail"]["kernel_variant"], j["value"], "MLUPS", round(j["ms_per_step"]*1000,3), "us/step", j["roofline"]["achieved"], "GB/s")
except Exception as e: print("FAILED", e)')" | tee -a $OUT/small.txt
}
for n in 128 256 512 1024; do
  run "persistent_${n}x${n}" X=1 -- --nx $n --ny $n --steps 20000 --warmup 1000
  for v in lbm-graphcore_b200/exp/lib_pm*.so; do [ -f $v ] && run "persistent_$(basename $v)_${n}x${n}" LBM_B200_LIB=$PWD/$v -- --nx $n --ny $n --steps 20000 --warmup 1000; done
  run "perstep_pdl_${n}x${n}" LBM_B200_PERSISTENT=0 -- --nx $n --ny $n --steps 20000 --warmup 1000
done
run "forced_persistent_4096" LBM_B200_PERSISTENT=1 -- --nx 4096 --ny 4096 --steps 300 --warmup 20
run "perstep_4096" X=1 -- --nx 4096 --ny 4096 --steps 300 --warmup 20
