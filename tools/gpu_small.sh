#!/bin/bash
# per-step latency on the reference's small grid sizes (synthetic channel of that size), PDL on/off
OUT=gpurun_out/${1:-small}; mkdir -p $OUT
run() { # label, env... -- args
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs+=("$1"); shift; done; shift
  local r=$(env "${envs[@]}" timeout 300 python bench.py --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1)
  echo "$label $(echo "$r" | python -c 'import sys,json
try:
    j=json.loads(sys.stdin.read()); print(j["value"], "MLUPS", round(j["ms_per_step"]*1000,3), "us/step", j["roofline"]["achieved"], "GB/s")
except Exception as e: print("FAILED", e)')" | tee -a $OUT/small.txt
}
for pdl in 1 0; do
  for n in 128 256 1024 4096; do
    run "pdl${pdl}_${n}x${n}" LBM_B200_PDL=$pdl -- --nx $n --ny $n --steps 20000 --warmup 1000
  done
  run "pdl${pdl}_16384" LBM_B200_PDL=$pdl -- --steps 100 --warmup 10
done
