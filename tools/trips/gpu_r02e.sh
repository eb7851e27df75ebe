#!/bin/bash
# round 2, trip E (1 GPU): whole GPU suite, driver-style bench, then ncu: launch list of the bench, full captures of the fused
# compute+halo kernel (as 8 / 2 slabs sharing the GPU: the per-GPU launch of an 8- / 2-GPU run), of the per-step kernel on the
# L2-resident 1024^2 grid, and of the cluster-resident kernel
set -u
OUT=gpurun_out/r02e; mkdir -p $OUT
echo "== full GPU suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench (driver style)"; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo rc=$?; tail -c 3000 $OUT/bench.json; tail -3 $OUT/bench.err
B="python bench.py --no-e2e --no-cpu-baseline --no-parity --no-weak --no-shipped"
cap() {  # cap <name> <kernel regex> <skip> <env...> -- <command...>
  local name=$1 rx=$2 skip=$3; shift 3
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 "$@" > $OUT/$name.plain.log 2>&1 || { echo "$name: plain run failed"; tail -3 $OUT/$name.plain.log; return; }
  env "${envs[@]}" timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $OUT/$name "$@" > $OUT/$name.ncu.log 2>&1
  echo "$name ncu rc=$?"
  ncu -i $OUT/$name.ncu-rep --page raw --csv > $OUT/$name.raw.csv 2>/dev/null
  ncu -i $OUT/$name.ncu-rep --page details > $OUT/$name.details.txt 2>/dev/null
}
echo "== ncu launch list"
timeout 600 $B --steps 3 --warmup 3 > $OUT/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $B --steps 3 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
cap sweep_single sweep_kernel 3 X=1 -- $B --steps 3 --warmup 3
cap sweep_multi8 sweep_kernel 30 X=1 -- $B --steps 3 --warmup 3 --local-slabs 8
cap sweep_multi2 sweep_kernel 8 X=1 -- $B --steps 3 --warmup 3 --local-slabs 2
cap sweep_1024 sweep_kernel 50 LBM_B200_PERSISTENT=0 LBM_B200_GRAPH=0 -- $B --nx 1024 --ny 1024 --steps 100 --warmup 100
cap cluster_128 sweep_cluster 0 X=1 -- python tools/small_grids.py --child 128x128 2000
cap cluster_256 sweep_cluster 0 X=1 -- python tools/small_grids.py --child 256x256 2000
rm -f $OUT/sweep_multi2.ncu-rep $OUT/sweep_single.ncu-rep $OUT/cluster_256.ncu-rep
ls -la $OUT
