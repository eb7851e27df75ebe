#!/bin/bash
# One GPU-box trip: parity tests, smoke, bench, then ncu (launch list + one full capture of the sweep).
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench (kernel only, quick)"
timeout 600 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | tail -2 | tee $OUT/bench_quick.json
echo "== bench (full default)"
timeout 1200 python bench.py 2>&1 | tail -2 | tee $OUT/bench.json
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== ncu launch list"
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 $CMD > $OUT/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
echo "== ncu full (sweep kernel)"
timeout 600 $CMD > $OUT/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -f -o $OUT/sweep $CMD > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
