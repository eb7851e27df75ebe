#!/bin/bash
set -u
OUT=gpurun_out/r02h; mkdir -p $OUT
echo "== solve_host tests"; timeout 900 python -m pytest tests/test_solve_host_gpu.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_solve_host.txt
for w in 3 2 4; do
echo "== workers=$w"; LBM_B200_SOLVE_WORKERS=$w timeout 600 python tools/solve_host_probe.py 16384 20 2>&1 | tail -4 | tee $OUT/probe_w$w.txt
done
echo "== trace workers=3"; LBM_B200_SOLVE_TRACE=1 timeout 600 python tools/solve_host_probe.py 16384 20 2>&1 | tail -18 | tee $OUT/trace_w3.txt
echo "== bench (driver style)"; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-weak --no-shipped --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo rc=$?
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02h/bench.json').read().strip().splitlines()[-1])
print(json.dumps(j['e2e'], indent=1))
PY
tail -3 $OUT/bench.err
