#!/bin/bash
# round 2, trip S (2 GPUs): window pipelines -- parity on one GPU, then the driver-style N=2 line with the per-rank pipelined e2e
set -u
OUT=gpurun_out/r02s; mkdir -p $OUT
timeout 600 python -m pytest tests/test_solve_host_gpu.py -x -q 2>&1 | tail -4 | tee $OUT/pytest_solve_host.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --no-weak > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo rc=$?
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02s/bench_n2.json').read().strip().splitlines()[-1])
print(j['value'], json.dumps(j['e2e'], indent=1)); print(j['parity']['ok'])
PY
tail -3 $OUT/bench_n2.err
