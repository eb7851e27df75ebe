#!/bin/bash
# round 2, trip U (8 GPUs): the N=8 line with the per-rank pipelined e2e (kept short: the round's GPU budget is nearly spent)
set -u
OUT=gpurun_out/r02u; mkdir -p $OUT
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 --no-weak > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo rc=$?
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02u/bench_n8.json').read().strip().splitlines()[-1])
e=j['e2e']; print(j['value'], e['value'], e['job'][:150], e.get('pcie_gbs_both_ways_per_gpu'), e.get('equals_sequential_job'), e.get('window_pipeline_skipped'), 'seq', e.get('sequential_job', {}).get('value'), 'parity', j['parity']['ok'])
PY
tail -2 $OUT/bench_n8.err
