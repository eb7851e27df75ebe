#!/bin/bash
# round 2, trip T (4 GPUs): the driver-style N=4 line with the per-rank pipelined e2e
set -u
OUT=gpurun_out/r02t; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/bench_n4.json 2> $OUT/bench_n4.err; echo rc=$?
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02t/bench_n4.json').read().strip().splitlines()[-1])
e=j['e2e']; print(j['value'], e['value'], e['job'][:120], e.get('pcie_gbs_both_ways_per_gpu'), e['equals_sequential_job'], 'seq', e['sequential_job']['value'], e['sequential_job']['h2d_gbs_per_gpu'], e['sequential_job']['d2h_gbs_per_gpu'], 'parity', j['parity']['ok'], 'weak', j['weak']['mlups'])
PY
tail -2 $OUT/bench_n4.err
