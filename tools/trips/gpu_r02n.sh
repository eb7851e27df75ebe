#!/bin/bash
# round 2, trip N (8 GPUs): final multi-GPU validation -- all multi-GPU tests (row slabs, rank mode, blocks v2 on 2/4/8 GPUs), the
# partition comparison with the v2 block kernel, the driver-style N=8 and N=4 lines
set -u
OUT=gpurun_out/r02n; mkdir -p $OUT
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -4 | tee $OUT/pytest_multi_gpu.txt
for args in "--in-process 8" "--in-process 8 --grid 4x2" "--in-process 8 --grid 2x4" "--in-process 4" "--in-process 4 --grid 2x2"; do
  timeout 600 python bench.py --steps 200 --warmup 20 $args 2>&1 | tail -1 | python -c '
import sys, json
j = json.loads(sys.stdin.read()); print(sys.argv[1], "|", j["detail"]["partition"], "|", j["ms_per_step"], "ms/step", j["value"], "MLUPS", j["roofline"]["achieved"], "GB/s per GPU")' "$args" | tee -a $OUT/partition_compare.txt
done
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err; echo "N=$n rc=$?"
  python - <<PY
import json
j=json.loads(open('gpurun_out/r02n/bench_n$n.json').read().strip().splitlines()[-1])
print(j['n_gpus'], j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['traffic'], 'e2e', j['e2e']['value'], j['e2e']['h2d_gbs_per_gpu'], j['e2e']['d2h_gbs_per_gpu'], 'parity', j['parity']['ok'], 'weak', j['weak']['mlups'])
PY
done
