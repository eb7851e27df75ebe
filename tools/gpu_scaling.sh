#!/bin/bash
# 8-GPU box: multi-GPU parity tests, strong scaling 1/2/4/8 on the 16384^2 channel, weak scaling 32768^2 per GPU at 8
OUT=gpurun_out/${1:-scaling}; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_multi_gpu.txt
for n in 1 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n \
     bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline > $OUT/bench_n$n.log 2>&1
  echo "strong n=$n rc=$? $(tail -1 $OUT/bench_n$n.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["value"], j["ms_per_step"], j["roofline"]["achieved"], (j.get("e2e") or {}).get("value"))' 2>&1 | tail -1)"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 \
   bench.py --gpus 8 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --workload weak32768 > $OUT/bench_weak8.log 2>&1
echo "weak n=8 rc=$? $(tail -1 $OUT/bench_weak8.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["value"], j["ms_per_step"], j["roofline"]["achieved"])' 2>&1 | tail -1)"
