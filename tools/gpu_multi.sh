#!/bin/bash
# Multi-GPU trip (gpurun --gpus N): parity tests incl. tests/test_multi_gpu.py, then bench.py at N GPUs.
# Usage: bash tools/gpu_multi.sh <tag> <N> [extra bench args]
TAG=${1:-multi}; N=${2:-2}; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
for n in 1 $N; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline "$@" > $OUT/bench_n$n.log 2>&1
  echo "bench n=$n rc=$?"; tail -1 $OUT/bench_n$n.log | cut -c1-1500
done
