#!/bin/bash
# round 2, trip I (2 GPUs): the driver-style N=2 line after the bench.py changes (parity, weak, e2e), reference arm under torchrun
set -u
OUT=gpurun_out/r02i; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo rc=$?
tail -c 3500 $OUT/bench_n2.json; tail -3 $OUT/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
