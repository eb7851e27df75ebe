#!/bin/bash
# round 2, trip C (8 GPUs): PCIe ceiling at 1/2/4/8 GPUs, multi-GPU parity tests with the two-phase teardown, the driver-style
# N=8 bench line (parity + weak records), NVLink byte counters around a known number of timesteps
set -u
OUT=gpurun_out/r02c; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== pcie probe"; timeout 400 tools/probes/bin/pcie_probe 2147483648 2 2>&1 | tee $OUT/pcie_probe.txt
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -8 | tee $OUT/pytest_multi_gpu.txt
echo "== bench N=8 (driver style)"
nvidia-smi nvlink -gt d > $OUT/nvlink_before.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 200 --warmup 20 --no-weak --no-parity --no-e2e > $OUT/bench_n8_kernel_only.json 2> $OUT/bench_n8_kernel_only.err; echo rc=$?
nvidia-smi nvlink -gt d > $OUT/nvlink_after.txt 2>&1
tail -c 1500 $OUT/bench_n8_kernel_only.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo rc=$?
tail -c 4000 $OUT/bench_n8.json; tail -3 $OUT/bench_n8.err
