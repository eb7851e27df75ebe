#!/bin/bash
# round 2, trip A: new tests (save/load, graph replay, runner flags), the driver-style bench line, the reference arm, PCIe ceiling
set -u
OUT=gpurun_out/r02a; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== bench (driver style)"; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo rc=$?; tail -c 6000 $OUT/bench.json; tail -5 $OUT/bench.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== pcie probe"; timeout 300 tools/probes/bin/pcie_probe 4294967296 3 2>&1 | tee $OUT/pcie_probe_n1.txt
