#!/bin/bash
# round 2, trip G (8 GPUs): block partition across GPUs (2x2, 2x4, 4x2) -- parity, then the same 16384^2 channel as 8 row slabs vs
# 4x2 vs 2x4 blocks (one process driving all GPUs for all three, so only the partition differs), then the rank-mode strong-scaling line
set -u
OUT=gpurun_out/r02g; mkdir -p $OUT
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -6 | tee $OUT/pytest_multi_gpu.txt
B="python bench.py --steps 200 --warmup 20"
for args in "--in-process 8" "--in-process 8 --grid 4x2" "--in-process 8 --grid 2x4" "--in-process 4" "--in-process 4 --grid 2x2" "--in-process 2" "--in-process 2 --grid 2x1"; do
  timeout 600 $B $args 2>&1 | tail -1 | python -c '
import sys, json
j = json.loads(sys.stdin.read()); print(sys.argv[1], "|", j["detail"]["partition"], "|", j["ms_per_step"], "ms/step", j["value"], "MLUPS", j["roofline"]["achieved"], "GB/s per GPU")' "$args" | tee -a $OUT/partition_compare.txt
done
echo "== lbm_runner -n 8 on the shipped 1024x1024 input: row slabs vs --grid auto"
for extra in "" "--grid auto"; do
  mkdir -p $OUT/run8; timeout 600 lbm-graphcore_b200/bin/lbm_runner --params tests/golden/params/input_1024x1024.params --obstacles tests/golden/params/obstacles_1024x1024.dat -n 8 $extra --stats -d --reruns 1 --outdir $OUT/run8 2>&1 | grep -E "Running LBM|Throughput|numGpus|^8,|Free-cell|MLUPS|Kernel" | sed "s/^/[$extra] /" | tee -a $OUT/runner_n8.txt
  python tools/check.py --ref-av-vels-file <(zcat tests/golden/check/1024x1024.av_vels.dat.gz) --ref-final-state-file $OUT/run8/final_state.dat --av-vels-file $OUT/run8/av_vels.dat --final-state-file $OUT/run8/final_state.dat | tail -1 | tee -a $OUT/runner_n8.txt
done
rm -rf $OUT/run8
