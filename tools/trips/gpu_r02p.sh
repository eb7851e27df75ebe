#!/bin/bash
# round 2, trip P (8 GPUs): block partition v3 on 4 and 8 GPUs -- parity, then the partition comparison
set -u
OUT=gpurun_out/r02p; mkdir -p $OUT
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -4 | tee $OUT/pytest_multi_gpu.txt
for args in "--in-process 8" "--in-process 8 --grid 4x2" "--in-process 8 --grid 2x4" "--in-process 4" "--in-process 4 --grid 2x2"; do
  timeout 600 python bench.py --steps 200 --warmup 20 $args 2>&1 | tail -1 | python -c '
import sys, json
j = json.loads(sys.stdin.read()); print(sys.argv[1], "|", j["detail"]["partition"], "|", j["ms_per_step"], "ms/step", j["value"], "MLUPS", j["roofline"]["achieved"], "GB/s per GPU")' "$args" | tee -a $OUT/partition_compare.txt
done
echo "== lbm_runner -n 8 --grid auto on the shipped 1024x1024 input"
mkdir -p $OUT/run8; timeout 600 lbm-graphcore_b200/bin/lbm_runner --params tests/golden/params/input_1024x1024.params --obstacles tests/golden/params/obstacles_1024x1024.dat -n 8 --grid auto -d --reruns 1 --outdir $OUT/run8 2>&1 | grep -E "Running LBM|Throughput|MLUPS" | tee -a $OUT/runner_n8_grid.txt
python tools/check.py --ref-av-vels-file <(zcat tests/golden/check/1024x1024.av_vels.dat.gz) --ref-final-state-file $OUT/run8/final_state.dat --av-vels-file $OUT/run8/av_vels.dat --final-state-file $OUT/run8/final_state.dat | tail -1 | tee -a $OUT/runner_n8_grid.txt
rm -rf $OUT/run8
