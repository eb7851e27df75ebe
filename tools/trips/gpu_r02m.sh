#!/bin/bash
# round 2, trip M2 (2 GPUs, one process): block partition with per-tile-row ghost-column flags and a 2-D launch grid
set -u
OUT=gpurun_out/r02m; mkdir -p $OUT
timeout 600 python -m pytest tests/test_blocks_gpu.py tests/test_multi_gpu.py -x -q 2>&1 | tail -3 | tee $OUT/pytest_blocks_v2.txt
LBM_B200_LIB=$PWD/lbm-graphcore_b200/liblbm_b200_boundscheck.so timeout 600 python -m pytest tests/test_blocks_gpu.py -x -q 2>&1 | tail -2
for args in "--in-process 2" "--in-process 2 --grid 2x1" "--in-process 2 --grid 1x2"; do
  timeout 600 python bench.py --steps 200 --warmup 20 $args 2>&1 | tail -1 | python -c '
import sys, json
j = json.loads(sys.stdin.read()); print(sys.argv[1], "|", j["detail"]["partition"], "|", j["ms_per_step"], "ms/step", j["value"], "MLUPS")' "$args" | tee -a $OUT/compare_v2.txt
done
B="python bench.py --in-process 2 --grid 2x1 --steps 3 --warmup 3"
M=nvltx__bytes.sum,nvltx__bytes_data_user.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:sweep_kernel -s 6 -c 2 --csv --log-file $OUT/blocks_v2_nvlink.csv $B > $OUT/blocks_v2_nvlink.log 2>&1; echo rc=$?
grep -v "^==" $OUT/blocks_v2_nvlink.csv | cut -d, -f1,10,13,15 | head -14
