#!/bin/bash
# round 2, trip W2 (1 GPU): ncu --set full of one v3 block launch, as a 1x1 "grid" whose ghost columns wrap onto the block itself.
# (With a real neighbour a v3 launch waits for the neighbour's launch of the SAME step to start, which a serialising profiler
# never lets happen: the 2-GPU attempt ended in the kernel's own 20 s timeout, gpurun_out/r02w/blocks_v3.log.)
set -u
OUT=gpurun_out/r02w; mkdir -p $OUT
B="python bench.py --in-process 1 --grid 1x1 --steps 3 --warmup 3"
timeout 300 $B > $OUT/plain.log 2>&1 || { echo plain failed; tail -3 $OUT/plain.log; exit 1; }
tail -1 $OUT/plain.log | cut -c1-200
timeout 600 ncu --set full --clock-control none -k regex:sweep_kernel -s 4 -c 1 -f -o $OUT/blocks_v3 $B > $OUT/blocks_v3_1x1.log 2>&1; echo rc=$?
ncu -i $OUT/blocks_v3.ncu-rep --page details > $OUT/blocks_v3.details.txt 2>/dev/null
ncu -i $OUT/blocks_v3.ncu-rep --page raw --csv > $OUT/blocks_v3.raw.csv 2>/dev/null
rm -f $OUT/blocks_v3.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02w/blocks_v3.raw.csv'))); idx={n:i for i,n in enumerate(rows[0])}
for n in ('Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','l1tex__t_sectors_pipe_lsu_mem_local_op_st_lookup_hit.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size'):
    print(n, rows[2][idx[n]], rows[1][idx[n]])
PY
