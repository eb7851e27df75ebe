#!/bin/bash
# round 2, trip J (1 GPU): persistent vs per-step + graph on L2-resident grids; ncu of the fused compute+halo launch at the N=4 size
set -u
OUT=gpurun_out/r02j; mkdir -p $OUT
echo "== mid-size grids"; timeout 1200 python tools/mid_grids.py 2>&1 | tee $OUT/mid_grids.txt
B="python bench.py --no-e2e --no-cpu-baseline --no-parity --no-weak --no-shipped --steps 3 --warmup 3 --local-slabs 4"
timeout 600 $B > $OUT/multi4.plain.log 2>&1 && timeout 1200 ncu --set full --clock-control none -k regex:sweep_kernel -s 16 -c 1 -f -o $OUT/sweep_multi4 $B > $OUT/multi4.ncu.log 2>&1
ncu -i $OUT/sweep_multi4.ncu-rep --page raw --csv > $OUT/sweep_multi4.raw.csv 2>/dev/null; ncu -i $OUT/sweep_multi4.ncu-rep --page details > $OUT/sweep_multi4.details.txt 2>/dev/null; rm -f $OUT/sweep_multi4.ncu-rep
grep -E "gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum" $OUT/sweep_multi4.details.txt | head; python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02j/sweep_multi4.raw.csv')))
idx={n:i for i,n in enumerate(rows[0])}
for n in ('Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__grid_size'):
    print(n, rows[2][idx[n]], rows[1][idx[n]])
PY
