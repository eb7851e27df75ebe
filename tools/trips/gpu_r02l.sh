#!/bin/bash
# NVLink byte counters of ncu, asked for by name, on one fused compute+halo launch per GPU of a 2-GPU single-process run
set -u
OUT=gpurun_out/r02l; mkdir -p $OUT
B="python bench.py --in-process 2 --steps 3 --warmup 3"
M=nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_packet_request_data_user.sum,nvlrx__bytes_packet_request_data_user.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:"sweep_kernel|init_halo" -c 14 --csv --log-file $OUT/nvlink_counters.csv $B > $OUT/nvlink_counters.log 2>&1; echo "rc=$?"
grep -v "^==" $OUT/nvlink_counters.csv | cut -d, -f1,5,10,13,14,15 | head -80
