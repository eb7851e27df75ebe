#!/bin/bash
OUT=gpurun_out/${1:-probe}; mkdir -p $OUT
P=tools/probes/bin/stream_probe
run() { timeout 120 $P "$@" 2>&1 | tail -1 | tee -a $OUT/probe.txt; }
run copy4
for m in blob blobldg blobst; do
  run $m 16384 16384 128 8
  run $m 16384 16384 128 4
  run $m 16384 16384 128 2
  run $m 16384 16384 256 4
done
run tile 16384 16384 128 4 3 3
run tile 16384 16384 128 2 4 4
run tile 16384 16384 128 2 3 6
run tile 16384 16384 128 4 6 2
run rowchunk 16384 16384 128 4 3 3
run rowchunk 16384 16384 256 2 4 2
run rowchunk 16384 16384 128 2 4 4
