#!/bin/bash
# phases of a whole lbm_runner job on the largest shipped input (where does an end-to-end run spend its time?)
OUT=gpurun_out/${1:-runner}; mkdir -p $OUT
G=tests/golden/params
for mode in "" "--download-cells"; do
  for c in 256x256 1024x1024; do
    echo "== $c $mode" | tee -a $OUT/runner_times.txt
    ( TIMEFORMAT="wall %R s"; time lbm-graphcore_b200/bin/lbm_runner --params $G/input_$c.params --obstacles $G/obstacles_$c.dat --outdir /tmp --reruns 1 -d $mode ) 2>&1 | grep -v "^$" | tee -a $OUT/runner_times.txt
  done
done
