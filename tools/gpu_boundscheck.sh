#!/bin/bash
# Parity suites + sanitize cases on a build whose sweep kernel bounds-checks every lattice access, ghost-column and corner
# pushes of the block partition included (our substitute for compute-sanitizer memcheck, which is closed on the GPU pool).
# Build first:  make -C lbm-graphcore_b200/csrc ../liblbm_b200_boundscheck.so
OUT=gpurun_out/${1:-boundscheck}; mkdir -p $OUT
export LBM_B200_LIB=$PWD/lbm-graphcore_b200/liblbm_b200_boundscheck.so
timeout 600 python tools/sanitize_cases.py 2>&1 | tail -8 | tee $OUT/sanitize_cases.txt
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_blocks_gpu.py tests/test_solve_host_gpu.py tests/test_prepared_gpu.py tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_boundscheck.txt
