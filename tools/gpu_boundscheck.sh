#!/bin/bash
# Parity suite + sanitize cases on a build whose sweep kernel bounds-checks every lattice access (our substitute for
# compute-sanitizer memcheck, which is closed on the GPU pool).  Build first:
#   nvcc ... -DLBM_BOUNDS_CHECK -shared -o lbm-graphcore_b200/exp/lib_boundscheck.so lbm-graphcore_b200/csrc/lbm_capi.cu
OUT=gpurun_out/${1:-boundscheck}; mkdir -p $OUT
export LBM_B200_LIB=$PWD/lbm-graphcore_b200/exp/lib_boundscheck.so
timeout 600 python tools/sanitize_cases.py 2>&1 | tail -8 | tee $OUT/sanitize_cases.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_boundscheck.txt
