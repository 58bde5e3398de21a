#!/bin/bash
# r02 call 3: Jacobi noise-floor fix -- regression tests, the full -m gpu suite, the flagged C3 trajectories over 20 steps
OUT=gpurun_out/r02c3
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "jacobi or svd" > $OUT/pytest_jacobi.txt 2>&1; echo "rc=$?" >> $OUT/pytest_jacobi.txt; tail -15 $OUT/pytest_jacobi.txt
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -15 $OUT/pytest_gpu.txt
DBG_TRAJ=103,113,136,0,1,2,3,4 DBG_STEPS=20 timeout 900 python tools/debug_status.py > $OUT/debug_status.txt 2>&1; echo "debug rc=$?" >> $OUT/debug_status.txt; tail -6 $OUT/debug_status.txt
