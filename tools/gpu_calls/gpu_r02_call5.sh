#!/bin/bash
# r02 call 5: fused RCONTRACT + MPO epilogue: kernel test, micro-benchmark (8- vs 16-deep k tiles), engine fixtures with it, C3 bench lines
OUT=gpurun_out/r02c5
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "rcontract_mpo" > $OUT/pytest_rw.txt 2>&1; echo "rc=$?" >> $OUT/pytest_rw.txt; tail -8 $OUT/pytest_rw.txt
timeout 300 python tools/bench_matvec.py > $OUT/matvec.json 2> $OUT/matvec.err; tail -c 3000 $OUT/matvec.json; tail -3 $OUT/matvec.err
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -6 $OUT/pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c3_fused.json 2> $OUT/bench_c3_fused.err; echo "fused rc=$?"; tail -c 1500 $OUT/bench_c3_fused.json; tail -5 $OUT/bench_c3_fused.err
YB_GEMM_BK16=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c3_fused_bk16.json 2> $OUT/bench_c3_fused_bk16.err; echo "bk16 rc=$?"; tail -c 1500 $OUT/bench_c3_fused_bk16.json
