#!/bin/bash
# Last GPU call of the round: full bench.py code path on a tiny chain (all sections), GPU tests at the final code state,
# row-ring depth sweep of the eigen-solver, then the default bench line for profiles/ (cut by the remaining budget if need be).
TAG=${1:-r01f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 200 python __graft_entry__.py smoke > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.txt
timeout 200 python bench.py --length 12 --chi 8 --batch 8 --steps 2 --warmup 3 > $OUT/bench_tiny.json 2> $OUT/bench_tiny.err; echo "tiny bench rc=$?"; tail -c 600 $OUT/bench_tiny.json; tail -3 $OUT/bench_tiny.err
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -4 $OUT/pytest_gpu.txt
timeout 200 python tools/bench_heev.py > $OUT/heev_microbench.json 2> $OUT/heev_microbench.err; cat $OUT/heev_microbench.json; tail -2 $OUT/heev_microbench.err
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err; tail -c 2500 $OUT/bench.json; tail -2 $OUT/bench.err
