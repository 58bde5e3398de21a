#!/bin/bash
# r02 call 4: host-side restructuring (batched observables / jump plan / ambiguity kernel / Lanczos read schedule / streaming API) on the GPU:
# full suite, the flagged C3 trajectories over 20 steps, the new bench (tiny code path, then C3 short with 2 lanes and with 1 lane)
OUT=gpurun_out/r02c4
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -12 $OUT/pytest_gpu.txt
DBG_TRAJ=103,113,136,0,1,2,3,4 DBG_STEPS=20 timeout 900 python tools/debug_status.py > $OUT/debug_status.txt 2>&1; echo "debug rc=$?" >> $OUT/debug_status.txt; tail -4 $OUT/debug_status.txt
timeout 300 python bench.py --length 12 --chi 8 --batch 8 --steps 2 --warmup 1 --cpu-cores 2 > $OUT/bench_tiny.json 2> $OUT/bench_tiny.err; echo "tiny rc=$?"; tail -c 1500 $OUT/bench_tiny.json; tail -5 $OUT/bench_tiny.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --roofline-every 3 > $OUT/bench_c3_l2.json 2> $OUT/bench_c3_l2.err; echo "c3 l2 rc=$?"; tail -c 2500 $OUT/bench_c3_l2.json; tail -5 $OUT/bench_c3_l2.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --roofline-every 3 --lanes 1 > $OUT/bench_c3_l1.json 2> $OUT/bench_c3_l1.err; echo "c3 l1 rc=$?"; tail -c 1200 $OUT/bench_c3_l1.json; tail -5 $OUT/bench_c3_l1.err
