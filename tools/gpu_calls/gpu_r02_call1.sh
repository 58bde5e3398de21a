#!/bin/bash
# r02 call 1: reproduce the C3 status failure on the trajectories the driver flagged; first hardware run of config 5; lane A/B at batch 128
OUT=gpurun_out/r02c1
mkdir -p $OUT
export PYTHONUNBUFFERED=1
DBG_TRAJ=103,113,136,0,1,2,3,4 DBG_STEPS=20 timeout 900 python tools/debug_status.py > $OUT/debug_status.txt 2>&1; echo "debug rc=$?" >> $OUT/debug_status.txt; tail -20 $OUT/debug_status.txt
mv gpurun_out/dbg_w_*.npy $OUT/ 2>/dev/null
YB_RUN_C5=1 timeout 700 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "stress" --durations=3 > $OUT/pytest_c5.txt 2>&1; echo "c5 rc=$?" >> $OUT/pytest_c5.txt; tail -8 $OUT/pytest_c5.txt
timeout 400 python bench.py --batch 128 --lanes 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/bench_b128_l1.json 2> $OUT/bench_b128_l1.err; tail -c 1800 $OUT/bench_b128_l1.json
timeout 400 python bench.py --batch 128 --lanes 2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-roofline > $OUT/bench_b128_l2.json 2> $OUT/bench_b128_l2.err; tail -c 600 $OUT/bench_b128_l2.json
