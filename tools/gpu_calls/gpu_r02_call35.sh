#!/bin/bash
# r02 call 35 (the round's last GPU seconds): ncu launch list of the bench command itself, a window of one lane of 64 trajectories inside
# the first TDVP sweep of a C3 step (bulk sites); --kill ends the run after the window, so no bench line is printed (a number under
# the profiler is never a bench value)
OUT=gpurun_out/r02c35
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 62 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 --launch-count 4000 --kill 1 --csv \
  --log-file $OUT/launches_bench_c3_window.csv \
  python bench.py --batch 64 --lanes 1 --steps 1 --warmup 0 --no-cpu-baseline --roofline-steps 0 > $OUT/bench.log 2>&1
echo "rc=$?"; wc -l $OUT/launches_bench_c3_window.csv
