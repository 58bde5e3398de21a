#!/bin/bash
# First GPU call of the next round (1 GPU): everything that was written after round 1's GPU budget ran out, in priority order.
#   1 full -m gpu suite (first execution of: measure_shots, scheduled jumps, SWAP ladders, backend boundary, config-4 fixture)
#   2 the opt-in config-5 stress fixture (Heisenberg L=200, chi=256; k = 512 eigen-problems)   3 default bench line
#   4 ncu launch list of one whole lane-step at batch 64 (no launch-count cut)   5 ncu --set full of k_gemm on ragged (chi ~ 113) tiles
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q --durations=10 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -15 $OUT/pytest_gpu.txt
YB_RUN_C5=1 timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "stress" --durations=3 > $OUT/pytest_c5.txt 2>&1; echo "c5 rc=$?" >> $OUT/pytest_c5.txt; tail -8 $OUT/pytest_c5.txt
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err; tail -c 2500 $OUT/bench.json
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 --csv --log-file $OUT/launches.csv \
    python bench.py --batch 64 --lanes 1 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-roofline > $OUT/launches_bench.log 2>&1
python tools/summarise_launches.py $OUT/launches.csv > $OUT/launches_summary.json 2>> $OUT/launches_bench.log; gzip -f $OUT/launches.csv
head -c 1500 $OUT/launches_summary.json
