#!/bin/bash
# r02 call 9: long-run parity tests (C3 / C2 ten steps, C4 at its own size), second-level cut-off A/B, first bench line at the driver's settings
OUT=gpurun_out/r02c9
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "long_run or c4_full or gram_svd or baseline_sizes" --durations=6 > $OUT/pytest_long.txt 2>&1; echo "rc=$?" >> $OUT/pytest_long.txt; tail -12 $OUT/pytest_long.txt
timeout 900 python -m pytest tests -m gpu -q --durations=3 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -5 $OUT/pytest_gpu.txt
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --roofline-steps 0 --batch 128 > $OUT/bench_b128_tau1e-8.json 2> $OUT/bench_b128_tau.err
python -c "
import json;d=json.load(open('$OUT/bench_b128_tau1e-8.json'));print('tau1e-8 b128', 'value', round(d['value'],3), d['step_ms'], 'launches', d['gpu_launches'], d['consistency']['status_nonzero'])"
timeout 1200 python bench.py --steps 20 --warmup 5 > $OUT/bench_driver.json 2> $OUT/bench_driver.err; echo "driver-settings bench rc=$?"; tail -c 3000 $OUT/bench_driver.json; tail -3 $OUT/bench_driver.err
