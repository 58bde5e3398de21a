#!/bin/bash
# r02 call 17: capped splits with top rows only (yb_heev_rows_top) -- kernel tests, eigen-solver micro-benchmark, engine suite, C3 short bench
OUT=gpurun_out/r02c17
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "cholqr or heev or svd" > $OUT/pytest_kernels.txt 2>&1; echo "rc=$?" >> $OUT/pytest_kernels.txt; tail -12 $OUT/pytest_kernels.txt
timeout 300 python tools/bench_heev.py > $OUT/heev.txt 2>&1; grep "^k" $OUT/heev.txt | cut -c1-250
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x --durations=3 > $OUT/pytest_engine.txt 2>&1; echo "rc=$?" >> $OUT/pytest_engine.txt; tail -8 $OUT/pytest_engine.txt
timeout 700 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --roofline-steps 1 > $OUT/bench_C3.json 2> $OUT/bench_C3.err; echo "C3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c17/bench_C3.json'))
print('C3 value', d.get('value'), d.get('step_ms'), 'frac', d.get('roofline',{}).get('frac'), d.get('consistency',{}).get('status_nonzero'), d.get('error'))
print(d.get('counters'))
for r in d.get('kernel_breakdown', [])[:16]: print(r)
PY
tail -3 $OUT/bench_C3.err
