#!/bin/bash
# r02 call 16: Cholesky-QR gauge shifts + unitary-jump rule on the B200 -- kernel test, engine tests (Gram / cholqr routes forced on the
# small fixtures, ten-step C3 / C2 runs against the reference), C3 short bench with the kernel breakdown (A/B: YB_CHOLQR=0)
OUT=gpurun_out/r02c16
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "cholqr" > $OUT/pytest_cholqr.txt 2>&1; echo "rc=$?" >> $OUT/pytest_cholqr.txt; tail -15 $OUT/pytest_cholqr.txt
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x --durations=5 > $OUT/pytest_engine.txt 2>&1; echo "rc=$?" >> $OUT/pytest_engine.txt; tail -12 $OUT/pytest_engine.txt
timeout 700 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --roofline-steps 1 > $OUT/bench_C3.json 2> $OUT/bench_C3.err; echo "C3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c16/bench_C3.json'))
print('C3 value', d.get('value'), d.get('step_ms'), 'frac', d.get('roofline',{}).get('frac'), d.get('consistency',{}).get('status_nonzero'), d.get('error'))
print(d.get('counters'))
for r in d.get('kernel_breakdown', [])[:24]: print(r)
PY
tail -3 $OUT/bench_C3.err
