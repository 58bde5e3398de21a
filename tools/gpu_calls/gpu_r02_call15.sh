#!/bin/bash
# r02 call 15 (second session): state at HEAD re-verified after the container was re-created -- full GPU suite, config-5 parity test,
# C3 short bench with the per-(family, size, phase) kernel breakdown, C5 bench line
OUT=gpurun_out/r02c15
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -9 $OUT/pytest_gpu.txt
YB_RUN_C5=1 timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "stress" --durations=3 > $OUT/pytest_c5.txt 2>&1; echo "c5 rc=$?" >> $OUT/pytest_c5.txt; tail -5 $OUT/pytest_c5.txt
timeout 700 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --roofline-steps 1 > $OUT/bench_C3.json 2> $OUT/bench_C3.err; echo "C3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c15/bench_C3.json'))
print('C3 value', d.get('value'), d.get('step_ms'), 'frac', d.get('roofline',{}).get('frac'), d.get('consistency',{}).get('status_nonzero'), d.get('error'))
for r in d.get('kernel_breakdown', [])[:40]: print(r)
PY
timeout 900 python bench.py --workload C5 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_C5.json 2> $OUT/bench_C5.err; echo "C5 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_C5.json'));print('C5', 'value', d.get('value'), d.get('step_ms'), 'frac', d.get('roofline',{}).get('frac'), d.get('consistency'), d.get('gpu_launches'), d.get('error'))"; tail -2 $OUT/bench_C5.err
