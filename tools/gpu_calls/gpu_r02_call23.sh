#!/bin/bash
# r02 call 23: CUDA graphs for the gauge shifts -- engine tests (graph tests, ten-step runs on a side stream), C4 / C1 / C3 lines
OUT=gpurun_out/r02c23
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x --durations=4 > $OUT/pytest_engine.txt 2>&1; echo "rc=$?" >> $OUT/pytest_engine.txt; tail -12 $OUT/pytest_engine.txt
timeout 900 python bench.py --workload C4 --steps 4 --warmup 3 --no-cpu-baseline > $OUT/bench_C4.json 2> $OUT/bench_C4.err; echo "C4 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_C4.json'));print('C4', 'value', d.get('value'), d.get('ms_per_step'), d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('counters'), d.get('error'))"; tail -3 $OUT/bench_C4.err
timeout 600 python bench.py --workload C1 --steps 4 --warmup 3 --no-cpu-baseline > $OUT/bench_C1.json 2> $OUT/bench_C1.err; echo "C1 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c23/bench_C1.json'))
print('C1 value', d.get('value'), d.get('step_ms'), d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('counters'), d.get('error'))
for r in d.get('kernel_breakdown', [])[:12]: print(r)
PY
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --roofline-steps 0 > $OUT/bench_C3.json 2> $OUT/bench_C3.err; echo "C3 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_C3.json'));print('C3 value', d.get('value'), d.get('step_ms'), d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('counters'), d.get('error'))"
