#!/bin/bash
# r02 call 27: final build -- smoke(), full GPU suite, C4 / C1 lines, the bench at the driver's settings
OUT=gpurun_out/r02c27
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.txt
timeout 900 python -m pytest tests -m gpu -q -x --durations=4 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -8 $OUT/pytest_gpu.txt
for wl in C4 C1; do
  timeout 900 python bench.py --workload $wl --steps 4 --warmup 3 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "$wl rc=$?"
  python -c "
import json;d=json.load(open('$OUT/bench_$wl.json'));print('$wl', 'value', d.get('value'), 'e2e', d.get('e2e',{}).get('value'), d.get('ms_per_step'), 'frac', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('error'))"
done
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_C3_driver.json 2> $OUT/bench_C3_driver.err; echo "C3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c27/bench_C3_driver.json'))
print('C3 value', d.get('value'), 'e2e', d.get('e2e',{}).get('value'), d.get('step_ms'), 'frac', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'),
      d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('bench_wall_s'), d.get('config',{}).get('lanes_per_gpu'), d.get('error'))
print(d.get('counters'))
for r in d.get('kernel_breakdown', [])[:10]: print(r)
PY
