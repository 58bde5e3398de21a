#!/bin/bash
# r02 call 28: bench instruments ONE lane (the others follow concurrently); grid_batch target 8 CTAs per SM: C3 short with the instrumented step, C4, C2
OUT=gpurun_out/r02c28
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --roofline-steps 1 > $OUT/bench_C3.json 2> $OUT/bench_C3.err; echo "C3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c28/bench_C3.json'))
print('C3 value', d.get('value'), d.get('step_ms'), 'frac', d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('share_of_step'), d.get('tdvp_sweep_ms'), d.get('consistency',{}).get('status_nonzero'), d.get('error'))
for r in d.get('kernel_breakdown', [])[:6]: print(r)
PY
tail -3 $OUT/bench_C3.err
for wl in C4 C2; do
  timeout 900 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "$wl rc=$?"
  python -c "
import json;d=json.load(open('$OUT/bench_$wl.json'));print('$wl', 'value', d.get('value'), 'e2e', d.get('e2e',{}).get('value'), d.get('ms_per_step'), 'frac', d.get('roofline',{}).get('frac'), d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('error'))"
done
