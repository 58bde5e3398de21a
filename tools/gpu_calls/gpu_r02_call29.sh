#!/bin/bash
# r02 call 29: final build -- smoke(), full GPU suite (known-answer tests included), the bench at the driver's settings with the one-lane instrumented step
OUT=gpurun_out/r02c29
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.txt
timeout 900 python -m pytest tests -m gpu -q -x --durations=4 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -8 $OUT/pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_C3_driver.json 2> $OUT/bench_C3_driver.err; echo "C3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c29/bench_C3_driver.json'))
print('C3 value', d.get('value'), 'e2e', d.get('e2e',{}).get('value'), d.get('step_ms'), 'frac', d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('share_of_step'), 'cpu', d.get('cpu_baseline',{}).get('value'),
      d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('bench_wall_s'), d.get('config',{}).get('lanes_per_gpu'), d.get('tdvp_sweep_ms'), d.get('whole_step',{}).get('frac_of_fp64_peak_counters'), d.get('error'))
print(d.get('roofline',{}).get('other_kernels_ms_per_step'), d.get('roofline',{}).get('gemm_ms_per_step'))
PY
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; cut -c1-200 $OUT/bench_reference.json
