#!/bin/bash
# r02 call 19: full GPU suite on the fused-launch build, ncu evidence of the new kernels (launch list + --set full with source), the bench at the
# driver's settings (--steps 20 --warmup 5)
OUT=gpurun_out/r02c19
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -9 $OUT/pytest_gpu.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_target.csv python tools/ncu_target.py > $OUT/launches_target.log 2>&1
grep -E "k_cholqr|k_heev_order|k_ormtr|k_stedc|k_hetrd" $OUT/launches_target.csv | awk -F'","' '{print $5, $NF}' | cut -c1-160 | tail -16
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cholqr_factor|k_ormtr_rows|k_heev_order" -s 2 -c 5 -o $OUT/prof_new python tools/ncu_target.py > $OUT/ncu_full.log 2>&1; tail -3 $OUT/ncu_full.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_C3_driver.json 2> $OUT/bench_C3_driver.err; echo "C3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c19/bench_C3_driver.json'))
print('C3 value', d.get('value'), 'e2e', d.get('e2e',{}).get('value'), d.get('step_ms'), 'frac', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'),
      d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('bench_wall_s'), d.get('config',{}).get('lanes_per_gpu'), d.get('error'))
print(d.get('counters'))
for r in d.get('kernel_breakdown', [])[:14]: print(r)
PY
tail -3 $OUT/bench_C3_driver.err
ls -la $OUT
