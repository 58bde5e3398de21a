#!/bin/bash
# r02 call 21: observables / diagnostics kernels, k_cholqr_factor with the triangle-following thread split: full GPU suite, compute-sanitizer
# (racecheck + memcheck) on the kernels of this session, launch times of the factorisation
OUT=gpurun_out/r02c21
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x --durations=4 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -8 $OUT/pytest_gpu.txt
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python tools/racecheck_new.py > $OUT/racecheck_new.txt 2>&1; tail -6 $OUT/racecheck_new.txt
timeout 600 compute-sanitizer --tool memcheck python tools/racecheck_new.py > $OUT/memcheck_new.txt 2>&1; tail -4 $OUT/memcheck_new.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_target.csv python tools/ncu_target.py > $OUT/launches_target.log 2>&1
grep -E "k_cholqr" $OUT/launches_target.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | tail -3
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --roofline-steps 1 > $OUT/bench_C3.json 2> $OUT/bench_C3.err; echo "C3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c21/bench_C3.json'))
print('C3 value', d.get('value'), d.get('step_ms'), d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('error'))
for r in d.get('kernel_breakdown', []):
    if r['family'] in ('cholqr','local'): print(r)
PY
