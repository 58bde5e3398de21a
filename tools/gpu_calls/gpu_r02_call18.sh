#!/bin/bash
# r02 call 18: top-mode kernel test (1 x 1 fix), C4 bench line (digital workload), lanes A/B on C3 (3 x 64 and 4 x 48 against 2 x 96)
OUT=gpurun_out/r02c18
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "top or cholqr" > $OUT/pytest_kernels.txt 2>&1; echo "rc=$?" >> $OUT/pytest_kernels.txt; tail -5 $OUT/pytest_kernels.txt
timeout 900 python bench.py --workload C4 --steps 4 --warmup 3 > $OUT/bench_C4.json 2> $OUT/bench_C4.err; echo "C4 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_C4.json'));print('C4', 'value', d.get('value'), 'e2e', d.get('e2e',{}).get('value'), d.get('ms_per_step'), d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('consistency'), d.get('gpu_launches'), d.get('error'), d.get('bench_wall_s'))"; tail -3 $OUT/bench_C4.err
for lanes in 3 4; do
  timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --roofline-steps 0 --lanes $lanes > $OUT/bench_C3_l$lanes.json 2> $OUT/bench_C3_l$lanes.err; echo "C3 lanes $lanes rc=$?"
  python -c "
import json;d=json.load(open('$OUT/bench_C3_l$lanes.json'));print('C3 lanes $lanes value', d.get('value'), d.get('step_ms'), d.get('consistency',{}).get('status_nonzero'), d.get('error'))"
done
