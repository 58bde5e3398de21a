#!/bin/bash
# r02 call 24: batch-sized grids + tiny-product path of the tile GEMM: kernel tests, engine tests, C4 / C1 / C2 / C3 lines
OUT=gpurun_out/r02c24
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > $OUT/pytest_kernels.txt 2>&1; echo "rc=$?" >> $OUT/pytest_kernels.txt; tail -5 $OUT/pytest_kernels.txt
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_kernels_gpu.py --durations=3 > $OUT/pytest_rest.txt 2>&1; echo "rc=$?" >> $OUT/pytest_rest.txt; tail -8 $OUT/pytest_rest.txt
for wl in C4 C1 C2; do
  timeout 900 python bench.py --workload $wl --steps 4 --warmup 3 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "$wl rc=$?"
  python -c "
import json;d=json.load(open('$OUT/bench_$wl.json'));print('$wl', 'value', d.get('value'), 'e2e', d.get('e2e',{}).get('value'), d.get('ms_per_step'), 'frac', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('error'))"
done
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --roofline-steps 0 > $OUT/bench_C3.json 2> $OUT/bench_C3.err; echo "C3 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_C3.json'));print('C3 value', d.get('value'), d.get('step_ms'), d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('error'))"
