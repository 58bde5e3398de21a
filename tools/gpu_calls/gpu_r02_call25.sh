#!/bin/bash
# r02 call 25: grid target 64 CTAs per SM (config 2 regression of call 24): kernel tests, C2 / C4 / C1 lines
OUT=gpurun_out/r02c25
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or small or svd" > $OUT/pytest_kernels.txt 2>&1; echo "rc=$?" >> $OUT/pytest_kernels.txt; tail -3 $OUT/pytest_kernels.txt
for wl in C2 C4 C1; do
  timeout 900 python bench.py --workload $wl --steps 4 --warmup 3 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "$wl rc=$?"
  python -c "
import json;d=json.load(open('$OUT/bench_$wl.json'));print('$wl', 'value', d.get('value'), 'e2e', d.get('e2e',{}).get('value'), d.get('ms_per_step'), 'frac', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('error'))"
done
