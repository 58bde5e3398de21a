#!/bin/bash
# r02 call 26: small-product kernel for slots up to 32 x 32 (default) against 64 x 64 (YB_GEMM_SMALL_MAX=64, the earlier rule) on C2 / C4; kernel tests
OUT=gpurun_out/r02c26
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or small" > $OUT/pytest_kernels.txt 2>&1; echo "rc=$?" >> $OUT/pytest_kernels.txt; tail -3 $OUT/pytest_kernels.txt
for sm in 32 64; do
for wl in C2 C4; do
  YB_GEMM_SMALL_MAX=$sm timeout 900 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > $OUT/bench_${wl}_sm$sm.json 2> $OUT/bench_${wl}_sm$sm.err; echo "$wl small_max=$sm rc=$?"
  python -c "
import json;d=json.load(open('$OUT/bench_${wl}_sm$sm.json'));print('$wl sm$sm', 'value', d.get('value'), d.get('ms_per_step'), 'frac', d.get('roofline',{}).get('frac'), d.get('consistency',{}).get('status_nonzero'), d.get('error'))
for r in d.get('kernel_breakdown', [])[:5]: print(r)"
done
done
