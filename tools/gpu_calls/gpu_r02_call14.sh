#!/bin/bash
# r02 call 14: thread-per-element kernel for small products (chi <= 16 chains, chain ends): suite, bench lines C1 / C5 / C3 (short)
OUT=gpurun_out/r02c14
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x --durations=3 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -6 $OUT/pytest_gpu.txt
timeout 600 python bench.py --workload C1 --steps 4 --warmup 3 > $OUT/bench_C1.json 2> $OUT/bench_C1.err; echo "C1 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_C1.json'));print('C1', 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['step_ms'], {k:(round(v,5) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k in ('bound','achieved','peak','frac')}, 'cpu', d.get('cpu_baseline',{}).get('value'), d['consistency']['status_nonzero'], d['gpu_launches'])"; tail -2 $OUT/bench_C1.err
timeout 900 python bench.py --workload C5 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_C5.json 2> $OUT/bench_C5.err; echo "C5 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_C5.json'));print('C5', 'value', d.get('value'), d.get('step_ms'), 'frac', d.get('roofline',{}).get('frac'), d.get('consistency'), d.get('gpu_launches'), d.get('error'))"; tail -2 $OUT/bench_C5.err
timeout 700 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --roofline-steps 1 > $OUT/bench_C3.json 2> $OUT/bench_C3.err
python -c "
import json;d=json.load(open('$OUT/bench_C3.json'));print('C3', 'value', round(d['value'],3), d['step_ms'], 'frac', round(d['roofline']['frac'],3), 'gemm ms', round(d['roofline']['gemm_ms_per_step']), d['roofline']['other_kernels_ms_per_step'], d['consistency']['status_nonzero'])"
