#!/bin/bash
# r02 call 13: eigen-solver tweaks (division chains in k_stedc, split dot accumulators in k_ormtr_rows); bench lines of the other BASELINE configs
OUT=gpurun_out/r02c13
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "heev" > $OUT/pytest_heev.txt 2>&1; echo "rc=$?" >> $OUT/pytest_heev.txt; tail -3 $OUT/pytest_heev.txt
timeout 300 python tools/bench_heev.py > $OUT/heev.txt 2>&1; grep "^k" $OUT/heev.txt | cut -c1-120
for wl in C1 C2; do
  timeout 600 python bench.py --workload $wl --steps 4 --warmup 3 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "$wl rc=$?"
  python -c "
import json;d=json.load(open('$OUT/bench_$wl.json'));print('$wl', 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['step_ms'], 'roofline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k in ('bound','achieved','peak','frac')}, 'cpu', d.get('cpu_baseline',{}).get('value'), d['consistency']['status_nonzero'], d['gpu_launches'])"; tail -2 $OUT/bench_$wl.err
done
timeout 900 python bench.py --workload C5 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_C5.json 2> $OUT/bench_C5.err; echo "C5 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_C5.json'));print('C5', 'value', round(d['value'],3), d['step_ms'], 'frac', round(d['roofline']['frac'],3), d['consistency'], d['gpu_launches'])"; tail -2 $OUT/bench_C5.err
