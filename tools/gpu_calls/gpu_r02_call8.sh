#!/bin/bash
# r02 call 8: 64 x 96 fused tiles + division-free DMMA epilogue, two-row ormtr: tests, micro-benchmarks, A/B bench lines
OUT=gpurun_out/r02c8
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "heev_rows or rcontract_mpo or workspace" > $OUT/pytest_new.txt 2>&1; echo "rc=$?" >> $OUT/pytest_new.txt; tail -6 $OUT/pytest_new.txt
timeout 300 python tools/bench_heev.py > $OUT/heev.txt 2>&1; grep "^k" $OUT/heev.txt | cut -c1-200
timeout 300 python tools/bench_matvec.py > $OUT/matvec.json 2> $OUT/matvec.err; python -c "
import json;d=json.load(open('$OUT/matvec.json'))
for k,v in d.items(): print(k,{a:round(b,3) for a,b in v.items() if a in ('rcontract_ms','wapply_ms','lcontract_ms','fused_bk8_ms','fused_bk8_frac','matvec_round1_frac','matvec_fused_frac','fused_vs_chain_max_abs_diff')})"; tail -3 $OUT/matvec.err
timeout 900 python -m pytest tests -m gpu -q -x --durations=3 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -6 $OUT/pytest_gpu.txt
for cfg in "default::" "oldheev:YB_HEEV_ROWS=0:" "baseline:YB_HEEV_ROWS=0 YB_FUSED_MPO=0:" "b160::--batch 160"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; extra=${rest#*:}
  env $envs timeout 700 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --roofline-steps 0 $extra > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python -c "
import json;d=json.load(open('$OUT/bench_$name.json'));print('$name', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), d['step_ms'], 'launches', d['gpu_launches'], d['consistency']['status_nonzero'])"
done
