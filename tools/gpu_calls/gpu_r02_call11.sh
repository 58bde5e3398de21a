#!/bin/bash
# r02 call 11: DMMA issue order (dependent pairs a tile sweep apart): bit-identity tests, micro-benchmark, bench line
OUT=gpurun_out/r02c11
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or rcontract" > $OUT/pytest_gemm.txt 2>&1; echo "rc=$?" >> $OUT/pytest_gemm.txt; tail -3 $OUT/pytest_gemm.txt
timeout 300 python tools/bench_matvec.py > $OUT/matvec.json 2> $OUT/matvec.err; python -c "
import json;d=json.load(open('$OUT/matvec.json'))
for k,v in d.items(): print(k,{a:round(b,3) for a,b in v.items() if a in ('rcontract_ms','lcontract_ms','fused_bk8_ms','rcontract_frac','lcontract_frac','fused_bk8_frac','rcontract_warps16_ms','fused_warps16_ms','matvec_fused_frac','matvec_fused_warps16_frac','cublas_rcontract_frac')})"; tail -3 $OUT/matvec.err
timeout 700 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --roofline-steps 1 > $OUT/bench_default.json 2> $OUT/bench_default.err
python -c "
import json;d=json.load(open('$OUT/bench_default.json'));print('default', 'value', round(d['value'],3), d['step_ms'], 'frac', round(d['roofline']['frac'],3), 'gemm ms', round(d['roofline']['gemm_ms_per_step']), d['consistency']['status_nonzero'])"
