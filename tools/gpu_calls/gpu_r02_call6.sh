#!/bin/bash
# r02 call 6: yb_heev_rows (blocked hetrd + ormtr) and the DMMA epilogue of yb_rcontract_mpo: kernel tests, micro-benchmarks, suite, bench
OUT=gpurun_out/r02c6
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "heev_rows or rcontract_mpo" > $OUT/pytest_new.txt 2>&1; echo "rc=$?" >> $OUT/pytest_new.txt; tail -25 $OUT/pytest_new.txt
timeout 300 python tools/bench_heev.py > $OUT/heev.txt 2>&1; tail -12 $OUT/heev.txt | cut -c1-300
timeout 300 python tools/bench_matvec.py > $OUT/matvec.json 2> $OUT/matvec.err; python -c "
import json;d=json.load(open('$OUT/matvec.json'))
for k,v in d.items(): print(k,{a:round(b,3) for a,b in v.items()})"; tail -3 $OUT/matvec.err
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -12 $OUT/pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "bench rc=$?"; tail -c 1200 $OUT/bench_c3.json; tail -5 $OUT/bench_c3.err
YB_WAPPLY_ONEPASS=1 YB_FUSED_MPO=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c3_onepass.json 2> $OUT/bench_c3_onepass.err; echo "onepass rc=$?"; tail -c 600 $OUT/bench_c3_onepass.json
