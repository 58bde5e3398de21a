#!/bin/bash
# r02 call 20: jump test / selection kernels (incl. in-kernel Philox) on the B200: kernel test, engine + ensemble + simulator GPU tests, C3 short bench
OUT=gpurun_out/r02c20
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "jump or small_util or philox" > $OUT/pytest_kernels.txt 2>&1; echo "rc=$?" >> $OUT/pytest_kernels.txt; tail -6 $OUT/pytest_kernels.txt
timeout 900 python -m pytest tests -m gpu -q -x --durations=4 --deselect tests/test_kernels_gpu.py > $OUT/pytest_rest.txt 2>&1; echo "rc=$?" >> $OUT/pytest_rest.txt; tail -9 $OUT/pytest_rest.txt
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --roofline-steps 0 > $OUT/bench_C3.json 2> $OUT/bench_C3.err; echo "C3 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_C3.json'));print('C3 value', d.get('value'), d.get('step_ms'), d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('error'))"
tail -3 $OUT/bench_C3.err
