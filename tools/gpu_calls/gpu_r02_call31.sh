#!/bin/bash
# r02 call 31: graph handles count their kernels in yb_launch_count: graph tests + short C3 line
OUT=gpurun_out/r02c31
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "graphed or gram_svd_route" > $OUT/pytest_graphs.txt 2>&1; echo "rc=$?" >> $OUT/pytest_graphs.txt; tail -3 $OUT/pytest_graphs.txt
timeout 400 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --roofline-steps 0 > $OUT/bench_C3.json 2> $OUT/bench_C3.err; echo "C3 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_C3.json'));print('C3 value', d.get('value'), d.get('step_ms'), d.get('consistency',{}).get('status_nonzero'), d.get('gpu_launches'), d.get('counters',{}).get('graph_launches'), d.get('error'))"
