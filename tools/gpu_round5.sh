#!/bin/bash
# Second GPU call of the session: new tests (Philox kernel, ensemble 3-sigma, row-ring bit identity), eigen-solver A/B
# (cp.async row ring on/off), short bench A/B over lanes.  Ordered by priority; bounded by the remaining GPU minutes.
TAG=${1:-r01e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1
rc=$?
echo "pytest rc=$rc" >> $OUT/pytest_gpu.txt
tail -4 $OUT/pytest_gpu.txt
if [[ $rc != 0 ]]; then
  YB_HEEV_NORING=1 timeout 400 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_noring.txt 2>&1
  echo "pytest (YB_HEEV_NORING=1, all failures listed) rc=$?" >> $OUT/pytest_gpu_noring.txt
  tail -15 $OUT/pytest_gpu_noring.txt
fi
timeout 200 python tools/bench_heev.py > $OUT/heev_microbench.json 2> $OUT/heev_microbench.err
cat $OUT/heev_microbench.json; tail -2 $OUT/heev_microbench.err
A="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-roofline"
timeout 300 python bench.py $A > $OUT/bench_ring_l2.json 2> $OUT/bench_ring_l2.err; python tools/bench_brief.py $OUT/bench_ring_l2.json
timeout 300 python bench.py $A --lanes 4 > $OUT/bench_ring_l4.json 2> $OUT/bench_ring_l4.err; python tools/bench_brief.py $OUT/bench_ring_l4.json
YB_HEEV_NORING=1 timeout 300 python bench.py $A > $OUT/bench_noring_l2.json 2> $OUT/bench_noring_l2.err; python tools/bench_brief.py $OUT/bench_noring_l2.json
timeout 300 python bench.py $A --lanes 3 --batch 288 > $OUT/bench_ring_l3_b288.json 2> $OUT/bench_ring_l3_b288.err; python tools/bench_brief.py $OUT/bench_ring_l3_b288.json
ls -la $OUT
