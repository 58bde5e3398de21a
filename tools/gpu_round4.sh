#!/bin/bash
# One gpurun call, ordered by priority (the call may be cut by the GPU-minute budget at any point).
#   1 parity tests (divide & conquer eigen-solver on by default; falls back to the implicit-QL kernel for the rest if they fail)
#   2 eigen-solver micro-benchmark   3 default bench line   4 ncu launch list of one step   5 CPU reference arm
#   6 ncu --set full of the eigen-solver kernels   7 compute-sanitizer racecheck of the eigen-solver
TAG=${1:-r01d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt

timeout 420 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1
rc=$?
echo "pytest rc=$rc" >> $OUT/pytest_gpu.txt
tail -4 $OUT/pytest_gpu.txt
if [[ $rc != 0 ]]; then
  export YB_HEEV_QL=1
  timeout 420 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_ql.txt 2>&1
  echo "pytest (YB_HEEV_QL=1) rc=$?" >> $OUT/pytest_gpu_ql.txt
  tail -4 $OUT/pytest_gpu_ql.txt
  timeout 300 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_all_failures.txt 2>&1
fi
echo "YB_HEEV_QL=${YB_HEEV_QL:-unset}" >> $OUT/gpu.txt

timeout 200 python tools/bench_heev.py > $OUT/heev_microbench.json 2> $OUT/heev_microbench.err
cat $OUT/heev_microbench.json

timeout 720 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/bench.err
tail -c 3500 $OUT/bench.json
tail -3 $OUT/bench.err

timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 --launch-count 24000 --csv \
    --log-file $OUT/launches.csv python bench.py --batch 64 --lanes 1 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $OUT/launches_bench.log 2>&1
python tools/summarise_launches.py $OUT/launches.csv > $OUT/launches_summary.json 2>> $OUT/launches_bench.log
gzip -f $OUT/launches.csv
head -c 1800 $OUT/launches_summary.json

timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
tail -c 1200 $OUT/bench_reference.json

timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_hetrd|k_ungtr|k_stedc|k_steqr' --launch-skip 6 --launch-count 6 -f -o $OUT/heev_full \
    python tools/bench_heev.py > $OUT/ncu_heev.log 2>&1
ncu -i $OUT/heev_full.ncu-rep --page raw --csv > $OUT/heev_full_raw.csv 2>/dev/null
python tools/ncu_pick.py $OUT/heev_full_raw.csv > $OUT/heev_full_summary.json 2>> $OUT/ncu_heev.log

timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/racecheck_heev.py > $OUT/racecheck_heev.txt 2>&1
tail -5 $OUT/racecheck_heev.txt
ls -la $OUT
