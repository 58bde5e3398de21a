#!/bin/bash
# One gpurun call: GPU parity tests, default bench, ncu launch list and ncu --set full captures.  Outputs -> gpurun_out/<tag>/
# usage: tools/gpu_round.sh <tag> [stages]     stages: any of  t(ests) b(ench) r(eference arm) l(aunch list) f(ull ncu) m(icrobench)
TAG=${1:-run}
STAGES=${2:-tblfm}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
if [[ $STAGES == *t* ]]; then
  timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1
  echo "pytest rc=$?" >> $OUT/pytest_gpu.txt
  tail -3 $OUT/pytest_gpu.txt
fi
if [[ $STAGES == *m* ]]; then
  timeout 300 python tools/bench_heev.py > $OUT/heev_microbench.json 2> $OUT/heev_microbench.err
  timeout 300 python tools/bench_matvec.py > $OUT/matvec_microbench.json 2> $OUT/matvec_microbench.err
  tail -2 $OUT/heev_microbench.json
fi
if [[ $STAGES == *b* ]]; then
  timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
  echo "bench rc=$?" >> $OUT/bench.err
  tail -c 3000 $OUT/bench.json
fi
if [[ $STAGES == *r* ]]; then
  timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
  tail -c 1500 $OUT/bench_reference.json
fi
if [[ $STAGES == *l* ]]; then
  # launch list of one sampled order-2 step (same workload, one lane, batch 64): skip the initialisation, list 24000 launches
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 --launch-count 24000 --csv \
      --log-file $OUT/launches.csv python bench.py --batch 64 --lanes 1 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $OUT/launches_bench.log 2>&1
  python tools/summarise_launches.py $OUT/launches.csv > $OUT/launches_summary.json 2>> $OUT/launches_bench.log
  gzip -f $OUT/launches.csv
  cat $OUT/launches_summary.json | head -c 2500
fi
if [[ $STAGES == *f* ]]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm --launch-skip 6 --launch-count 3 -f -o $OUT/gemm_full \
      python tools/bench_matvec.py > $OUT/ncu_gemm.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_hetrd|k_ungtr|k_steqr|k_wapply' --launch-skip 3 --launch-count 4 -f -o $OUT/heev_full \
      python tools/bench_heev.py > $OUT/ncu_heev.log 2>&1
  ls -la $OUT
fi
