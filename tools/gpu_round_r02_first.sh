#!/bin/bash
# First GPU call of the next round (1 GPU): everything that was written after round 1's GPU budget ran out, in priority order.
#   1 full -m gpu suite.  First execution on a B200 of: measure_shots, scheduled jumps, SWAP ladders, backend boundary, config-4
#     fixture, SimulationProgram (4 fixtures), the gate-MPO route + yb_mpo_site_apply, tdvp_mode="dynamic" (uncapped fixture + exactness
#     at full bond dimension), the "gemm_ragged" A/B option (bit-identity
#     against the default tiling + two engine fixtures).  tests/conftest.py collects these LAST (FIRST_RUN_ON_GPU): prune the list
#     from the log of this call.
#   2 the opt-in config-5 stress fixture (Heisenberg L=200, chi=256; k = 512 eigen-problems)
#   3 default bench line, then the A/B line with YB_GEMM_RAGGED=1 (interleaved 8x8 sub-tiles that skip padding: expected -15..20 %
#     on k_gemm at chi ~ 113 of 128, i.e. -7..9 % on the step; make it the default if the A/B agrees and the suite is green with it),
#     YB_WAPPLY_ONEPASS=1 (k_wapply_onepass: every input element read once; k_wapply is 7 % of the step), and both together
#   4 ncu launch list of one whole lane-step at batch 64   5 ncu --set full of k_gemm on ragged (chi ~ 113) tiles, both variants
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --durations=10 > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -15 $OUT/pytest_gpu.txt
YB_RUN_C5=1 timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "stress" --durations=3 > $OUT/pytest_c5.txt 2>&1; echo "c5 rc=$?" >> $OUT/pytest_c5.txt; tail -8 $OUT/pytest_c5.txt
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err; tail -c 2500 $OUT/bench.json
YB_GEMM_RAGGED=1 timeout 600 python bench.py --no-e2e --no-cpu-baseline > $OUT/bench_gemm_ragged.json 2> $OUT/bench_gemm_ragged.err; tail -c 1500 $OUT/bench_gemm_ragged.json
YB_WAPPLY_ONEPASS=1 timeout 600 python bench.py --no-e2e --no-cpu-baseline > $OUT/bench_wapply_onepass.json 2> $OUT/bench_wapply_onepass.err; tail -c 1500 $OUT/bench_wapply_onepass.json
YB_GEMM_RAGGED=1 YB_WAPPLY_ONEPASS=1 timeout 600 python bench.py --no-e2e --no-cpu-baseline > $OUT/bench_both_options.json 2> $OUT/bench_both_options.err; tail -c 1500 $OUT/bench_both_options.json
timeout 600 python bench.py --no-e2e --no-cpu-baseline > $OUT/bench_gemm_default_ab.json 2> $OUT/bench_gemm_default_ab.err; tail -c 1500 $OUT/bench_gemm_default_ab.json
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 --csv --log-file $OUT/launches.csv \
    python bench.py --batch 64 --lanes 1 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-roofline > $OUT/launches_bench.log 2>&1
python tools/summarise_launches.py $OUT/launches.csv > $OUT/launches_summary.json 2>> $OUT/launches_bench.log; gzip -f $OUT/launches.csv
head -c 1500 $OUT/launches_summary.json
