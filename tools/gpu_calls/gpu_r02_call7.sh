#!/bin/bash
# r02 call 7: where does the time go?  launch list + ncu --set full with source for the matvec kernels and the new eigen-solver kernels
OUT=gpurun_out/r02c7
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_target.csv python tools/ncu_target.py > $OUT/launches_target.log 2>&1
grep -E "k_gemm|k_wapply|k_hetrd|k_ormtr|k_stedc" $OUT/launches_target.csv | awk -F'","' '{print $5, $NF}' | cut -c1-200 | tail -24
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm|k_hetrd_blocked|k_ormtr_rows|k_stedc" -s 4 -c 7 -o $OUT/prof_target python tools/ncu_target.py > $OUT/ncu_full.log 2>&1; tail -3 $OUT/ncu_full.log
ls -la $OUT
