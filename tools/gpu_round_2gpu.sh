#!/bin/bash
# 2-GPU plumbing check under torchrun (NCCL init, sharded Simulator.run, the all-reduce, rank-0-only JSON), then a short C3 run.
OUT=gpurun_out/r01h
mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 150 $TR bench.py --gpus 2 --length 12 --chi 8 --batch 8 --steps 2 --warmup 3 > $OUT/bench_2gpu_tiny.json 2> $OUT/bench_2gpu_tiny.err; echo "tiny rc=$?"
tail -c 900 $OUT/bench_2gpu_tiny.json; tail -3 $OUT/bench_2gpu_tiny.err
timeout 60 $TR bench.py --impl reference --gpus 2 --length 12 --chi 8 --steps 1 --warmup 1 > $OUT/bench_2gpu_tiny_reference.json 2> $OUT/bench_2gpu_tiny_reference.err; echo "ref rc=$?"
tail -c 300 $OUT/bench_2gpu_tiny_reference.json
timeout 200 $TR bench.py --gpus 2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-roofline > $OUT/bench_2gpu_c3_short.json 2> $OUT/bench_2gpu_c3_short.err; echo "c3 rc=$?"
python tools/bench_brief.py $OUT/bench_2gpu_c3_short.json
