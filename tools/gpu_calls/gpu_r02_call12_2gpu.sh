#!/bin/bash
# r02 call 12 (2 GPUs): the driver's launch line at N = 2, short, for both arms; NCCL all-reduce inside AnalogRun.result()
OUT=gpurun_out/r02c12
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "n2 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_n2.json'));print('N=2', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), d['step_ms'], d['n_gpus'], d['config']['global_batch'], d['consistency'])"; tail -3 $OUT/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > $OUT/ref_n2.json 2> $OUT/ref_n2.err; echo "ref rc=$?"; head -c 600 $OUT/ref_n2.json
