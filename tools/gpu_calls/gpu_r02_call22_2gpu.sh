#!/bin/bash
# r02 call 22 (2 GPUs): the driver's launch line for N = 2 on the final build (short run) + the reference arm under the same launcher
OUT=gpurun_out/r02c22
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L; nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err; echo "rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_2gpu.json'));print('2gpu value', d.get('value'), 'e2e', d.get('e2e',{}).get('value'), d.get('step_ms'), d.get('n_gpus'), d.get('config',{}).get('global_batch'), d.get('config',{}).get('lanes_per_gpu'), d.get('consistency',{}).get('status_nonzero'), d.get('consistency',{}).get('trajectories'), d.get('error'))"
tail -3 $OUT/bench_2gpu.err
