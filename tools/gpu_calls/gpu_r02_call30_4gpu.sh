#!/bin/bash
# r02 call 30 (4 GPUs): the driver's launch line for N = 4 on the final build (short run)
OUT=gpurun_out/r02c30
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l; nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 2 --warmup 2 > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.err; echo "rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_4gpu.json'));print('4gpu value', d.get('value'), 'e2e', d.get('e2e',{}).get('value'), d.get('step_ms'), d.get('n_gpus'), d.get('config',{}).get('global_batch'), d.get('config',{}).get('lanes_per_gpu'), d.get('consistency',{}).get('status_nonzero'), d.get('consistency',{}).get('trajectories'), d.get('error'))"
tail -2 $OUT/bench_4gpu.err
