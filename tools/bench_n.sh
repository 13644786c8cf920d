#!/bin/bash
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 10 --warmup 3 --quick 2>gpurun_out/r2f_gpu$N.err | grep '^{' > gpurun_out/r2f_gpu$N.jsonl; cut -c1-200 gpurun_out/r2f_gpu$N.jsonl
