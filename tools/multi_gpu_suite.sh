#!/bin/bash
# usage: tools/multi_gpu_suite.sh N TAG   -- the per-N measurements kept under profiles/ (weak-scaling bench line, dialogue-sharded
# scoring sweeps of BASELINE configs 3/4, fixed-size training sweep); one JSON line per run appended to gpurun_out/TAG_gpuN.jsonl
N=$1; TAG=${2:-r2}
out=gpurun_out/${TAG}_gpu${N}.jsonl
: > $out
run() { if [ "$N" = "1" ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; fi; }
if [ "$N" = "2" ]; then python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/${TAG}_pytest_dp_2gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_dp_2gpu.log; fi
run bench.py --gpus $N --steps 10 --warmup 3 --quick 2>gpurun_out/${TAG}_gpu${N}.err | grep '^{' >> $out
run tools/score_sweep.py --utterances 1000000 2>>gpurun_out/${TAG}_gpu${N}.err | grep '^{' >> $out
run tools/score_sweep.py --meld --utterances 200000 2>>gpurun_out/${TAG}_gpu${N}.err | grep '^{' >> $out
run tools/train_sweep.py --utterances 1000000 2>>gpurun_out/${TAG}_gpu${N}.err | grep '^{' >> $out
wc -l $out; cut -c1-400 $out
