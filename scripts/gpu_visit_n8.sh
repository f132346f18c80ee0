#!/bin/bash
# Eight-GPU visit: the bench line at N=8 (torchrun), nothing else (8x charge).
TAG=${1:-r1_n8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1; cat $OUT/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "rc=$?"; tail -c 3500 $OUT/bench_n8.json; tail -5 $OUT/bench_n8.err
