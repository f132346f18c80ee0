#!/bin/bash
# Two-GPU visit: NCCL parity test of the sharded path + the bench line at N=2 (torchrun).
TAG=${1:-r1_n2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1; cat $OUT/gpus.txt
stamp pytest-dist; timeout 600 python -m pytest tests/test_dist_gpu.py -q -m gpu > $OUT/pytest_dist.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_dist.log
stamp bench-n2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?"; tail -c 3000 $OUT/bench_n2.json; tail -5 $OUT/bench_n2.err
stamp bench-n1
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"; tail -c 600 $OUT/bench_n1.json
stamp done
