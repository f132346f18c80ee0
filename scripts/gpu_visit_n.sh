#!/bin/bash
# Multi-GPU visit: the sharded tests (peer-memory communicator in-process and over CUDA IPC, NCCL
# baseline) and the strong-scaling bench line at N = $1 ranks (torchrun), then N=1 on the same box.
N=${1:-2}; TAG=${2:-r4_n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1; cat $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
stamp pytest-dist; timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_comm_gpu.py -q -m gpu > $OUT/pytest_dist.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_dist.log
stamp bench-n$N
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?"; tail -c 5000 $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err
stamp isolated
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    scripts/time_sharded.py > $OUT/time_sharded_n$N.txt 2> $OUT/time_sharded_n$N.err; echo "rc=$?"; cat $OUT/time_sharded_n$N.txt; tail -3 $OUT/time_sharded_n$N.err
stamp bench-n1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-e2e > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"; tail -c 400 $OUT/bench_n1.json
stamp done
