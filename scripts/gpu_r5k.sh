#!/bin/bash
# Round-2 visit r5k (one GPU): call_reduce with the payload tile staged in shared memory.
TAG=${1:-r5k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests/test_call_reduce_gpu.py tests/test_gpu_parity.py -q -m gpu -k "call_reduce or mkperm" --maxfail=10 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -20
stamp prims; timeout 300 python scripts/time_prims.py mkperm call_reduce > $OUT/prims.txt 2>&1; cat $OUT/prims.txt
stamp sanitizer; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_call_reduce_gpu.py -q -m gpu -k "staged and 4096-1" > $OUT/memcheck.log 2>&1; tail -4 $OUT/memcheck.log
stamp done
