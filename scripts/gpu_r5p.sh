#!/bin/bash
# Round-2 visit r5p (one GPU): compute-sanitizer racecheck over the GPU test suite (the 2^30 / baseline-size
# tests excluded: they take too long under the tool), incl. sort, call_reduce and the fused sharded paths at world 1.
TAG=${1:-r5p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp racecheck-parity; timeout 1500 compute-sanitizer --tool racecheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
   -k "not 2_30 and not baseline and not large and not reference_grid and not small_arrays and not int_ops and not literals" > $OUT/racecheck_parity.log 2>&1; tail -3 $OUT/racecheck_parity.log
stamp racecheck-other; timeout 1500 compute-sanitizer --tool racecheck --print-limit 30 python -m pytest tests/test_sort_gpu.py tests/test_call_reduce_gpu.py tests/test_comm_gpu.py tests/test_history_graph_gpu.py -q -m gpu \
   -k "not 2_26 and not staged" > $OUT/racecheck_other.log 2>&1; tail -3 $OUT/racecheck_other.log
stamp done
