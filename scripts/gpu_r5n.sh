#!/bin/bash
# Round-2 visit r5n (one GPU): one-CTA scan of small arrays (prefix_small.cu): parity (all prefix tests, the
# sharded paths at world 1, graph capture), per-call table against the reference's CUDA kernels.
TAG=${1:-r5n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_comm_gpu.py tests/test_history_graph_gpu.py tests/test_incumbent_gpu.py tests/test_insitu_gpu.py -q -m gpu -k "prefix or scan or fused or graph or history or insitu or reference" --maxfail=10 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -20
stamp small; timeout 600 python scripts/small_sizes.py 2>/dev/null > $OUT/small_sizes.txt; grep -E "primitive|prefix" $OUT/small_sizes.txt
stamp sanitizer; timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "small_arrays and u32" > $OUT/racecheck.log 2>&1; tail -3 $OUT/racecheck.log
stamp done
