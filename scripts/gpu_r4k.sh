#!/bin/bash
# Round-2 visit K (one GPU, experiments build): cluster-of-two mkperm scatter kernel, parity + timing.
TAG=${1:-r4k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
export DRJIT_B200_LIB=$PWD/build/exp/libdrjit_b200.so
stamp parity; DRJIT_B200_MKPERM_PAIR=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_call_reduce_gpu.py -q -m gpu -k "mkperm or call_reduce" > $OUT/pytest_pair.log 2>&1; echo "pair rc=$?"; tail -6 $OUT/pytest_pair.log
stamp timing
for P in 0 1; do DRJIT_B200_MKPERM_PAIR=$P timeout 60 python scripts/time_prims.py mkperm --reps 20 2>&1 | tee $OUT/prims_mkperm_pair$P.txt; done
exit 0
stamp sanitize
DRJIT_B200_MKPERM_PAIR=1 timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mkperm_unordered_tiles_ragged" > $OUT/racecheck_pair.txt 2>&1; tail -5 $OUT/racecheck_pair.txt
stamp done
