#!/bin/bash
# Round-2 visit 1 (one GPU): full GPU test-suite on the new library (incl. the world=1 peer-exchange
# paths, kernel history, graph capture), then the acceptance run of the experimental 60 Ki-key
# mkperm kernel (experiments build only).
TAG=${1:-r3a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
stamp "pytest -m gpu"
timeout 900 python -m pytest tests -q -m gpu -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/pytest.log
stamp "mkperm KPT=60 acceptance (experiments build)"
export DRJIT_B200_LIB=$PWD/build/exp/libdrjit_b200.so
DRJIT_B200_MKPERM_KPT=60 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mkperm" > $OUT/pytest_kpt60.log 2>&1; echo "kpt60 rc=$?"
tail -5 $OUT/pytest_kpt60.log
for K in 48 60; do
  stamp "prims KPT=$K"; DRJIT_B200_MKPERM_KPT=$K timeout 60 python scripts/time_prims.py mkperm --reps 20 > $OUT/prims_kpt$K.log 2>&1; cat $OUT/prims_kpt$K.log
done
unset DRJIT_B200_LIB
stamp "prims (shipped library)"
timeout 120 python scripts/time_prims.py all --reps 10 > $OUT/prims.txt 2>&1; cat $OUT/prims.txt
stamp done
