#!/bin/bash
TAG=${1:-r1u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=20 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -30
stamp prims; timeout 300 python scripts/time_prims.py mkperm mkperm256 > $OUT/prims.log 2>&1; cat $OUT/prims.log
stamp prims-unordered; DRJIT_B200_MKPERM_UNORDERED=1 timeout 300 python scripts/time_prims.py mkperm256 > $OUT/prims_unordered.log 2>&1; cat $OUT/prims_unordered.log
stamp prims-warp; DRJIT_B200_MKPERM_TILES=0 timeout 300 python scripts/time_prims.py mkperm256 > $OUT/prims_warp.log 2>&1; cat $OUT/prims_warp.log
stamp done
