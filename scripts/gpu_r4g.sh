#!/bin/bash
# Round-2 visit G (one GPU, experiments build): mkperm tile order A/B, 8-byte scan geometries.
TAG=${1:-r4g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
export DRJIT_B200_LIB=$PWD/build/exp/libdrjit_b200.so
stamp mkperm-order
for D in 0 16; do DRJIT_B200_MKPERM_DEBUG=$D timeout 60 python scripts/time_prims.py mkperm --reps 20 > $OUT/prims_mkperm_order$D.txt 2>&1; cat $OUT/prims_mkperm_order$D.txt; done
DRJIT_B200_MKPERM_DEBUG=16 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mkperm" > $OUT/pytest_mkperm_order16.log 2>&1; echo "order16 rc=$?"; tail -3 $OUT/pytest_mkperm_order16.log
stamp scan64
for G in 0 1 2 3 4; do echo "geom $G"; DRJIT_B200_SCAN64_GEOM=$G timeout 60 python scripts/time_prims.py scan64 --reps 20 2>&1 | tee -a $OUT/prims_scan64_geom.txt; done
DRJIT_B200_SCAN64_GEOM=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "prefix or scan" > $OUT/pytest_scan_geom1.log 2>&1; echo "geom1 rc=$?"; tail -3 $OUT/pytest_scan_geom1.log
stamp done
