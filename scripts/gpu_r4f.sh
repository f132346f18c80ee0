#!/bin/bash
# Round-2 visit F (one GPU): full GPU test-suite with the paired compress staging in the product,
# call_reduce payload timings, mkperm tile-order A/B (experiments build), bench line.
TAG=${1:-r4f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=10 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -20
stamp prims; timeout 300 python scripts/time_prims.py call_reduce compress compress01 compress99 --reps 10 > $OUT/prims.txt 2>&1; cat $OUT/prims.txt
stamp mkperm-order
export DRJIT_B200_LIB=$PWD/build/exp/libdrjit_b200.so
for D in 0 16; do DRJIT_B200_MKPERM_DEBUG=$D timeout 60 python scripts/time_prims.py mkperm --reps 20 > $OUT/prims_mkperm_order$D.txt 2>&1; cat $OUT/prims_mkperm_order$D.txt; done
DRJIT_B200_MKPERM_DEBUG=16 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mkperm" > $OUT/pytest_mkperm_order16.log 2>&1; echo "order16 rc=$?"; tail -3 $OUT/pytest_mkperm_order16.log
unset DRJIT_B200_LIB
stamp bench; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench.json')); print(d['value'], d['ms_per_step'], d['verified'], {k:(v['ms'],v['frac_of_peak_per_gpu']) for k,v in d['primitives'].items()}, d['e2e']['value'])"
stamp done
