#!/bin/bash
# Round-2 visit r5f (one GPU): why the 2^30 scan went from 1.305 ms (r4a) to 1.346 ms (r5a): shipped build
# (single-CTA branches compiled out of the TMA kernel again) against the same build without evict-first
# stores; segmented-scan sweep with the 1 MiB block limit.
TAG=${1:-r5f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
for V in shipped ab_NO_EVICT_FIRST; do
  if [ $V = shipped ]; then unset DRJIT_B200_LIB; else export DRJIT_B200_LIB=$PWD/build/$V/libdrjit_b200.so; fi
  stamp "prims $V"; timeout 300 python scripts/time_prims.py scan compress --reps 20 > $OUT/prims_$V.txt 2>&1; cat $OUT/prims_$V.txt
  stamp "bench $V"; timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e > $OUT/bench_$V.json 2>/dev/null
  python -c "
import json; d=json.loads([l for l in open('$OUT/bench_$V.json') if l.startswith('{')][-1]); print('$V:', d['ms_per_step'], {k:v['ms'] for k,v in d['primitives'].items() if 'ms' in v})"
done
unset DRJIT_B200_LIB
stamp pytest; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "prefix" --maxfail=10 > $OUT/pytest_prefix.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_prefix.log | head
stamp segsweep; timeout 300 python scripts/time_scanseg.py > $OUT/scanseg_sweep.txt 2>&1; grep -E "1000|100000|1048576|   256 " $OUT/scanseg_sweep.txt
stamp done
