#!/bin/bash
# Round-2 visit r5i (one GPU): full GPU suite at HEAD, all primitive timings (64-bit scan after the
# single-CTA branches left the TMA instantiation: spills 80 -> 8 bytes), small-size table, bench line.
TAG=${1:-r5i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=10 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -20
stamp prims; timeout 300 python scripts/time_prims.py all > $OUT/prims.txt 2>&1; cat $OUT/prims.txt
stamp small; timeout 600 python scripts/small_sizes.py 2>/dev/null > $OUT/small_sizes.txt; head -40 $OUT/small_sizes.txt
stamp bench; timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"; tail -2 $OUT/bench_n1.err
python - <<PY
import json
d=json.loads([l for l in open('$OUT/bench_n1.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], {k:(v['ms'], v['frac_of_peak_per_gpu']) for k,v in d['primitives'].items() if 'ms' in v})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pcie_GBps'], 'cpu', d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)
PY
stamp done
