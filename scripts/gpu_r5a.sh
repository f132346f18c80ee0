#!/bin/bash
# Round-2 visit r5a (one GPU): full GPU suite after the exchange fix + fused count exchange + flattened
# all-reduce (world 1 runs every exchange code path), bench line with the reordered step, e2e A/B of
# write-combined upload buffers, launch list.
TAG=${1:-r5a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=10 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -20
stamp smoke; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
stamp bench; timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"; tail -3 $OUT/bench_n1.err
python - <<PY
import json
d=json.loads([l for l in open('$OUT/bench_n1.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], {k:(v['ms'], v['frac_of_peak_per_gpu']) for k,v in d['primitives'].items() if 'ms' in v})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pcie_GBps'], 'cpu', d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)
PY
stamp e2e-wc; DRJIT_B200_E2E_WC=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > $OUT/bench_e2e_wc.json 2> $OUT/bench_e2e_wc.err; echo "rc=$?"; tail -3 $OUT/bench_e2e_wc.err
stamp e2e-plain; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > $OUT/bench_e2e_plain.json 2> $OUT/bench_e2e_plain.err; echo "rc=$?"
python - <<PY
import json
for f in ('bench_e2e_wc','bench_e2e_plain'):
    d=json.loads([l for l in open('$OUT/'+f+'.json') if l.startswith('{')][-1])
    print(f, d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pcie_GBps'], d['e2e'].get('upload_buffers'))
PY
stamp prims; timeout 300 python scripts/time_prims.py all > $OUT/prims.txt 2>&1; cat $OUT/prims.txt
stamp ncu-launches; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
stamp done
