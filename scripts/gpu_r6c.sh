#!/bin/bash
# r6c: scatter_packet.cu third run (vectorised packet kernel, prefetching scatter_inc kernels) + the full single-GPU
# evidence at this state: whole GPU suite, smoke, bench with default arguments
TAG=${1:-r6c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc > $OUT/nproc.txt
stamp pytest-new; timeout 400 python -m pytest tests/test_scatter_packet_gpu.py -q -m gpu --maxfail=20 -p no:cacheprovider > $OUT/pytest_scatter_packet.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" $OUT/pytest_scatter_packet.log | head -40
stamp time; timeout 300 python scripts/time_prims.py packet scatter_inc --reps 10 > $OUT/prims_scatter_packet.txt 2>&1; echo "time rc=$?"; cat $OUT/prims_scatter_packet.txt
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -20
stamp smoke; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
stamp bench; timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('$OUT/bench_n1.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], {k:(v['ms'], v['frac_of_peak_per_gpu']) for k,v in d['primitives'].items() if 'ms' in v})
print('roofline', d['roofline']['kernel'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
PY
stamp done
