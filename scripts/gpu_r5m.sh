#!/bin/bash
# Round-2 final single-GPU visit r5m: full GPU suite, smoke, both bench arms with the default arguments
# (what the driver runs), launch list of the bench command, ncu --set full of the hot kernels (raw pages).
TAG=${1:-r5m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc > $OUT/nproc.txt
stamp pytest; timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -20
stamp smoke; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
stamp bench-ref; timeout 900 python bench.py --impl reference > $OUT/bench_reference_cpu.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"; tail -c 700 $OUT/bench_reference_cpu.json
stamp bench; timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('$OUT/bench_n1.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], {k:(v['ms'], v['frac_of_peak_per_gpu']) for k,v in d['primitives'].items() if 'ms' in v})
print('roofline', d['roofline']['kernel'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pcie_GBps'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
PY
stamp ncu-launches; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
stamp ncu-full
timeout 600 ncu --set full --clock-control none -k "regex:reduce_chunk|reduce_group|prefix_reduce|compress_kernel|mkperm|scatter_reduce" -s 20 -c 11 -f -o /tmp/full_bench \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-verify --no-extras > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/full_bench.ncu-rep --page raw --csv > $OUT/full_bench.csv 2>/dev/null; ls -la $OUT/full_bench.csv
stamp done
