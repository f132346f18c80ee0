#!/bin/bash
# r6a: first GPU run of scatter_packet.cu (packet scatter-reduce, scatter_inc): parity tests, timings, one ncu capture
TAG=${1:-r6a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
stamp pytest-new; timeout 400 python -m pytest tests/test_scatter_packet_gpu.py -q -m gpu --maxfail=20 -p no:cacheprovider > $OUT/pytest_scatter_packet.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" $OUT/pytest_scatter_packet.log | head -40
stamp time; timeout 300 python scripts/time_prims.py packet scatter_inc scatter --reps 10 > $OUT/prims_scatter_packet.txt 2>&1; echo "time rc=$?"; cat $OUT/prims_scatter_packet.txt
stamp ncu
timeout 300 ncu --set full --clock-control none -k "regex:scatter_packet|scatter_inc" -c 8 -f -o /tmp/full_sp \
    python scripts/time_prims.py packet4 inc_queue inc_16 inc_2^20 --reps 1 --warm 1 > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/full_sp.ncu-rep --page raw --csv > $OUT/full_scatter_packet.csv 2>/dev/null; ls -la $OUT/full_scatter_packet.csv
stamp done
