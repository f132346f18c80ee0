#!/bin/bash
# r6d: scatter_inc with the per-tile coherence probe; the reference's JIT kernels timed through KernelHistory; ncu of the
# vectorised packet kernel and the prefetching scatter_inc kernels
TAG=${1:-r6d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest-new; timeout 400 python -m pytest tests/test_scatter_packet_gpu.py tests/test_gpu_parity.py -k "scatter" -q -m gpu --maxfail=20 -p no:cacheprovider > $OUT/pytest_scatter.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" $OUT/pytest_scatter.log | head -40
stamp time; timeout 300 python scripts/time_prims.py scatter_inc --reps 10 > $OUT/prims_scatter_inc.txt 2>&1; echo "time rc=$?"; cat $OUT/prims_scatter_inc.txt
stamp ref; timeout 200 python scripts/time_scatter_ref.py 2>&1 | grep "^reference\|^drjit_b200\|Error\|error" > $OUT/time_scatter_ref.txt; cat $OUT/time_scatter_ref.txt
stamp ncu
timeout 300 ncu --set full --clock-control none -k "regex:scatter_packet|scatter_inc" -c 8 -f -o /tmp/full_sp \
    python scripts/time_prims.py packet4 packet8 inc_queue_m inc_16 --reps 1 --warm 1 > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/full_sp.ncu-rep --page raw --csv > $OUT/full_scatter_packet.csv 2>/dev/null; ls -la $OUT/full_scatter_packet.csv
stamp done
