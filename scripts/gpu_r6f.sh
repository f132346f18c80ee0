#!/bin/bash
# r6f: wider sanitizer pass over tests/test_scatter_packet_gpu.py (racecheck: every scatter_inc test but the 2^22 one;
# memcheck: the packet tests but the film / reference-JIT ones)
TAG=${1:-r6f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp racecheck; timeout 62 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_scatter_packet_gpu.py -q -m gpu -p no:cacheprovider \
   -k "scatter_inc and not 4194381 and not reference_cuda and not 65536" > $OUT/racecheck_scatter_inc_all.log 2>&1; tail -3 $OUT/racecheck_scatter_inc_all.log
stamp memcheck; timeout 45 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_scatter_packet_gpu.py -q -m gpu -p no:cacheprovider \
   -k "packet_other_ops or packet_f16 or (modes_agree and Local)" > $OUT/memcheck_packet_all.log 2>&1; tail -3 $OUT/memcheck_packet_all.log
stamp done
