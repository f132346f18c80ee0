#!/bin/bash
# r6e: compute-sanitizer racecheck over the shared-memory kernels of scatter_packet.cu (scatter_inc private / queue forms)
# and memcheck over the packet kernels (small cases; the round's last GPU seconds)
TAG=${1:-r6e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp racecheck; timeout 85 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest -q -m gpu -p no:cacheprovider 'tests/test_scatter_packet_gpu.py::test_scatter_inc_coherent_and_skewed_warps' 'tests/test_scatter_packet_gpu.py::test_scatter_inc_queue_form[5]' 'tests/test_scatter_packet_gpu.py::test_scatter_inc_queue_form[2049]' 'tests/test_scatter_packet_gpu.py::test_scatter_inc_errors_and_out_of_range' 'tests/test_scatter_packet_gpu.py::test_scatter_inc_reference_grid[3]' 'tests/test_scatter_packet_gpu.py::test_scatter_inc_reference_grid[100]' > $OUT/racecheck_scatter_inc.log 2>&1; tail -3 $OUT/racecheck_scatter_inc.log
stamp memcheck; timeout 50 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest -q -m gpu -p no:cacheprovider 'tests/test_scatter_packet_gpu.py::test_packet_unaligned_target_and_inputs[1]' 'tests/test_scatter_packet_gpu.py::test_packet_unaligned_target_and_inputs[2]' 'tests/test_scatter_packet_gpu.py::test_packet_argument_errors' 'tests/test_scatter_packet_gpu.py::test_packet_f16[2-add]' 'tests/test_scatter_packet_gpu.py::test_packet_f16[12-max]' 'tests/test_scatter_packet_gpu.py::test_scatter_inc_queue_form[2048]' > $OUT/memcheck_scatter_packet.log 2>&1; tail -3 $OUT/memcheck_scatter_packet.log
stamp done
