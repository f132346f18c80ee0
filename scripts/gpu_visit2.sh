#!/bin/bash
OUT=gpurun_out/r1b; mkdir -p $OUT
echo "== microbench"; timeout 300 build/microbench > $OUT/microbench.log 2>&1; echo "rc=$?"; cat $OUT/microbench.log
echo "== failing tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_block_reduce_float_ops and mul-f32 or golden_fixture or test_dot" --tb=short 2>&1 | tail -60 | tee $OUT/failing.log
echo "== durations"; timeout 1100 python -m pytest tests -q -m gpu --durations=40 --timeout=400 > $OUT/pytest.log 2>&1; echo "rc=$?"; grep -A45 "slowest" $OUT/pytest.log | head -60; tail -5 $OUT/pytest.log
