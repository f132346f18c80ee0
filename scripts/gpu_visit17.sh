#!/bin/bash
TAG=${1:-r1q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python scripts/small_sizes.py > $OUT/small_sizes.txt 2>&1; echo "rc=$?"; grep -v "jit_\|launching\|Found CUDA" $OUT/small_sizes.txt | tail -40
