#!/bin/bash
TAG=${1:-r1n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python scripts/incumbent.py --log2-shift 4 > $OUT/incumbent_small.txt 2>&1; echo "rc=$?"; tail -12 $OUT/incumbent_small.txt
timeout 600 python scripts/incumbent.py > $OUT/incumbent.txt 2>&1; echo "rc=$?"; tail -12 $OUT/incumbent.txt
