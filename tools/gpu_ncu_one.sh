#!/bin/bash
# usage: gpu_ncu_one.sh <kernel-regex> <out-name> [skip]
mkdir -p gpurun_out
SMALL="python bench.py --clips 2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --batch-frames 300"
timeout 300 $SMALL > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-2} -c 1 -f -o gpurun_out/$2 $SMALL > gpurun_out/ncu_$2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$2.log
