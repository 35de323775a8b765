#!/bin/bash
# ncu launch list + full capture of one layer's kernels on a 2-clip (300-frame) step.
mkdir -p gpurun_out
SMALL="python bench.py --clips 2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --gallery-rows 100000"
timeout 300 $SMALL > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 $SMALL > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_tn|attention|preprocess|layernorm|final_norm" -s 0 -c 11 -f -o gpurun_out/prof_layer0 $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/plain.log
