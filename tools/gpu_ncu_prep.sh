#!/bin/bash
# ncu --set full of the K1 TMA kernels (second form = mode 2, first form = mode 1) at 600 x 1080p
mkdir -p gpurun_out
CMD="env PREP_SIZES=1080x1920 python tools/prep_bench.py 2 1"
timeout 300 $CMD > gpurun_out/prep_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:preprocess_tma -s 5 -c 1 -f -o gpurun_out/r02d_prof_k1_v2 $CMD > gpurun_out/ncu_k1_v2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_k1_v2.log
ncu -i gpurun_out/r02d_prof_k1_v2.ncu-rep --page raw --csv > gpurun_out/r02d_prof_k1_v2_raw.csv 2>/dev/null
ncu -i gpurun_out/r02d_prof_k1_v2.ncu-rep --page source --csv > gpurun_out/r02d_prof_k1_v2_src.csv 2>/dev/null
