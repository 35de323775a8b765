#!/bin/bash
# ncu --set full of the attention-out projection (K = 768) with the split-residual epilogues, standalone at the bench's launch size
mkdir -p gpurun_out
CMD="python tools/resid_bench.py --k 768 --epis sp sp3 --iters 1"
timeout 300 $CMD > gpurun_out/resid_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 3 -c 5 -f -o gpurun_out/r02c_prof_resid_sp $CMD > gpurun_out/ncu_resid_sp.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_resid_sp.log
ncu -i gpurun_out/r02c_prof_resid_sp.ncu-rep --page raw --csv > gpurun_out/r02c_prof_resid_sp_raw.csv 2>/dev/null
ncu -i gpurun_out/r02c_prof_resid_sp.ncu-rep --page source --csv > gpurun_out/r02c_prof_resid_sp_src.csv 2>/dev/null
ls -la gpurun_out/r02c_prof_resid_sp*
