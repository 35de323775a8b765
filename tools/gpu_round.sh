#!/bin/bash
# One gpurun call: GPU parity tests, bench, ncu launch list and a full capture of the first layer's kernels
# (batch 1130 frames = the bench's launch size, so per-launch DRAM traffic is comparable with bench.py's roofline).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os, torch; print('cpus', os.cpu_count(), 'threads', torch.get_num_threads())" >> gpurun_out/gpu.txt 2>&1
free -g >> gpurun_out/gpu.txt 2>&1
# (GPU parity tests: tools/gpu_quick2.sh)

timeout 600 python bench.py --steps 3 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -30 gpurun_out/bench.err; cat gpurun_out/bench.json
SMALL="python bench.py --clips 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --gallery-rows 100000"
timeout 300 $SMALL > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 $SMALL > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_tn|attention_split|preprocess|layernorm|final_norm" -s 0 -c 11 -f -o gpurun_out/prof_layer0 $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
# the remaining kernels of the step (final norm + token mean, row statistics, prefix fill) ...
timeout 300 $SMALL > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"final_norm_mean|row_stats|fill_prefix|attention_exact" -c 4 -f -o gpurun_out/prof_tail $SMALL > gpurun_out/ncu_tail.log 2>&1
echo "ncu tail rc=$?"
# ... and K3 + both forms of K4 at the bench's sizes (second round of tools/topk_profile.py: 5 launches after the first 5)
timeout 300 python tools/topk_profile.py > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"pool_clips|topk_prepare|merge_topk|gallery_scan_small|gemm_tn_kernel<6" -s 5 -c 5 -f -o gpurun_out/prof_reid python tools/topk_profile.py > gpurun_out/ncu_reid.log 2>&1
echo "ncu reid rc=$?"
# the long-sequence attention kernel at 518 x 518 (configs[2]), first launch of a small run
LONG="python bench.py --resize 518 --height 518 --width 518 --clips 2 --frames-per-clip 111 --batch-frames 222 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-baseline"
timeout 300 $LONG > gpurun_out/plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attention_long" -s 1 -c 1 -f -o gpurun_out/prof_attn_long2 $LONG > gpurun_out/ncu_attn_long2.log 2>&1
echo "ncu long rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
for r in prof_layer0 prof_tail prof_reid prof_attn_long2; do
  [ -f gpurun_out/$r.ncu-rep ] && ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
