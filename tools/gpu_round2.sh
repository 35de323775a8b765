#!/bin/bash
# full GPU parity + attention standalone timings + headline bench (device-timed) + configs[2] lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/attn_variants.py > gpurun_out/attn_variants9.log 2>&1; grep "n=1130\|n= 600" gpurun_out/attn_variants9.log
timeout 300 python tools/resid_bench.py > gpurun_out/resid_bench2.log 2>&1; cat gpurun_out/resid_bench2.log
i=0
for flags in "" "--resize 518 --height 518 --width 518 --clips 4 --frames-per-clip 111 --batch-frames 222" "--resize 592 --height 592 --width 592 --clips 4 --frames-per-clip 111 --batch-frames 222"; do
  i=$((i+1))
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline --breakdown $flags > gpurun_out/bench_e$i.json 2> gpurun_out/bench_e$i.err
  echo "== [$flags] rc=$?"; tail -16 gpurun_out/bench_e$i.err; python -c "
import json;d=json.load(open('gpurun_out/bench_e$i.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'fwd',d['roofline']['vit_forward'], 'clocks',d['clocks'])"
done
