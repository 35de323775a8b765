#!/bin/bash
# GELU epilogue change: parity suite, headline bench, then the other BASELINE configs on one GPU (ViT-L/16, 224x224 input clips)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "gelu epilogue|passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | head
timeout 600 python bench.py --steps 3 --warmup 3 --breakdown > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
head -8 gpurun_out/bench_r2.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r2.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'gemm TF/s',round(d['roofline']['achieved']),'fwd',round(d['roofline']['vit_forward']['tflops']),'e2e',round(d['e2e']['value']),'cpu',d['cpu_baseline']['value'],'clocks',d['clocks'])"
timeout 600 python bench.py --model vitl16 --height 224 --width 224 --steps 2 --warmup 2 --breakdown --no-cpu-baseline > gpurun_out/bench_vitl.json 2> gpurun_out/bench_vitl.err; echo "vitl rc=$?"
head -8 gpurun_out/bench_vitl.err
python -c "
import json;d=json.load(open('gpurun_out/bench_vitl.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'gemm TF/s',round(d['roofline']['achieved']),'fwd',round(d['roofline']['vit_forward']['tflops']),'vit_tflops',round(d['vit_tflops']),'e2e',round(d['e2e']['value']))"
timeout 600 python bench.py --height 224 --width 224 --clips 1024 --steps 1 --warmup 1 --breakdown --no-cpu-baseline > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; echo "config4 n1 rc=$?"
head -6 gpurun_out/bench_c4_n1.err
python -c "
import json;d=json.load(open('gpurun_out/bench_c4_n1.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'vit_tflops',round(d['vit_tflops']),'e2e',round(d['e2e']['value']))"
