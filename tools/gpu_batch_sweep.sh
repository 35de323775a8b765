#!/bin/bash
mkdir -p gpurun_out
for bf in 565 600 640 1130; do
  timeout 300 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --batch-frames $bf > gpurun_out/bench_bf$bf.json 2> gpurun_out/bench_bf$bf.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_bf$bf.json'));k=d['kernels'];print('batch $bf frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'gemm TF/s',round(d['roofline']['achieved']), 'ln ms',round(k['layernorm']['ms'],1),'attn ms',round(k['attention']['ms'],1),'resid',round(k['gemm_resid']['ms'],1),'gelu',round(k['gemm_gelu']['ms'],1),'qkv',round(k['gemm_qkv']['ms'],1),'mhz',d['clocks']['sm_mhz'])"
done
