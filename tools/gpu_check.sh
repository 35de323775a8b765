#!/bin/bash
# correctness + quick perf after a kernel change
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | head -20
timeout 300 python tools/gemm_tune.py > gpurun_out/gemm_tune.log 2>&1; tail -32 gpurun_out/gemm_tune.log
for cg in 1 2; do
  timeout 600 python bench.py --steps 2 --warmup 2 --breakdown --no-e2e --no-cpu-baseline --cta-group $cg > gpurun_out/bench_cg$cg.json 2> gpurun_out/bench_cg$cg.err
  echo "== cta_group $cg: rc=$?"; head -12 gpurun_out/bench_cg$cg.err; python -c "
import json;d=json.load(open('gpurun_out/bench_cg$cg.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'gemm TF/s',round(d['roofline']['achieved']),'clocks',d['clocks'])"
done
