#!/bin/bash
# parity + a short A/B bench: usage gpu_quick.sh "<bench flags A>" "<bench flags B>" ...
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu.log | head -30
i=0
for flags in "$@"; do
  i=$((i+1))
  timeout 600 python bench.py --steps 2 --warmup 2 --breakdown --no-e2e --no-cpu-baseline $flags > gpurun_out/bench_q$i.json 2> gpurun_out/bench_q$i.err
  echo "== [$flags] rc=$?"; head -12 gpurun_out/bench_q$i.err; python -c "
import json;d=json.load(open('gpurun_out/bench_q$i.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'gemm TF/s',round(d['roofline']['achieved']),'clocks',d['clocks'])"
done
