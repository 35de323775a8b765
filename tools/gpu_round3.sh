#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | head
timeout 300 python tools/topk_bench.py > gpurun_out/topk_bench.log 2>&1; echo "topk_bench rc=$?"; cat gpurun_out/topk_bench.log
