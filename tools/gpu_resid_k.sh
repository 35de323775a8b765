mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -q -x -k "resid or folded or forward_tokens or golden" > gpurun_out/pytest_resid.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_resid.log
timeout 300 python tools/resid_bench.py --k 64 768 3072 --epis sp sp3 > gpurun_out/resid_bench_k2.log 2>&1; cat gpurun_out/resid_bench_k2.log
