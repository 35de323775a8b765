mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "preprocess" > gpurun_out/pytest_prep.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_prep.log
timeout 300 python tools/prep_bench.py > gpurun_out/prep_bench_r02e.log 2>&1; cat gpurun_out/prep_bench_r02e.log
