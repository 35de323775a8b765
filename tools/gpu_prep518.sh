mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -q -x -k "preprocess or large_grid" > gpurun_out/pytest_prep.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_prep.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --breakdown --resize 518 --height 518 --width 518 --clips 4 --frames-per-clip 111 --batch-frames 222 > gpurun_out/bench_518.json 2> gpurun_out/bench_518.err
echo "rc=$?"; grep "preprocess\|attention  " gpurun_out/bench_518.err; python -c "
import json;d=json.load(open('gpurun_out/bench_518.json'));print('frames/s',round(d['value']),'e2e',round(d['e2e']['value']),'fwd',d['roofline']['vit_forward']['frac_of_burst_peak'])"
