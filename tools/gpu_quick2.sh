mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline --breakdown > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
echo "== rc=$?"; tail -16 gpurun_out/bench_h.err; python -c "
import json;d=json.load(open('gpurun_out/bench_h.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'fwd',d['roofline']['vit_forward']['frac_of_burst_peak'], 'clocks',d['clocks']['sm_mhz'])"
timeout 300 python tools/prep_bench.py > gpurun_out/prep_bench_r02c.log 2>&1; tail -8 gpurun_out/prep_bench_r02c.log
