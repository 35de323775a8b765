mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "resid or folded or row_stats" > gpurun_out/pytest_resid.log 2>&1; echo "pytest kernels rc=$?"; tail -5 gpurun_out/pytest_resid.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "forward_tokens or golden or cosine" > gpurun_out/pytest_fwd.log 2>&1; echo "pytest fwd rc=$?"; tail -5 gpurun_out/pytest_fwd.log
timeout 300 python tools/resid_bench.py > gpurun_out/resid_bench.log 2>&1; cat gpurun_out/resid_bench.log
i=0
for flags in "--tune resid_split=1" "--tune resid_split=0" "--tune resid_split=1 --tune resid_ln_deep=3" "--tune resid_split=1 --tune resid_ln_deep=0"; do
  i=$((i+1))
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $flags > gpurun_out/bench_sp$i.json 2> gpurun_out/bench_sp$i.err
  echo "== [$flags] rc=$?"; tail -3 gpurun_out/bench_sp$i.err; python -c "
import json;d=json.load(open('gpurun_out/bench_sp$i.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'fwd',d['roofline']['vit_forward']['frac_of_burst_peak'], 'resid us',d['kernels']['gemm_resid']['avg_us'],'clocks',d['clocks']['sm_mhz'])"
done
