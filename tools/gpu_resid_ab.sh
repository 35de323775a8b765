mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -q -x -k "resid or folded or forward_tokens or golden" > gpurun_out/pytest_resid.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_resid.log
timeout 300 python tools/resid_bench.py > gpurun_out/resid_bench3.log 2>&1; cat gpurun_out/resid_bench3.log
CRE_B200_LIB=tools/libcre_b200_trace.so timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace_r02c.log 2>&1; tail -32 gpurun_out/attn_trace_r02c.log
for flags in "" "--tune resid_ln_deep=3"; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline --breakdown $flags > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
echo "== [$flags] rc=$?"; grep "gemm_resid \|gemm_resid_mlp\|attention  " gpurun_out/bench_f.err; python -c "
import json;d=json.load(open('gpurun_out/bench_f.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'fwd',d['roofline']['vit_forward']['frac_of_burst_peak'], 'clocks',d['clocks']['sm_mhz'])"
done
