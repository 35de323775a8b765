mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_attn.log
timeout 600 python tools/attn_variants.py only split_default,split_s0a_late,split_delay2400,split_delay4000 > gpurun_out/attn_variants11.log 2>&1; grep "n= 70\|n=1130" gpurun_out/attn_variants11.log
CRE_B200_LIB=tools/libcre_b200_trace.so timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace_r02d.log 2>&1; tail -22 gpurun_out/attn_trace_r02d.log
for flags in "" "--tune attention_split_mode=4"; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline --breakdown $flags > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
echo "== [$flags] rc=$?"; grep "attention  " gpurun_out/bench_g.err; python -c "
import json;d=json.load(open('gpurun_out/bench_g.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'fwd',d['roofline']['vit_forward']['frac_of_burst_peak'], 'clocks',d['clocks']['sm_mhz'])"
done
