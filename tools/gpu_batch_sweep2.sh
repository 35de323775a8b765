mkdir -p gpurun_out
for bf in 1130 1600 3200 4800 1130; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline --batch-frames $bf > gpurun_out/bench_bf$bf.json 2> gpurun_out/bench_bf$bf.err
echo "== [batch $bf] rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_bf$bf.json'));k=d['kernels'];print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'kernel-sum ms',round(sum(v['ms'] for v in k.values()),1),'launches',d['gpu_launches'],'fwd',round(d['roofline']['vit_forward']['frac_of_burst_peak'],4), 'clocks',d['clocks']['sm_mhz'])"
done
