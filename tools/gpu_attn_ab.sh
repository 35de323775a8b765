mkdir -p gpurun_out
timeout 600 python tools/attn_variants.py only split_default,split_spin_softmax,split_spin_issuer,split_spin_both > gpurun_out/attn_variants10.log 2>&1; cat gpurun_out/attn_variants10.log
for flags in "" "--tune attention_split_mode=12"; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline --breakdown $flags > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
echo "== [$flags] rc=$?"; grep "attention  " gpurun_out/bench_g.err; python -c "
import json;d=json.load(open('gpurun_out/bench_g.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'fwd',d['roofline']['vit_forward']['frac_of_burst_peak'], 'clocks',d['clocks']['sm_mhz'])"
done
