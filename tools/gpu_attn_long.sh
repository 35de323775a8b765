mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_attn.log
timeout 600 python tools/attn_variants.py only split_default,long_kernel > gpurun_out/attn_variants12.log 2>&1; cat gpurun_out/attn_variants12.log
for flags in "--resize 518 --height 518 --width 518 --clips 4 --frames-per-clip 111 --batch-frames 222" "--resize 592 --height 592 --width 592 --clips 4 --frames-per-clip 111 --batch-frames 222"; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --breakdown $flags > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
  echo "== [$flags] rc=$?"; grep "attention  " gpurun_out/bench_d.err; python -c "
import json;d=json.load(open('gpurun_out/bench_d.json'));print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'fwd',d['roofline']['vit_forward']['frac_of_burst_peak'], 'clocks',d['clocks']['sm_mhz'])"
  cp gpurun_out/bench_d.json gpurun_out/bench_long_$(echo $flags | cut -d' ' -f2).json
done
