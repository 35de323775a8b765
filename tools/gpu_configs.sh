mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_attn.log
i=0
for flags in "--resize 518 --height 518 --width 518 --clips 4 --frames-per-clip 111 --batch-frames 222" "--resize 592 --height 592 --width 592 --clips 4 --frames-per-clip 111 --batch-frames 222" "--height 224 --width 224 --clips 1024 --frames-per-clip 150 --steps 2" "--model vitl16 --height 224 --width 224"; do
  i=$((i+1))
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --breakdown $flags > gpurun_out/bench_c$i.json 2> gpurun_out/bench_c$i.err
  echo "== [$flags] rc=$?"; grep "attention  \|gemm_topk\|merge_topk" gpurun_out/bench_c$i.err; python -c "
import json;d=json.load(open('gpurun_out/bench_c$i.json'));print(d['config']['workload'][:60]);print('frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'e2e',d['e2e']['value'] if d.get('e2e') else None,'fwd',d['roofline']['vit_forward']['frac_of_burst_peak'], 'clocks',d['clocks']['sm_mhz'])"
done
