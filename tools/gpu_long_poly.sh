mkdir -p gpurun_out
for flags in "--tune attention_poly=0" "--tune attention_poly=1" "--tune attention_poly=2"; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline --breakdown --resize 518 --height 518 --width 518 --clips 4 --frames-per-clip 111 --batch-frames 222 $flags > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
  echo "== [$flags] rc=$?"; grep "attention  " gpurun_out/bench_d.err
done
