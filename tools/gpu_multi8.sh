#!/bin/bash
# usage: gpu_multi8.sh N -- configs[4] (ViT-L/16, 1024-d gallery row-sharded over N ranks) and configs[3] (1 024 clips of 224x224 over N ranks)
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $RUN bench.py --gpus $N --model vitl16 --height 224 --width 224 --steps 2 --warmup 3 > gpurun_out/bench_vitl_n$N.json 2> gpurun_out/bench_vitl_n$N.err; echo "configs[4] N=$N rc=$?"
tail -3 gpurun_out/bench_vitl_n$N.err
python -c "
import json;d=json.load(open('gpurun_out/bench_vitl_n$N.json'));print('N',d['n_gpus'],'frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']),'sharded',d.get('sharded_check'),'fwd',d['roofline']['vit_forward']['frac_of_burst_peak'],'clocks',d['clocks'])"
C=$((1024 / N))
timeout 600 $RUN bench.py --gpus $N --height 224 --width 224 --clips $C --steps 2 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; echo "configs[3] N=$N rc=$?"
tail -3 gpurun_out/bench_c3_n$N.err
python -c "
import json;d=json.load(open('gpurun_out/bench_c3_n$N.json'));print('N',d['n_gpus'],'frames/s',round(d['value']),'clips/s',round(d['clips_per_s'],1),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']),'sharded',d.get('sharded_check'))"
