#!/bin/bash
# usage: gpu_multi.sh N  -- headline config (weak scaling, configs[1] per rank) and configs[3] (1 024 clips of 224x224 sharded over N ranks)
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $RUN bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "headline N=$N rc=$?"
tail -3 gpurun_out/bench_n$N.err
python -c "
import json;d=json.load(open('gpurun_out/bench_n$N.json'));print('N',d['n_gpus'],'frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']),'clocks',d['clocks'])"
C=$((1024 / N))
timeout 600 $RUN bench.py --gpus $N --height 224 --width 224 --clips $C --steps 2 --warmup 2 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "configs[3] N=$N rc=$?"
tail -3 gpurun_out/bench_c4_n$N.err
python -c "
import json;d=json.load(open('gpurun_out/bench_c4_n$N.json'));print('N',d['n_gpus'],'frames/s',round(d['value']),'clips/s',round(d['clips_per_s'],1),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']))"
