#!/bin/bash
# headline config (configs[1] per rank, weak scaling) on N GPUs of one box
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
timeout 900 $RUN bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "headline N=$N rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_n$N.json'));print('N',d['n_gpus'],'frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']),'h2d',d['e2e'].get('h2d_only_gbs'),'sharded',d.get('sharded_check'),'fwd',d['roofline']['vit_forward']['frac_of_burst_peak'],'clocks',d['clocks'])"
