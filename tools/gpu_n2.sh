#!/bin/bash
# 2-GPU check of the final code: the sharded matcher test over NCCL + the headline bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "sharded_matcher_on_two_gpus" > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_n2.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $RUN bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "headline N=2 rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_n2.json'));print('N',d['n_gpus'],'frames/s',round(d['value']),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']),'sharded',d.get('sharded_check'),'fwd',d['roofline']['vit_forward']['frac_of_burst_peak'],'clocks',d['clocks'])"
