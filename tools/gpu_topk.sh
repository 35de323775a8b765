#!/bin/bash
# K4: parity of every gallery / merge / kNN test, then standalone timing (stacked hi / lo form on and off)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "gallery or topk or knn or matcher or sharded_topk or smoke or reid" > gpurun_out/pytest_topk.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_topk.log
timeout 300 python - > gpurun_out/topk_bench_r02d.log 2>&1 <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import torch
from vision_sam3_yolo_lameless_b200 import _lib
from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig
from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit
from gemm_tune import timeit
model = random_init_vit(layers=1)
eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
dev = eng.device
g = torch.Generator(device=dev).manual_seed(3)
gal = torch.nn.functional.normalize(torch.randn(100_000, 768, device=dev, generator=g), dim=1).to(torch.bfloat16)
for q in (3, 16, 64, 65, 128):
    qs = torch.nn.functional.normalize(torch.randn(q, 768, device=dev, generator=g), dim=1)
    for st in (1, 0):
        _lib.set_tuning("topk_stacked", st)
        eng.gallery_topk(qs, gal, k=5)
        _lib.profile_start(4096)
        for _ in range(5):
            eng.gallery_topk(qs, gal, k=5)
        recs = _lib.profile_stop(4096)
        per = {}
        for name, ms, work in recs:
            per.setdefault(name, []).append(ms)
        ms_all = timeit(lambda: eng.gallery_topk(qs, gal, k=5), iters=20)
        parts = "  ".join(f"{n}={min(v) * 1e3:6.1f}us" for n, v in per.items())
        print(f"Q={q:4d} stacked={st}: call {ms_all * 1e3:7.1f} us   {parts}   scan {100000 * 1536 / min(per['gemm_topk']) / 1e6:6.0f} GB/s", flush=True)
    _lib.set_tuning("topk_stacked", 1)
PY
cat gpurun_out/topk_bench_r02d.log
