"""K3 / K4 standalone timing: clip pooling and the gallery scan (+ merge) at the bench's sizes and at serving sizes (Q = 1)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from vision_sam3_yolo_lameless_b200 import _lib
from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig
from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit
from gemm_tune import timeit

model = random_init_vit(layers=1)
eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
dev = eng.device
g = torch.Generator(device=dev).manual_seed(3)
for rows in (100_000, 12_500):
    gal = torch.nn.functional.normalize(torch.randn(rows, 768, device=dev, generator=g), dim=1).to(torch.bfloat16)
    for q in (1, 2, 64, 128, 1024):
        qs = torch.nn.functional.normalize(torch.randn(q, 768, device=dev, generator=g), dim=1)
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
        scan = min(per["gemm_topk"])
        print(f"gallery {rows:6d} x 768, Q={q:4d}: call {ms_all * 1e3:7.1f} us   {parts}   scan {rows * 1536 / scan / 1e6:6.0f} GB/s", flush=True)
for clips, fpc in ((64, 150), (1024, 150), (1, 5)):
    emb = torch.randn(clips * fpc, 768, device=dev, generator=g)
    offs = torch.arange(0, clips * fpc + 1, fpc, dtype=torch.int32, device=dev)
    ms = timeit(lambda: eng.pool_clips(emb, offs), iters=20)
    print(f"pool_clips {clips} x {fpc} frames: {ms * 1e3:7.1f} us  {emb.numel() * 4 / ms / 1e6:6.0f} GB/s", flush=True)
