"""Attention fast kernel, one timing (600 frames x 12 heads, T = 201)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit
from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig
from gemm_tune import timeit

model = random_init_vit(layers=1)
eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
n, t, heads = 600, 201, 12
qkv = (torch.randn(n * t, 3 * heads * 64, device=eng.device) * 0.5).to(torch.bfloat16)
best = min(timeit(lambda: eng.attention(qkv, n, t, heads), iters=20) for _ in range(3))
print(f"{sys.argv[1] if len(sys.argv) > 1 else ''} attention 600 x 12 x 201: {best * 1e3:7.1f} us", flush=True)
