"""Attention kernels over T: the kernel the launcher picks (split-S for 160 < T <= 208, single-S persistent for other T <= 256,
long-sequence persistent above) against the long-sequence kernel forced for every T (attention_fast = 0).
    python tools/attn_t_sweep.py > gpurun_out/attn_t_sweep.log"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import torch
from gemm_tune import timeit
from vision_sam3_yolo_lameless_b200 import _lib
from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig
from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit

model = random_init_vit(layers=1)
eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
heads, d = 12, 768
for t in [int(x) for x in sys.argv[1:]] or (65, 129, 149, 160, 201, 230, 256, 261, 401, 785, 1029):
    n = max(8, 230000 // t)
    qkv = (torch.randn(n * t, 3 * d, device=eng.device) * 0.5).to(torch.bfloat16)
    row = []
    for fast in (1, 0):
        _lib.set_tuning("attention_fast", fast)
        ms = min(timeit(lambda: eng.attention(qkv, n, t, heads), iters=10) for _ in range(2))
        row.append(f"{'picked' if fast else 'long  '} {ms * 1e3:8.1f} us {4.0 * t * t * 64 * heads * n / ms / 1e9:6.1f} TFLOP/s")
    _lib.set_tuning("attention_fast", 1)
    print(f"T={t:4d} n={n:5d}:  " + "   |   ".join(row), flush=True)
