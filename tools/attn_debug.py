"""Attention fast-kernel component timing (debug masks drop parts of the pipeline; results are garbage)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit
from vision_sam3_yolo_lameless_b200 import _lib
from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig
from gemm_tune import timeit

model = random_init_vit(layers=1)
eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
n, t, heads = 600, 201, 12
qkv = (torch.randn(n * t, 3 * heads * 64, device=eng.device) * 0.5).to(torch.bfloat16)
names = {0: "full", 1: "no softmax math", 2: "no PV mma", 4: "no S mma", 8: "no O read/store", 16: "no TMA loads",
         1 | 8: "no softmax, no O", 1 | 2 | 4 | 8: "barriers + TMA only", 1 | 2 | 4 | 8 | 16: "barriers only", 2 | 4: "no MMAs",
         16 | 2 | 4: "softmax + O only (no TMA, no MMA)"}
for mask, name in names.items():
    _lib.set_tuning("attention_debug", mask)
    ms = timeit(lambda: eng.attention(qkv, n, t, heads), iters=10)
    print(f"mask {mask:2d} {name:36s}: {ms * 1e3:8.1f} us", flush=True)
_lib.set_tuning("attention_debug", 0)
_lib.set_tuning("attention_fast", 0)
ms = timeit(lambda: eng.attention(qkv, n, t, heads), iters=10)
print(f"general kernel: {ms * 1e3:8.1f} us")
