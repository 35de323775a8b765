"""cg=1 vs cg=2 main-loop micro-diagnosis: MMA-only and TMA-only rates (EPI_NONE)."""
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
dev = eng.device
m = 300 * 201
for n, k in [(3072, 768), (768, 3072)]:
    a = torch.randn(m, k, device=dev).to(torch.bfloat16)
    b = torch.randn(n, k, device=dev).to(torch.bfloat16)
    out = torch.zeros(m, n, device=dev, dtype=torch.float32)
    for mode, name in [(0, "normal"), (1, "mma-only"), (2, "tma-only")]:
        _lib.set_tuning("gemm_debug", mode)
        for cg in (1, 2):
            ms = timeit(lambda: eng.gemm(a, b, _lib.EPI_NONE, out=out, cta_group=cg))
            print(f"{m}x{n}x{k} {name:9s} cg={cg}: {ms * 1e3:8.1f} us  {2.0 * m * n * k / ms / 1e9:7.1f} TF/s-equivalent", flush=True)
    _lib.set_tuning("gemm_debug", 0)
