"""K1 standalone timing on 1080p / 720p noise frames (device resident)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit
from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig
from gemm_tune import timeit

model = random_init_vit(layers=1)
eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=600)
from vision_sam3_yolo_lameless_b200 import _lib
modes = [int(a) for a in sys.argv[1:]] or [2, 1, 0]    # preprocess_tma values: 2 = TMA-staged row pairs (default), 1 = TMA-staged, 0 = direct-load kernel
import os
SIZES = [(600, 1080, 1920), (600, 720, 1280), (600, 224, 224)]
if os.environ.get("PREP_SIZES"):          # e.g. PREP_SIZES=540x960,480x640
    SIZES = [(600, int(a), int(b)) for a, b in (x.split("x") for x in os.environ["PREP_SIZES"].split(","))]
for (n, h, w) in SIZES:
    fr = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=eng.device)
    for mode in modes:
        _lib.set_tuning("preprocess_tma", mode)
        ms = timeit(lambda: eng.preprocess(fr), iters=10)
        b = n * (h * w * 3 + 196 * 1536)
        print(f"preprocess[mode {mode}] {n}x{h}x{w}: {ms:.3f} ms  {ms / n * 1e3:.2f} us/frame  {b / ms / 1e6:.0f} GB/s", flush=True)
_lib.set_tuning("preprocess_tma", 2)
