"""Launches K3 and both forms of K4 once each at the bench's sizes (target of the ncu capture in tools/gpu_round.sh)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig
from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit

model = random_init_vit(layers=1)
eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
dev = eng.device
g = torch.Generator(device=dev).manual_seed(3)
gal = torch.nn.functional.normalize(torch.randn(100_000, 768, device=dev, generator=g), dim=1).to(torch.bfloat16)
emb = torch.randn(64 * 150, 768, device=dev, generator=g)
offs = torch.arange(0, 64 * 150 + 1, 150, dtype=torch.int32, device=dev)
for rep in range(2):                      # second round = warm instruction caches; the gallery (154 MB) never fits L2 (126 MB)
    _, unit = eng.pool_clips(emb, offs)                       # K3: 64 clips x 150 frames
    eng.gallery_topk(unit[:1].contiguous(), gal, k=5)         # K4 serving form (Q = 1): gallery_scan_small_kernel
    eng.gallery_topk(unit, gal, k=5)                          # K4 batched form (Q = 64): topk_prepare + gemm_tn_kernel<6,1,3> + merge_topk
torch.cuda.synchronize()
print("ok")
