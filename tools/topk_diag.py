"""Where does a multi-pass top-k result differ from the oracle rule?  (bring-up diagnostics)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from oracle import reid_ref
from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig
from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit

model = random_init_vit(layers=1)
eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
dev = eng.device
for q, n, k in [(3, 40, 256), (3, 40, 48), (3, 40, 40), (3, 40, 41), (3, 300, 256), (5, 40, 16)]:
    gen = torch.Generator(device=dev).manual_seed(q * 1000 + k)
    g = torch.nn.functional.normalize(torch.randn(n, 768, device=dev, generator=gen), dim=1)
    qv = torch.nn.functional.normalize(torch.randn(q, 768, device=dev, generator=gen), dim=1)
    if n > 30:
        dup = torch.randperm(n, device=dev, generator=gen)[:12]
        g[dup] = g[dup[0]].clone()
        qv[0] = torch.nn.functional.normalize(g[dup[0]] + 0.02 * torch.randn(768, device=dev, generator=gen), dim=0)
        print("dup rows", sorted((dup + 50).tolist()))
    gb = g.to(torch.bfloat16).contiguous()
    s, i, dump = eng.gallery_topk(qv, gb, k=k, row_base=50, dump_scores=True)
    kk = min(k, n)
    ref_top, ref_idx = reid_ref.topk_rule(dump.cpu().numpy(), kk, row_base=50)
    bad = np.argwhere(i.cpu().numpy()[:, :kk] != ref_idx)
    print(f"q={q} n={n} k={k}: {len(bad)} index mismatches; first {bad[:6].tolist()}")
    if len(bad):
        r, c = bad[0]
        print("   got ", i[r, max(0, c - 3):c + 6].tolist(), [round(x, 4) for x in s[r, max(0, c - 3):c + 6].tolist()])
        print("   want", ref_idx[r, max(0, c - 3):c + 6].tolist(), [round(float(x), 4) for x in ref_top[r, max(0, c - 3):c + 6]])
        print("   full got row", i[r, :kk].tolist())
        print("   full want row", ref_idx[r].tolist())
    tail_ok = (i.cpu().numpy()[:, kk:] == 0x7FFFFFFF).all()
    print("   tail fill ok:", bool(tail_ok))
