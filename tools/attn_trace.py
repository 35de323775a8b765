"""Timeline of the split-S attention kernel's warps on one SM (CTA 0), from a -DCRE_ATTN_TRACE build of the library:

    for f in api gemm attention elementwise preprocess profile; do nvcc <Makefile flags> -DCRE_ATTN_TRACE -c csrc/$f.cu -o /tmp/$f.o; done
    nvcc -shared -o tools/libcre_b200_trace.so /tmp/*.o
    CRE_B200_LIB=tools/libcre_b200_trace.so python tools/attn_trace.py

Prints, per unit, the cycle offsets of the events of softmax warps 4 (group 0) / 8 (group 1) and of the two MMA issuers."""
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("CRE_B200_LIB", str(ROOT / "tools" / "libcre_b200_trace.so"))
import torch

from vision_sam3_yolo_lameless_b200 import _lib
from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig
from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit

NAMES = {1: "I.wait_free", 2: "I.got_free", 3: "I.S0_issued", 4: "I.got_P0", 5: "I.PV0_issued", 6: "I.got_P1b0", 7: "I.got_P1b1",
         8: "I.got_P1b2", 9: "I.commit_O", 10: "I.S1next", 11: "I.pre_wait", 12: "I.post_wait", 13: "I.mma", 20: "S.wait_S0", 21: "S.got_S0", 22: "S.h0_done", 23: "S.got_S1",
         24: "S.h1_done", 25: "S.got_O", 26: "S.O_read", 27: "S.stage_free", 28: "S.unit_done"}

model = random_init_vit(layers=1)
eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
lib = _lib.load()
n, t, heads = 600, 201, 12
qkv = (torch.randn(n * t, 3 * heads * 64, device=eng.device) * 0.5).to(torch.bfloat16)
for _ in range(3):
    eng.attention(qkv, n, t, heads)
buf = torch.zeros(12 * 256, dtype=torch.int64, device=eng.device)
fn = lib.cre_debug_set_attention_trace
fn.argtypes, fn.restype = [C.c_void_p], C.c_int32
fn(buf.data_ptr())
eng.attention(qkv, n, t, heads)
torch.cuda.synchronize()
fn(None)
tr = buf.cpu().view(12, 256)
t0 = min(int(v) & ((1 << 56) - 1) for v in tr.flatten().tolist() if v != 0)
def events(warp):
    return [((int(v) >> 56) & 0xff, (int(v) & ((1 << 56) - 1)) - t0) for v in tr[warp].tolist() if v != 0]


# one table per unit: every softmax warp's milestones next to the issuers'
soft = {w: events(w) for w in range(4, 11)}
iss = {w: events(w) for w in (3, 11)}
for unit in (5, 6):
    print(f"=== unit {unit} (cycles since kernel start)")
    for w, ev in soft.items():
        per = [ev[i:i + 9] for i in range(0, len(ev), 9)]
        if unit < len(per):
            print(f"  softmax warp {w:2d} (g{(w - 4) // 4} q{w % 4}): " + "  ".join(f"{NAMES[e][2:]}@{c}" for e, c in per[unit]))
    for w, ev in iss.items():
        starts = [i for i, (e, _) in enumerate(ev) if e == 1]
        if unit < len(starts):
            seg = ev[starts[unit]:(starts[unit + 1] if unit + 1 < len(starts) else None)]
            print(f"  issuer  warp {w:2d} (g{0 if w == 3 else 1}):    " + "  ".join(f"{NAMES[e][2:]}@{c}" for e, c in seg))
