"""Attention at T = 201 (the 224 x 224 shape): error and time of every kernel variant, each in its own process (a trapped kernel
poisons the CUDA context).

    python tools/attn_variants.py                 # all variants -> stdout
    python tools/attn_variants.py one <name>      # one variant in-process
"""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

VARIANTS = {
    "single_S": {"attention_split": 0},
    "split_default": {},
    "split_delay0": {"attention_split_delay": 0},
    "split_delay2400": {"attention_split_delay": 2400},
    "split_delay4000": {"attention_split_delay": 4000},
    "split_poly0": {"attention_poly": 0},
    "split_poly50": {"attention_poly": 2},
    "split_pv1": {"attention_split_mode": 2},
    "long_kernel": {"attention_fast": 0},        # the T > 256 persistent kernel on T = 201 (three 96-key blocks, two 128-row tiles)
}


def one(name):
    import torch
    from gemm_tune import timeit
    from vision_sam3_yolo_lameless_b200 import _lib
    from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig
    from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit

    model = random_init_vit(layers=1)
    eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
    for k, v in VARIANTS[name].items():
        _lib.set_tuning(k, v)
    t, heads = 201, 12
    d = heads * 64
    gen = torch.Generator(device=eng.device).manual_seed(5)
    for n in (3, 70):
        q, k, v = (torch.randn(n, t, heads, 64, device=eng.device, generator=gen) for _ in range(3))
        qs, kb, vb = (q * 0.125).to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)
        qkv = torch.cat([qs.reshape(n * t, d), kb.reshape(n * t, d), vb.reshape(n * t, d)], dim=1).contiguous()
        out = eng.attention(qkv, n, t, heads)
        torch.cuda.synchronize()
        att = torch.softmax(qs.float().permute(0, 2, 1, 3) @ kb.float().permute(0, 2, 3, 1), dim=-1)
        ref = (att @ vb.float().permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(n * t, d)
        err = (out.float() - ref).abs()
        per_unit = err.reshape(n, t, heads, 64).amax(dim=(1, 3))
        print(f"{name:22s} n={n:3d}: rel_err {err.max().item() / ref.abs().max().item():.3e}  nonfinite {(~torch.isfinite(out.float())).sum().item()}"
              f"  worst units {[tuple(int(x) for x in divmod(int(i), heads)) for i in per_unit.flatten().topk(3).indices]}", flush=True)
    for n in (600, 1130):
        qkv = (torch.randn(n * t, 3 * d, device=eng.device) * 0.5).to(torch.bfloat16)
        best = min(timeit(lambda: eng.attention(qkv, n, t, heads), iters=20) for _ in range(3))
        fl = 4.0 * t * t * 64 * heads * n
        gb = n * t * (3 * d + d) * 2 / 1e9
        print(f"{name:22s} n={n:4d}: {best * 1e3:7.1f} us  {fl / best / 1e9:6.1f} TFLOP/s  {gb / best * 1e3:6.0f} GB/s (q,k,v read + out write)", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "one":
        one(sys.argv[2])
    else:
        names = sys.argv[2].split(",") if len(sys.argv) > 2 and sys.argv[1] == "only" else list(VARIANTS)
        for name in names:
            r = subprocess.run([sys.executable, __file__, "one", name], capture_output=True, text=True, timeout=240)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                print(f"{name}: rc={r.returncode}\n{r.stderr[-1500:]}", flush=True)
