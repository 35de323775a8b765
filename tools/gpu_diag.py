"""GPU bring-up diagnostics: each group runs in its own process (a trapped kernel poisons the CUDA
context), prints error statistics against plain PyTorch / the oracle, and never raises on mismatch.

    python tools/gpu_diag.py            # run every group, write gpurun_out/diag_<group>.log
    python tools/gpu_diag.py gemm       # run one group in-process
"""
from __future__ import annotations

import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

GROUPS = ["gemm1", "gemm2", "ln", "attn", "attn_long", "prep", "fwd2", "fwd12", "topk", "perf"]


def _engine(layers=2, max_frames=64, resize=(224, 224), hidden=768, mlp=3072, heads=12):
    import torch
    from oracle import common
    from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig

    model = common.hf_model(hidden=hidden, mlp=mlp, layers=layers, heads=heads)
    cfg = VitConfig.from_hf(model.config)
    eng = ClipEmbedEngine(cfg, model.state_dict(), max_frames=max_frames, resize=resize)
    return eng, model


def _stats(name, got, ref):
    import torch
    got = got.float().cpu()
    ref = ref.float().cpu()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    bad = (~torch.isfinite(got)).sum().item()
    print(f"  {name}: max_abs={err.max().item():.4e} mean_abs={err.mean().item():.4e} rel_max={err.max().item() / denom:.4e} "
          f"nonfinite={bad} shape={tuple(got.shape)}", flush=True)
    return err.max().item() / denom


def g_gemm(cg):
    import torch
    from vision_sam3_yolo_lameless_b200 import _lib
    eng, _ = _engine(layers=1, max_frames=8)
    dev = eng.device
    torch.manual_seed(1)
    for (m, n, k) in [(128, 256, 64), (256, 256, 128), (128, 256, 768), (1000, 768, 768), (4021, 2304, 768), (515, 768, 3072)]:
        a = (torch.randn(m, k, device=dev) * 0.5).to(torch.bfloat16)
        b = (torch.randn(n, k, device=dev) * 0.5).to(torch.bfloat16)
        bias = torch.randn(n, device=dev)
        ref = a.float() @ b.float().t() + bias
        print(f"gemm cg={cg} {m}x{n}x{k}", flush=True)
        out = eng.gemm(a, b, _lib.EPI_F32, bias=bias, cta_group=cg)
        torch.cuda.synchronize()
        _stats("f32", out, ref)
        out = eng.gemm(a, b, _lib.EPI_BF16, bias=bias, cta_group=cg)
        torch.cuda.synchronize()
        _stats("bf16", out, ref)
        out = eng.gemm(a, b, _lib.EPI_GELU, bias=bias, cta_group=cg)
        torch.cuda.synchronize()
        _stats("gelu", out, torch.nn.functional.gelu(ref))
        scale = torch.rand(n, device=dev) + 0.5
        res = torch.randn(m, n, device=dev)
        out = res.clone()
        eng.gemm(a, b, _lib.EPI_RESID, bias=bias, scale=scale, out=out, cta_group=cg)
        torch.cuda.synchronize()
        _stats("resid", out, res + scale * ref)


def g_ln():
    import torch
    eng, _ = _engine(layers=1, max_frames=8)
    dev = eng.device
    x = torch.randn(1003, 768, device=dev) * 3 + 1.5
    g = torch.randn(768, device=dev)
    b = torch.randn(768, device=dev)
    out = eng.layernorm(x, g, b)
    torch.cuda.synchronize()
    _stats("ln768", out, torch.nn.functional.layer_norm(x, (768,), g, b, 1e-5))
    x = torch.randn(77, 1024, device=dev)
    g = torch.randn(1024, device=dev)
    b = torch.randn(1024, device=dev)
    out = eng.layernorm(x, g, b)
    torch.cuda.synchronize()
    _stats("ln1024", out, torch.nn.functional.layer_norm(x, (1024,), g, b, 1e-5))


def g_attn(t, n):
    import torch
    eng, _ = _engine(layers=1, max_frames=8)
    dev = eng.device
    heads = 12
    d = heads * 64
    torch.manual_seed(2)
    q = torch.randn(n, t, heads, 64, device=dev)
    k = torch.randn(n, t, heads, 64, device=dev)
    v = torch.randn(n, t, heads, 64, device=dev)
    qs = (q * 0.125).to(torch.bfloat16)
    kb = k.to(torch.bfloat16)
    vb = v.to(torch.bfloat16)
    qk = torch.cat([qs.reshape(n * t, d), kb.reshape(n * t, d)], dim=1).contiguous()
    tpad = (t + 7) // 8 * 8
    vt = torch.zeros(n * heads * 64, tpad, device=dev, dtype=torch.bfloat16)
    vt.view(n, heads, 64, tpad)[:, :, :, :t] = vb.permute(0, 2, 3, 1)
    out = eng.attention(qk, vt, n, t, heads)
    torch.cuda.synchronize()
    att = torch.softmax(qs.float().permute(0, 2, 1, 3) @ kb.float().permute(0, 2, 3, 1), dim=-1)
    ref = (att @ vb.float().permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(n * t, d)
    _stats(f"attn t={t}", out, ref)


def g_prep():
    import numpy as np
    import torch
    from oracle import common, preprocess_ref
    eng, _ = _engine(layers=1, max_frames=8)
    for (h, w, kind) in [(1080, 1920, "noise"), (720, 1280, "smooth"), (224, 224, "noise"), (270, 482, "noise")]:
        fr = common.noise_frames(2, h, w, seed=3) if kind == "noise" else common.smooth_frames(2, h, w, seed=3)
        ref = preprocess_ref.patchify(preprocess_ref.preprocess(fr, bgr=True))
        out = eng.preprocess(torch.from_numpy(fr).to(eng.device), bgr=True)
        torch.cuda.synchronize()
        print(f"prep {h}x{w} {kind}")
        _stats("patches", out, torch.from_numpy(ref))


def g_fwd(layers):
    import numpy as np
    import torch
    from oracle import common, preprocess_ref, vit_ref
    eng, model = _engine(layers=layers, max_frames=16)
    fr = common.noise_frames(5, 224, 224, seed=4)
    pv = preprocess_ref.preprocess(fr, bgr=True)
    sd = model.state_dict()
    ref_tok = vit_ref.vit_forward(sd, torch.from_numpy(pv), heads=12, layers=layers)
    ref = ref_tok.mean(dim=1)
    for cg in (1, 2):
        from vision_sam3_yolo_lameless_b200.engine import set_cta_group
        set_cta_group(cg)
        patches = eng.preprocess(torch.from_numpy(fr).to(eng.device), bgr=True)
        emb, tok = eng.forward_patches(patches, fr.shape[0], want_tokens=True)
        torch.cuda.synchronize()
        print(f"forward layers={layers} cg={cg}")
        _stats("tokens", tok, ref_tok)
        _stats("frame_emb", emb, ref)
        print("  cosine per frame:", common.cosine(emb.cpu().numpy(), ref.numpy()), flush=True)


def g_topk():
    import numpy as np
    import torch
    from oracle import reid_ref
    eng, _ = _engine(layers=1, max_frames=8)
    dev = eng.device
    torch.manual_seed(7)
    for (q, n) in [(3, 1000), (130, 5000), (64, 100000)]:
        g = torch.nn.functional.normalize(torch.randn(n, 768, device=dev), dim=1)
        g[n // 2] = g[7]          # exact duplicate rows -> tie
        g[n - 1] = g[7]
        gb = g.to(torch.bfloat16).contiguous()
        qv = torch.nn.functional.normalize(torch.randn(q, 768, device=dev), dim=1)
        qv[0] = torch.nn.functional.normalize(g[7] + 0.05 * torch.randn(768, device=dev), dim=0)
        s, i, dump = eng.gallery_topk(qv, gb, k=5, row_base=1000, dump_scores=True)
        torch.cuda.synchronize()
        ref_scores = reid_ref.cosine_scores(qv.cpu().numpy(), gb.float().cpu().numpy())
        print(f"topk q={q} n={n}")
        _stats("scores", dump, torch.from_numpy(ref_scores))
        _, ridx = reid_ref.topk_rule(dump.cpu().numpy(), 5, row_base=1000)
        same = (ridx == i.cpu().numpy()).all()
        print(f"  indices bit-exact vs rule on GPU scores: {bool(same)}; row0 idx={i[0].tolist()} scores={s[0].tolist()}", flush=True)


def g_perf():
    import torch
    from vision_sam3_yolo_lameless_b200 import _lib
    from vision_sam3_yolo_lameless_b200.engine import set_cta_group
    eng, _ = _engine(layers=12, max_frames=256)
    dev = eng.device

    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    m = 256 * 201
    for (n, k, epi, name) in [(2304, 768, _lib.EPI_BF16, "qkv-like"), (3072, 768, _lib.EPI_GELU, "up+gelu"),
                              (768, 3072, _lib.EPI_F32, "down-like"), (768, 768, _lib.EPI_F32, "proj-like")]:
        a = torch.randn(m, k, device=dev).to(torch.bfloat16)
        b = torch.randn(n, k, device=dev).to(torch.bfloat16)
        bias = torch.randn(n, device=dev)
        dt = torch.float32 if epi == _lib.EPI_F32 else torch.bfloat16
        out = torch.empty(m, n, device=dev, dtype=dt)
        for cg in (1, 2):
            ms = timeit(lambda: eng.gemm(a, b, epi, bias=bias, out=out, cta_group=cg))
            print(f"perf gemm {name} {m}x{n}x{k} cg={cg}: {ms:.3f} ms  {2.0 * m * n * k / ms / 1e9:.1f} TFLOP/s", flush=True)
        ms = timeit(lambda: torch.matmul(a, b.t()))
        print(f"perf torch.matmul {m}x{n}x{k}: {ms:.3f} ms  {2.0 * m * n * k / ms / 1e9:.1f} TFLOP/s", flush=True)
    frames = torch.randint(0, 256, (256, 224, 224, 3), device=dev, dtype=torch.uint8)
    for cg in (1, 2):
        set_cta_group(cg)
        ms = timeit(lambda: eng.embed_frames(frames), iters=5)
        fl = eng.cfg.flops_per_frame(14, 14) * 256
        print(f"perf embed 256 frames 224x224 cg={cg}: {ms:.2f} ms  {256 / ms * 1e3:.0f} frames/s  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    big = torch.randint(0, 256, (32, 1080, 1920, 3), device=dev, dtype=torch.uint8)
    ms = timeit(lambda: eng.preprocess(big), iters=10)
    print(f"perf preprocess 32x1080p: {ms:.3f} ms  {32 * 6521856 / ms / 1e6:.1f} GB/s", flush=True)


def run_group(name):
    t0 = time.time()
    if name == "gemm1":
        g_gemm(1)
    elif name == "gemm2":
        g_gemm(2)
    elif name == "ln":
        g_ln()
    elif name == "attn":
        g_attn(201, 3)
    elif name == "attn_long":
        g_attn(1029, 1)
    elif name == "prep":
        g_prep()
    elif name == "fwd2":
        g_fwd(2)
    elif name == "fwd12":
        g_fwd(12)
    elif name == "topk":
        g_topk()
    elif name == "perf":
        g_perf()
    else:
        raise SystemExit(f"unknown group {name}")
    print(f"[{name}] done in {time.time() - t0:.1f}s", flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] != "all":
        for g in sys.argv[1:]:
            run_group(g)
        return
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    for g in GROUPS:
        log = out / f"diag_{g}.log"
        with open(log, "w") as f:
            try:
                rc = subprocess.run([sys.executable, __file__, g], stdout=f, stderr=subprocess.STDOUT, timeout=420).returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        tail = log.read_text().splitlines()[-12:]
        print(f"===== {g}: rc={rc}")
        print("\n".join(tail), flush=True)


if __name__ == "__main__":
    main()
