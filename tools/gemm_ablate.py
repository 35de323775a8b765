"""Epilogue / prefetch ablation of the tcgen05 GEMM at the bench's launch size (M = 1130 frames x 201 tokens).
gemm_debug bits: 4 = accumulators never read, 8 = TMEM read only, 16 = staged but never stored, 32 = L2 prefetch of A.
    python tools/gemm_ablate.py > gpurun_out/gemm_ablate.log
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit
from vision_sam3_yolo_lameless_b200 import _lib
from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig


def timeit(fn, iters=20):
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


def resid_variants():
    """RESID (TMA reduce-add, separate LayerNorm) vs RESID_LN / RESID_LN3 (x through the epilogue, LayerNorm folded)."""
    model = random_init_vit(layers=1)
    eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
    dev = eng.device
    m, n = 1130 * 201, 768
    x = torch.randn(m, n, device=dev)
    _, stats = eng.row_stats(x)
    bias, scale = torch.randn(n, device=dev), torch.ones(n, device=dev) * 1e-3
    g, b = torch.ones(n, device=dev), torch.zeros(n, device=dev)
    for k, name in ((3072, "down"), (768, "proj")):
        a = torch.randn(m, k, device=dev).to(torch.bfloat16)
        w = torch.randn(n, k, device=dev).to(torch.bfloat16)
        fl = 2.0 * m * n * k
        for rep in range(2):
            row = []
            ms = timeit(lambda: eng.gemm(a, w, _lib.EPI_RESID, bias=bias, scale=scale, out=x, cta_group=2))
            ms_ln = timeit(lambda: eng.layernorm(x, g, b))
            row.append(f"resid={fl / ms / 1e9:6.0f} ({ms * 1e3:6.1f} us) + layernorm {ms_ln * 1e3:6.1f} us")
            for e, label in ((_lib.EPI_RESID_LN, "resid_ln"), (_lib.EPI_RESID_LN3, "resid_ln3")):
                ms = timeit(lambda: eng.gemm_ln(a, w, e, stats, n, bias=bias, scale=scale, out=x, cta_group=2))
                row.append(f"{label}={fl / ms / 1e9:6.0f} ({ms * 1e3:6.1f} us)")
            print(f"{name:5s} {m}x{n}x{k}: " + "  ".join(row), flush=True)


def main():
    if "--resid" in sys.argv:
        return resid_variants()
    model = random_init_vit(layers=1)
    eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
    dev = eng.device
    m = 1130 * 201
    shapes = [(2304, 768, "qkv", _lib.EPI_BF16, "bf16"), (3072, 768, "up", _lib.EPI_GELU, "gelu"),
              (768, 3072, "down", _lib.EPI_RESID, "resid"), (768, 768, "proj", _lib.EPI_RESID, "resid")]
    for n, k, name, epi, ename in shapes:
        a = torch.randn(m, k, device=dev).to(torch.bfloat16)
        b = torch.randn(n, k, device=dev).to(torch.bfloat16)
        bias = torch.randn(n, device=dev)
        scale = torch.ones(n, device=dev)
        out = torch.zeros(m, n, device=dev, dtype=torch.float32)
        variants = [(_lib.EPI_NONE, 0, "none"), (epi, 0, ename), (epi, 32, ename + "+pf")]
        if epi != _lib.EPI_RESID:
            variants += [(epi, 4, ename + " no-ld"), (epi, 8, ename + " ld-only"), (epi, 16, ename + " no-store")]
        else:
            variants += [(_lib.EPI_NONE, 32, "none+pf"), (_lib.EPI_F32, 0, "f32-store")]
        for rep in range(2):
            ms = timeit(lambda: torch.matmul(a, b.t()))
            row = [f"cublas={2.0 * m * n * k / ms / 1e9:6.0f}"]
            for e, dbg, label in variants:
                _lib.set_tuning("gemm_debug", dbg)
                ms = timeit(lambda: eng.gemm(a, b, e, bias=bias, scale=scale, out=out, cta_group=2))
                row.append(f"{label}={2.0 * m * n * k / ms / 1e9:6.0f}")
            _lib.set_tuning("gemm_debug", 0)
            print(f"{name:5s} {m}x{n}x{k}: " + "  ".join(row), flush=True)


if __name__ == "__main__":
    main()
