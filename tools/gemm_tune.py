"""GEMM tuning sweep on a B200: cre_gemm_bf16 over (cta_group, pipeline stages, epilogue) on the ViT-B shapes.
    python tools/gemm_tune.py > gpurun_out/gemm_tune.log
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


def main():
    model = random_init_vit(layers=1)
    eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
    dev = eng.device
    m = 300 * 201
    shapes = [(2304, 768, "qkv"), (3072, 768, "up"), (768, 3072, "down"), (768, 768, "proj")]
    epis = [(_lib.EPI_NONE, "none"), (_lib.EPI_BF16, "bf16"), (_lib.EPI_GELU, "gelu"), (_lib.EPI_RESID, "resid")]
    for n, k, name in shapes:
        a = torch.randn(m, k, device=dev).to(torch.bfloat16)
        b = torch.randn(n, k, device=dev).to(torch.bfloat16)
        bias = torch.randn(n, device=dev)
        scale = torch.ones(n, device=dev)
        out = torch.zeros(m, n, device=dev, dtype=torch.float32)
        ms = timeit(lambda: torch.matmul(a, b.t()))
        print(f"{name:5s} {m}x{n}x{k} torch.matmul: {ms * 1e3:8.1f} us {2.0 * m * n * k / ms / 1e9:7.1f} TF/s", flush=True)
        for cg, stage_list in ((1, (3, 4)), (2, (4, 5, 6))):
            for st in stage_list:
                _lib.set_tuning("gemm_stages", st)
                row = []
                for epi, ename in epis:
                    ms = timeit(lambda: eng.gemm(a, b, epi, bias=bias, scale=scale, out=out, cta_group=cg))
                    row.append(f"{ename}={2.0 * m * n * k / ms / 1e9:7.1f}")
                print(f"   cg={cg} stages={st}: " + "  ".join(row) + " TF/s", flush=True)
        _lib.set_tuning("gemm_stages", 0)


if __name__ == "__main__":
    main()
