"""Residual-update GEMMs at the bench's launch size (1130 frames x 201 tokens): the fp32-stream epilogues (EPI_RESID_LN / LN3) against
the split-stream ones (EPI_RESID_SP / SP3) on both shapes of a block (attention-out K = 768, MLP-down K = 3072).
    python tools/resid_bench.py > gpurun_out/resid_bench.log
Reports time per launch, TFLOP/s and the HBM rate of the algorithmic bytes (A + residual stream in and out)."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from vision_sam3_yolo_lameless_b200 import _lib
from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig
from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit


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
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, nargs="*", default=[768, 3072])
    ap.add_argument("--epis", nargs="*", default=["ln", "ln3", "sp", "sp3"])
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    model = random_init_vit(layers=1)
    eng = ClipEmbedEngine(VitConfig.from_hf(model.config), model.state_dict(), max_frames=8)
    dev = eng.device
    m, n = 1130 * 201, 768
    x0 = torch.randn(m, n, device=dev)
    hi, lo, stats = eng.row_stats_split(x0)
    bias, scale = torch.randn(n, device=dev) * 0.1, torch.rand(n, device=dev) * 0.1
    for k in args.k:
        name = {768: "attention-out", 3072: "mlp-down"}.get(k, f"K={k}")
        a = (torch.randn(m, k, device=dev) * 0.1).to(torch.bfloat16)
        w = (torch.randn(n, k, device=dev) * 0.05).to(torch.bfloat16)
        flops = 2.0 * m * n * k
        for epi, ename, bytes_per_el in ((_lib.EPI_RESID_LN, "RESID_LN ", 10), (_lib.EPI_RESID_LN3, "RESID_LN3", 10),
                                         (_lib.EPI_RESID_SP, "RESID_SP ", 8), (_lib.EPI_RESID_SP3, "RESID_SP3", 8)):
            if ename.strip().lower()[6:] not in args.epis:
                continue
            x = x0.clone()
            h2, l2 = hi.clone(), lo.clone()
            out = (h2, l2) if epi in (_lib.EPI_RESID_SP, _lib.EPI_RESID_SP3) else x
            ms = timeit(lambda: eng.gemm_ln(a, w, epi, stats, n, bias=bias, scale=scale, out=out, cta_group=2), iters=args.iters)
            traffic = m * k * 2 + m * n * bytes_per_el
            print(f"{name:13s} {ename} {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s  {traffic / ms / 1e6:7.0f} GB/s "
                  f"({traffic / 1e9:.2f} GB algorithmic)", flush=True)


if __name__ == "__main__":
    main()
