"""Golden vectors for the per-track crop path: the REFERENCE'S OWN extract_embedding
(services/dinov3-pipeline/app/main.py:95-115, imported unmodified from /root/reference) applied to the numpy crop
frame[y0:y1, x0:x1] -- what "per-track embeddings" means for the tracking service (tracking main.py:332-334) -- plus the HF
processor's pixel_values for two of the crops.  Run in the authoring container only; outputs are committed.

    python -m oracle.make_golden_roi      ->  tests/golden/roi_crops.npz
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import common, reference_loader  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
ROI_FRAMES = ("smooth", 2, 720, 1280, 71)       # kind, n, h, w, seed (oracle.common generators)
# (name, frame, x0, y0, x1, y1): a large box, the whole frame, odd alignment, the right / bottom edges, an upscaled small box
ROI_CASES = [("box_large", 0, 100, 50, 500, 400), ("whole", 0, 0, 0, 1280, 720), ("odd", 1, 601, 333, 777, 700),
             ("edge", 1, 1000, 500, 1280, 720), ("small", 0, 37, 411, 150, 500)]


def main():
    assert reference_loader.available(), "/root/reference is required to generate golden vectors"
    import cv2
    from PIL import Image
    from transformers import DINOv3ViTImageProcessor

    proc = DINOv3ViTImageProcessor()
    pipe = reference_loader.make_reference_pipeline(common.hf_model(), proc)
    kind, n, h, w, seed = ROI_FRAMES
    fr = common.smooth_frames(n, h, w, seed)
    out = {}
    for name, f, x0, y0, x1, y1 in ROI_CASES:
        crop = np.ascontiguousarray(fr[f, y0:y1, x0:x1])
        out["emb_" + name] = pipe.extract_embedding(crop).astype(np.float32)
        if name in ("odd", "small"):
            pil = Image.fromarray(cv2.cvtColor(crop, cv2.COLOR_BGR2RGB))
            out["pix_" + name] = proc(images=pil, return_tensors="pt")["pixel_values"][0].numpy().astype(np.float16)
    np.savez_compressed(GOLDEN / "roi_crops.npz", **out)
    print("roi_crops.npz:", (GOLDEN / "roi_crops.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
