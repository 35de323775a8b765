"""CPU restatement of the reference's per-frame embedding call and decode loop, built directly on the same
third-party classes the reference instantiates (HF ``DINOv3ViTImageProcessor`` + ``DINOv3ViTModel``, cv2, PIL).
Used as (a) the ``cpu_baseline`` / ``--impl reference`` leg of bench.py on the GPU box, where /root/reference
does not exist, and (b) a checker in tests.  Pinned against the reference's own outputs in tests/golden/
(tests/test_oracle_golden.py).  Test / measurement infrastructure only -- never imported by the product.

  extract_embedding          services/dinov3-pipeline/app/main.py:95-115
  sampled_frame_indices      services/dinov3-pipeline/app/main.py:123-146 (fps truncation, 1 frame / second)
  extract_video_embeddings   services/dinov3-pipeline/app/main.py:117-163
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch


class ReferencePipelineCPU:
    def __init__(self, model, processor=None, device="cpu"):
        from transformers import DINOv3ViTImageProcessor

        self.device = torch.device(device)
        self.processor = processor if processor is not None else DINOv3ViTImageProcessor()
        self.model = model.to(self.device).eval()

    def extract_embedding(self, image: np.ndarray) -> np.ndarray:
        import cv2
        from PIL import Image

        rgb = cv2.cvtColor(image, cv2.COLOR_BGR2RGB) if image.ndim == 3 and image.shape[2] == 3 else image
        inputs = self.processor(images=Image.fromarray(rgb), return_tensors="pt").to(self.device)
        with torch.no_grad():
            out = self.model(**inputs)
        return out.last_hidden_state.mean(dim=1).squeeze().cpu().numpy()

    def extract_video_embeddings(self, video_path: Path):
        import cv2

        cap = cv2.VideoCapture(str(video_path))
        if not cap.isOpened():
            raise Exception(f"Failed to open video: {video_path}")
        fps = int(cap.get(cv2.CAP_PROP_FPS))
        total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        interval = max(1, fps)
        embs, count = [], 0
        while True:
            ret, frame = cap.read()
            if not ret:
                break
            if count % interval == 0:
                embs.append({"frame": count, "time": count / fps if fps > 0 else 0,
                             "embedding": self.extract_embedding(frame).tolist()})
            count += 1
        cap.release()
        canon = [embs[0], embs[len(embs) // 2], embs[-1]] if embs else []
        return {"embeddings": embs, "canonical_frames": canon, "total_frames": total, "fps": fps}


def sampled_frame_indices(total_frames: int, fps: float):
    interval = max(1, int(fps))
    return [i for i in range(total_frames) if i % interval == 0]
