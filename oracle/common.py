"""Seeded synthetic inputs and random-init weights shared by the oracle, the tests and bench.py
(SURVEY.md section 8(d)).  Test infrastructure only."""
from __future__ import annotations

import numpy as np
import torch


def hf_model(hidden=768, mlp=3072, layers=12, heads=12, registers=4, seed=0):
    """Seeded random-init HF DINOv3ViTModel (HF _init_weights: trunc-normal 0.02, zero bias, LayerScale 1.0)."""
    from transformers import DINOv3ViTConfig, DINOv3ViTModel

    cfg = DINOv3ViTConfig(hidden_size=hidden, intermediate_size=mlp, num_hidden_layers=layers,
                          num_attention_heads=heads, num_register_tokens=registers)
    torch.manual_seed(seed)
    return DINOv3ViTModel(cfg).eval()


def noise_frames(n, h, w, seed=0) -> np.ndarray:
    """uint8 [n, h, w, 3] uniform noise: the worst case for antialiased-resize parity."""
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


def smooth_frames(n, h, w, seed=0) -> np.ndarray:
    """uint8 [n, h, w, 3] sums of low-frequency sinusoids (image-like statistics)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h, dtype=np.float32), np.linspace(0, 1, w, dtype=np.float32), indexing="ij")
    out = np.empty((n, h, w, 3), dtype=np.uint8)
    for i in range(n):
        for c in range(3):
            acc = np.zeros((h, w), dtype=np.float32)
            for _ in range(4):
                fx, fy, ph = rng.uniform(0.5, 6.0), rng.uniform(0.5, 6.0), rng.uniform(0, 6.28)
                acc += np.sin(6.2832 * (fx * xx + fy * yy) + ph)
            out[i, :, :, c] = np.clip(127.5 + 30.0 * acc + rng.normal(0, 4.0, size=(h, w)), 0, 255).astype(np.uint8)
    return out


def cosine(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return (a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1) + 1e-30)
