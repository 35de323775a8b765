"""Oracle for K1: HF DINOv3ViTImageProcessor numerics, restated in numpy.

Follows HF:models/dinov3_vit/image_processing_dinov3_vit.py:45-86 (rescale 1/255 -> antialiased
bilinear resize -> normalise) whose resize is torchvision ``resize(antialias=True)`` -> aten
``_upsample_bilinear2d_aa`` (separable triangle filter, align_corners=False), and the caller
services/dinov3-pipeline/app/main.py:98-107 (BGR->RGB).  Test infrastructure only.
"""
from __future__ import annotations

import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def aa_weights(n_in: int, n_out: int):
    """aten HelperInterpBase::_compute_indices_min_size_weights_aa, bilinear filter, float32 math.
    Returns (lo [n_out] int, cnt [n_out] int, w [n_out, kmax] float32 normalised, zero padded)."""
    f = np.float32
    scale = f(n_in) / f(n_out)
    support = scale if scale >= 1.0 else f(1.0)
    invscale = f(1.0) / scale if scale >= 1.0 else f(1.0)
    kmax = int(np.ceil(support)) * 2 + 1
    lo = np.zeros(n_out, dtype=np.int64)
    cnt = np.zeros(n_out, dtype=np.int64)
    w = np.zeros((n_out, kmax), dtype=np.float32)
    for i in range(n_out):
        center = scale * (f(i) + f(0.5))
        xmin = max(int(center - support + f(0.5)), 0)
        xmax = min(int(center + support + f(0.5)), n_in)
        xs = xmax - xmin
        j = np.arange(xs, dtype=np.float32)
        x = np.abs((j + f(xmin) - center + f(0.5)) * invscale).astype(np.float32)
        v = np.where(x < 1.0, f(1.0) - x, f(0.0)).astype(np.float32)
        tot = v.sum(dtype=np.float32)
        if tot != 0:
            v = v / tot
        lo[i], cnt[i] = xmin, xs
        w[i, :xs] = v
    return lo, cnt, w


def resize_matrix(n_in: int, n_out: int) -> np.ndarray:
    lo, cnt, w = aa_weights(n_in, n_out)
    m = np.zeros((n_out, n_in), dtype=np.float32)
    for i in range(n_out):
        m[i, lo[i]:lo[i] + cnt[i]] = w[i, :cnt[i]]
    return m


def preprocess(frames_u8: np.ndarray, bgr: bool = True, size=(224, 224)) -> np.ndarray:
    """uint8 [n, H, W, 3] -> float32 pixel_values [n, 3, size_h, size_w] (what the HF processor returns)."""
    x = frames_u8.astype(np.float32)
    if bgr:
        x = x[..., ::-1]
    x = x * np.float32(0.00392156862745098)          # rescale (image_processing_backends.py:272-279)
    n, h, w, _ = x.shape
    rw = resize_matrix(w, size[1])                    # horizontal pass first, like aten's separable kernel
    rh = resize_matrix(h, size[0])
    x = np.einsum("nhwc,ow->nhoc", x, rw, optimize=True).astype(np.float32)
    x = np.einsum("nhoc,ph->npoc", x, rh, optimize=True).astype(np.float32)
    x = (x - MEAN) / STD
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2)).astype(np.float32)


def patchify(pixel_values: np.ndarray, patch: int = 16) -> np.ndarray:
    """[n, 3, H, W] -> [n * gh * gw, 3*patch*patch] in Conv2d-weight K order c*256 + ky*16 + kx
    (HF:modeling_dinov3_vit.py:71-81: conv stride 16 floor-divides, so only the top-left gh*16 x gw*16 is used)."""
    n, c, h, w = pixel_values.shape
    gh, gw = h // patch, w // patch
    x = pixel_values[:, :, : gh * patch, : gw * patch].reshape(n, c, gh, patch, gw, patch)
    x = x.transpose(0, 2, 4, 1, 3, 5).reshape(n * gh * gw, c * patch * patch)
    return np.ascontiguousarray(x)
