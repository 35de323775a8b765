"""Oracle for K2/K3a: fp32 restatement of HF DINOv3ViTModel.forward + token mean, written against a
plain state_dict with torch CPU ops (no HF modules).  Test infrastructure only.

Follows HF:models/dinov3_vit/modeling_dinov3_vit.py -- embeddings :75-92, RoPE tables :95-121,153-200,
rotate_half / apply_rotary_pos_emb :203-207,238-268, attention :294-334 (scaling head_dim^-0.5, no mask),
LayerScale :337-343, MLP :381-386 (exact-erf GELU), layer :424-450, final norm :546-548 -- and
services/dinov3-pipeline/app/main.py:113 (mean over ALL tokens).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def rope_tables(gh: int, gw: int, head_dim: int = 64, theta: float = 100.0):
    ch = torch.arange(0.5, gh, dtype=torch.float32) / gh
    cw = torch.arange(0.5, gw, dtype=torch.float32) / gw
    coords = torch.stack(torch.meshgrid(ch, cw, indexing="ij"), dim=-1).flatten(0, 1)
    coords = 2.0 * coords - 1.0
    inv_freq = 1 / theta ** torch.arange(0, 1, 4 / head_dim, dtype=torch.float32)
    angles = 2 * math.pi * coords[:, :, None] * inv_freq[None, None, :]
    angles = angles.flatten(1, 2).tile(2)
    return torch.cos(angles), torch.sin(angles)


def _rot_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def _key(sd, i, name):
    for k in (f"model.layer.{i}.{name}", f"layer.{i}.{name}"):
        if k in sd:
            return sd[k]
    return None


@torch.no_grad()
def vit_forward(sd, pixel_values: torch.Tensor, heads: int, layers: int, patch: int = 16, eps: float = 1e-5,
                theta: float = 100.0) -> torch.Tensor:
    """pixel_values f32 [n, 3, H, W] -> last_hidden_state f32 [n, T, D] (after the final LayerNorm)."""
    x = pixel_values.to(torch.float32)
    n = x.shape[0]
    w = sd["embeddings.patch_embeddings.weight"].float()
    d = w.shape[0]
    pe = F.conv2d(x, w, sd["embeddings.patch_embeddings.bias"].float(), stride=patch)
    gh, gw = pe.shape[2], pe.shape[3]
    pe = pe.flatten(2).transpose(1, 2)
    h = torch.cat([sd["embeddings.cls_token"].float().expand(n, -1, -1),
                   sd["embeddings.register_tokens"].float().expand(n, -1, -1), pe], dim=1)
    t = h.shape[1]
    prefix = t - gh * gw
    cos, sin = rope_tables(gh, gw, d // heads, theta)
    for i in range(layers):
        g = lambda name: _key(sd, i, name)
        y = F.layer_norm(h, (d,), g("norm1.weight").float(), g("norm1.bias").float(), eps)
        q = F.linear(y, g("attention.q_proj.weight").float(), g("attention.q_proj.bias"))
        k = F.linear(y, g("attention.k_proj.weight").float(), g("attention.k_proj.bias"))
        v = F.linear(y, g("attention.v_proj.weight").float(), g("attention.v_proj.bias"))
        q = q.view(n, t, heads, -1).transpose(1, 2)
        k = k.view(n, t, heads, -1).transpose(1, 2)
        v = v.view(n, t, heads, -1).transpose(1, 2)
        qp, kp = q[:, :, prefix:], k[:, :, prefix:]
        q = torch.cat([q[:, :, :prefix], qp * cos + _rot_half(qp) * sin], dim=2)
        k = torch.cat([k[:, :, :prefix], kp * cos + _rot_half(kp) * sin], dim=2)
        att = torch.softmax((q @ k.transpose(2, 3)) * (q.shape[-1] ** -0.5), dim=-1)
        o = (att @ v).transpose(1, 2).reshape(n, t, d)
        o = F.linear(o, g("attention.o_proj.weight").float(), g("attention.o_proj.bias"))
        h = h + o * g("layer_scale1.lambda1").float()
        y = F.layer_norm(h, (d,), g("norm2.weight").float(), g("norm2.bias").float(), eps)
        y = F.linear(F.gelu(F.linear(y, g("mlp.up_proj.weight").float(), g("mlp.up_proj.bias"))),
                     g("mlp.down_proj.weight").float(), g("mlp.down_proj.bias"))
        h = h + y * g("layer_scale2.lambda1").float()
    return F.layer_norm(h, (d,), sd["norm.weight"].float(), sd["norm.bias"].float(), eps)


@torch.no_grad()
def frame_embeddings(sd, pixel_values, heads, layers, batch: int = 16, **kw) -> torch.Tensor:
    """services/dinov3-pipeline/app/main.py:113: last_hidden_state.mean(dim=1) -> f32 [n, D]."""
    outs = []
    for s in range(0, pixel_values.shape[0], batch):
        outs.append(vit_forward(sd, pixel_values[s:s + batch], heads, layers, **kw).mean(dim=1))
    return torch.cat(outs, dim=0)
