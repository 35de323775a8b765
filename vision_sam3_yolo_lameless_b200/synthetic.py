"""Seeded random-init DINOv3 ViT weights for benchmarks and tuning tools (there is no network for ``from_pretrained``).

The reference builds its model with HuggingFace ``AutoModel`` (services/dinov3-pipeline/app/main.py:34-35); this builds the
same HF class from a config and lets HF ``_init_weights`` fill it (trunc-normal 0.02, zero bias, LayerScale 1.0,
HF:modeling_dinov3_vit.py:470-489).  Only the state dict is used: the forward pass runs in ``libcre_b200.so``.
"""
from __future__ import annotations

VIT_SHAPES = {
    # name: (hidden, mlp, layers, heads)            SURVEY.md section 8 "Model configs"
    "vitb16": (768, 3072, 12, 12),
    "vitl16": (1024, 4096, 24, 16),
}


def random_init_vit(name: str = "vitb16", seed: int = 0, layers: int | None = None):
    """HF ``DINOv3ViTModel`` (eval mode) of the named shape with seeded random-init weights."""
    import torch
    from transformers import DINOv3ViTConfig, DINOv3ViTModel

    hidden, mlp, depth, heads = VIT_SHAPES[name]
    cfg = DINOv3ViTConfig(hidden_size=hidden, intermediate_size=mlp, num_hidden_layers=layers or depth,
                          num_attention_heads=heads, num_register_tokens=4)
    torch.manual_seed(seed)
    return DINOv3ViTModel(cfg).eval()


def vit_flops_per_frame(name: str, tokens: int, patches: int) -> float:
    """Algorithmic FLOPs of one frame (2MNK per GEMM, 4 T^2 D attention per layer)."""
    hidden, mlp, depth, _ = VIT_SHAPES[name]
    return 2.0 * patches * 768 * hidden + depth * (8.0 * tokens * hidden * hidden + 4.0 * tokens * hidden * mlp
                                                   + 4.0 * tokens * tokens * hidden)
