"""kNN similarity graph over clip embeddings on the GPU: the second consumer of the cosine top-k kernel (K4).

Mirrors ``GraphBuilder.compute_knn_edges`` of the reference GNN service
(services/gnn-pipeline/app/main.py:55-100; the graph-transformer service builds the same graph): L2-normalise (+1e-8),
cosine similarity of every node against every node, self excluded, the k most similar nodes per node emitted in ASCENDING
similarity order as ``edge_index [2, E]`` / ``edge_weights [E]``.  The reference is an O(N^2) numpy matrix + a per-row
``argsort``; here it is ``cre_gallery_topk`` with Q = N queries against the N-row bf16 unit matrix, asking for
k + 1 hits and dropping the node itself.  Ties (unspecified in the reference) follow (score desc, index asc).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .engine import ClipEmbedEngine


class GraphBuilder:
    def __init__(self, engine: ClipEmbedEngine, k_neighbors: int = 5, embedding_dim: int = 64):
        if engine is None:
            raise RuntimeError("GraphBuilder needs a ClipEmbedEngine (GPU kernels); there is no CPU fallback")
        self.engine = engine
        self.k_neighbors = k_neighbors
        self.embedding_dim = embedding_dim

    def compute_knn_edges(self, embeddings: np.ndarray, k: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
        if k is None:
            k = self.k_neighbors
        n = len(embeddings)
        if n <= k:
            k = max(1, n - 1)
        if k + 1 > _lib.TOPK_LIMIT:
            raise ValueError(f"k={k} exceeds the kernel's top-k limit ({_lib.TOPK_LIMIT - 1} neighbours)")
        if n < 2:
            return np.zeros((2, 0), dtype=np.int64), np.zeros((0,), dtype=np.float64)
        eng = self.engine
        x = torch.as_tensor(np.asarray(embeddings, dtype=np.float32)).to(eng.device)
        _, unit = eng.pool_clips(x, torch.arange(n + 1, dtype=torch.int32))     # one-row "clips": e / (||e|| + 1e-8)
        gallery = unit.to(torch.bfloat16).contiguous()
        scores, idx = eng.gallery_topk(unit, gallery, k=k + 1)
        # Drop the node itself from its k + 1 hits (if a duplicate row outranked it on the index tie rule: the weakest hit)
        # and reverse to ascending similarity like argsort()[-k:] -- index bookkeeping on the device, no per-node host loop.
        idx = idx.to(torch.int64)
        node = torch.arange(n, device=idx.device)
        is_self = idx == node[:, None]
        drop = torch.where(is_self.any(dim=1), is_self.to(torch.int8).argmax(dim=1), torch.full_like(node, k))
        keep = torch.ones_like(is_self)
        keep[node, drop] = False
        dst = idx[keep].view(n, k).flip(1)
        w = scores[keep].view(n, k).flip(1).to(torch.float64)
        src = node.repeat_interleave(k)
        return torch.stack([src, dst.reshape(-1)]).cpu().numpy(), w.reshape(-1).cpu().numpy()
