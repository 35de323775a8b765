"""GPU-resident cosine gallery: the device mirror of a Qdrant COSINE collection.

Replaces the similarity + top-k arithmetic of the remote Qdrant calls at
services/tracking-service/app/reid/matcher.py:127-132 (query_points) and
services/dinov3-pipeline/app/main.py:168-172 (search); upserts (matcher.py:243-246,291-301,
dinov3 main.py:240-243) keep going to Qdrant as the durable store when a client is attached
("write-through"), and also land in the device matrix so that the next search sees them.

Vectors are stored L2-normalised in bf16 [capacity, D] (Qdrant normalises COSINE vectors on insert).
Row order = insertion order; search ties break on the lower row index.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .engine import ClipEmbedEngine


class ScoredPoint:
    """Shape of a qdrant_client ScoredPoint as the reference consumes it (.id, .score, .payload)."""

    __slots__ = ("id", "score", "payload", "vector")

    def __init__(self, id, score, payload, vector=None):
        self.id, self.score, self.payload, self.vector = id, score, payload, vector

    def __repr__(self):
        return f"ScoredPoint(id={self.id!r}, score={self.score:.6f})"


class GpuGallery:
    def __init__(self, engine: ClipEmbedEngine, dim: int, capacity: int = 4096):
        self.engine = engine
        self.dim = int(dim)
        self.capacity = int(capacity)
        self.matrix = torch.zeros((self.capacity, self.dim), dtype=torch.bfloat16, device=engine.device)
        self.ids: List[Any] = []
        self.payloads: List[Dict[str, Any]] = []
        self._row_of: Dict[Any, int] = {}

    def __len__(self) -> int:
        return len(self.ids)

    def _grow(self, need: int) -> None:
        if need <= self.capacity:
            return
        cap = max(need, 2 * self.capacity)
        m = torch.zeros((cap, self.dim), dtype=torch.bfloat16, device=self.engine.device)
        m[: len(self.ids)] = self.matrix[: len(self.ids)]
        self.matrix, self.capacity = m, cap

    def load(self, ids: List[Any], unit_vectors: torch.Tensor, payloads: Optional[List[Dict[str, Any]]] = None) -> None:
        """Bulk load already-normalised vectors [N, D] (any float dtype, host or device)."""
        n = len(ids)
        self._grow(n)
        self.matrix[:n] = unit_vectors.to(device=self.engine.device, dtype=torch.bfloat16)
        self.ids = list(ids)
        self.payloads = list(payloads) if payloads is not None else [{} for _ in ids]
        self._row_of = {pid: r for r, pid in enumerate(self.ids)}

    def upsert(self, point_id: Any, vector, payload: Optional[Dict[str, Any]] = None, momentum: float = 0.0) -> int:
        """Insert or overwrite one point; the row is normalised on the device (cre_gallery_update_row).
        momentum > 0 blends with the stored row: row <- norm(momentum * row + (1 - momentum) * unit(vector))."""
        v = torch.as_tensor(np.asarray(vector, dtype=np.float32)).to(self.engine.device).reshape(1, -1)
        _, v = self.engine.pool_clips(v, torch.tensor([0, 1], dtype=torch.int32))  # unit(vector), matcher.py:274
        row = self._row_of.get(point_id)
        if row is None:
            row = len(self.ids)
            self._grow(row + 1)
            self.ids.append(point_id)
            self.payloads.append(dict(payload or {}))
            self._row_of[point_id] = row
            momentum = 0.0
        elif payload is not None:
            self.payloads[row] = dict(payload)
        self.engine.gallery_update_row(self.matrix, row, v, momentum)
        return row

    def vector(self, point_id: Any) -> Optional[np.ndarray]:
        row = self._row_of.get(point_id)
        return None if row is None else self.matrix[row].float().cpu().numpy()

    def search(self, query, k: int = 5) -> List[ScoredPoint]:
        """One query -> up to k ScoredPoints in descending cosine order (Qdrant normalises the query too)."""
        return self.search_batch(np.asarray(query, dtype=np.float32)[None, :], k)[0]

    def search_batch(self, queries, k: int = 5) -> List[List[ScoredPoint]]:
        q = torch.as_tensor(np.asarray(queries, dtype=np.float32)).to(self.engine.device)
        n = len(self.ids)
        if n == 0:
            return [[] for _ in range(q.shape[0])]
        offs = torch.arange(q.shape[0] + 1, dtype=torch.int32)
        _, unit = self.engine.pool_clips(q, offs)  # one-frame "clips": L2 normalisation with the +1e-8 rule
        scores, idx = self.engine.gallery_topk(unit, self.matrix[:n], k=min(k, 8))
        scores, idx = scores.cpu().numpy(), idx.cpu().numpy()
        out = []
        for r in range(q.shape[0]):
            hits = []
            for s, i in zip(scores[r], idx[r]):
                if i >= n or not np.isfinite(s):
                    continue
                hits.append(ScoredPoint(self.ids[i], float(s), self.payloads[i]))
            out.append(hits)
        return out
