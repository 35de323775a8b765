"""GPU-resident cosine gallery: the device mirror of a Qdrant COSINE collection.

Replaces the similarity + top-k arithmetic of the remote Qdrant calls at
services/tracking-service/app/reid/matcher.py:127-132 (query_points) and
services/dinov3-pipeline/app/main.py:168-172 (search); upserts (matcher.py:243-246,291-301,
dinov3 main.py:240-243) keep going to Qdrant as the durable store when a client is attached
("write-through"), and also land in the device matrices so that the next search sees them.

Two device copies of every vector, both L2-normalised (Qdrant normalises COSINE vectors on insert):

* ``master`` f32 [capacity, D] -- the vector at the precision Qdrant keeps it.  The reference reads it back
  (``retrieve(with_vectors=True)``, matcher.py:267-271), blends the new embedding into it and upserts the result
  (:281-301): momentum updates therefore read and write THIS copy, and ``vector()`` / the Qdrant write-through return it;
* ``matrix`` bf16 [capacity, D] -- what the scan kernel (K4) streams: always a rounding of the master row, never a source.

Row order = insertion order; search ties break on the lower row index.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .engine import ClipEmbedEngine


class ScoredPoint:
    """Shape of a qdrant_client ScoredPoint as the reference consumes it (.id, .score, .payload)."""

    __slots__ = ("id", "score", "payload", "vector")

    def __init__(self, id, score, payload, vector=None):
        self.id, self.score, self.payload, self.vector = id, score, payload, vector

    def __repr__(self):
        return f"ScoredPoint(id={self.id!r}, score={self.score:.6f})"


def scroll_all(qdrant_client, collection_name: str, page: int = 1024) -> Iterable[Any]:
    """Every point of a collection with its vector, paged (``scroll`` returns (points, next_page_offset))."""
    offset = None
    while True:
        points, offset = qdrant_client.scroll(collection_name=collection_name, limit=page, offset=offset,
                                              with_vectors=True, with_payload=True)
        yield from points
        if offset is None or not points:
            return


class GpuGallery:
    def __init__(self, engine: ClipEmbedEngine, dim: int, capacity: int = 4096):
        self.engine = engine
        self.dim = int(dim)
        self.capacity = int(capacity)
        self.matrix = torch.zeros((self.capacity, self.dim), dtype=torch.bfloat16, device=engine.device)
        self.master = torch.zeros((self.capacity, self.dim), dtype=torch.float32, device=engine.device)
        self.ids: List[Any] = []
        self.payloads: List[Dict[str, Any]] = []
        self._row_of: Dict[Any, int] = {}

    def __len__(self) -> int:
        return len(self.ids)

    def _grow(self, need: int) -> None:
        if need <= self.capacity:
            return
        cap = max(need, 2 * self.capacity)
        n = len(self.ids)
        for name, dt in (("matrix", torch.bfloat16), ("master", torch.float32)):
            m = torch.zeros((cap, self.dim), dtype=dt, device=self.engine.device)
            m[:n] = getattr(self, name)[:n]
            setattr(self, name, m)
        self.capacity = cap

    def _unit(self, vectors) -> torch.Tensor:
        """[n, D] (host or device, any float dtype) -> device f32 unit rows, e / (||e|| + 1e-8) (K3b: one-row "clips")."""
        v = vectors if isinstance(vectors, torch.Tensor) else torch.as_tensor(np.asarray(vectors, dtype=np.float32))
        v = v.to(device=self.engine.device, dtype=torch.float32).reshape(-1, self.dim)
        _, unit = self.engine.pool_clips(v, torch.arange(v.shape[0] + 1, dtype=torch.int32))
        return unit

    def load(self, ids: List[Any], vectors, payloads: Optional[List[Dict[str, Any]]] = None) -> None:
        """Replace the contents with n vectors [n, D] in ONE normalisation launch (rows need not be normalised)."""
        n = len(ids)
        self._grow(n)
        if n:
            unit = self._unit(vectors)
            self.master[:n] = unit
            self.matrix[:n] = unit              # dtype cast on copy: round-to-nearest-even bf16
        self.ids = list(ids)
        self.payloads = [dict(p or {}) for p in payloads] if payloads is not None else [{} for _ in ids]
        self._row_of = {pid: r for r, pid in enumerate(self.ids)}

    def load_from_qdrant(self, qdrant_client, collection_name: str, page: int = 1024) -> int:
        """Mirror an existing collection (service start / restart): paged scroll, one bulk load.  Returns the point count."""
        ids, vecs, payloads = [], [], []
        for p in scroll_all(qdrant_client, collection_name, page):
            ids.append(p.id)
            vecs.append(np.asarray(p.vector, dtype=np.float32))
            payloads.append(dict(p.payload or {}))
        self.load(ids, np.stack(vecs) if vecs else np.zeros((0, self.dim), np.float32), payloads)
        return len(ids)

    def upsert(self, point_id: Any, vector, payload: Optional[Dict[str, Any]] = None, momentum: float = 0.0) -> int:
        """Insert or overwrite one point; the row is normalised on the device (cre_gallery_update_row).
        momentum > 0 blends with the stored MASTER row: row <- norm(momentum * row + (1 - momentum) * unit(vector))."""
        unit = self._unit(np.asarray(vector, dtype=np.float32)[None, :])          # unit(vector), matcher.py:274
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
        self.engine.gallery_update_row(self.matrix, row, unit, momentum, master=self.master)
        return row

    def vector(self, point_id: Any) -> Optional[np.ndarray]:
        """The stored unit vector at full (f32) precision -- what Qdrant's retrieve(with_vectors=True) returns."""
        row = self._row_of.get(point_id)
        return None if row is None else self.master[row].cpu().numpy()

    def search(self, query, k: int = 5) -> List[ScoredPoint]:
        """One query -> up to k ScoredPoints in descending cosine order (Qdrant normalises the query too)."""
        return self.search_batch(np.asarray(query, dtype=np.float32)[None, :], k)[0]

    def search_batch(self, queries, k: int = 5) -> List[List[ScoredPoint]]:
        k = int(k)
        if k < 1:
            raise ValueError(f"top_k={k}: must be >= 1")
        if k > _lib.TOPK_LIMIT:
            raise ValueError(f"top_k={k} exceeds the kernel limit ({_lib.TOPK_LIMIT}); the request is refused rather than truncated")
        q = np.asarray(queries, dtype=np.float32)
        n = len(self.ids)
        if n == 0:
            return [[] for _ in range(q.shape[0])]
        unit = self._unit(q)                     # L2 normalisation with the +1e-8 rule
        scores, idx = self.engine.gallery_topk(unit, self.matrix[:n], k=min(k, n))
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


class ShardedGpuGallery(GpuGallery):
    """The same gallery ROW-SHARDED over the ranks of a process group (one process per GPU): SURVEY.md section 8(e) behind the
    matcher / extractor interface instead of only inside bench.py.

    Contract (SPMD, like every torch.distributed program): every rank constructs it and calls the SAME methods with the SAME
    arguments in the SAME order -- e.g. each tracking-service replica consumes the same ``pipeline.dinov3`` subject.  Then

    * host metadata (ids, payloads, global row numbers = insertion order) is replicated and identical everywhere;
    * global row r lives on rank ``r % world`` at local row ``r // world`` (round-robin: ownership does not move when the gallery
      grows, unlike a contiguous split of a growing N); a write -- create or momentum update (matcher.py:203-301) -- runs its kernel
      on the OWNING rank only, in message order, and ``vector()`` broadcasts the owner's fp32 master row;
    * ``search`` / ``search_batch`` are collectives: every rank scans its shard (K4, candidates carry GLOBAL row numbers), the
      candidates are all-gathered and merged under (score desc, global row asc) -- ``ShardedReID.search_all`` -- so every rank returns
      the same hits a single-device gallery would.
    """

    def __init__(self, engine: ClipEmbedEngine, dim: int, capacity: int = 4096, group=None, sharded_reid=None):
        import torch.distributed as dist

        from .sharded import ShardedReID

        super().__init__(engine, dim, capacity)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        # per-shard scan with global row numbers; the merge is the engine's (or whatever the injected ShardedReID uses)
        self._reid = sharded_reid if sharded_reid is not None else ShardedReID(engine, group=group, local_topk=self._local_topk,
                                                                                  merge=lambda s, i: engine.merge_topk(s, i))

    # global row <-> (owner, local row)
    def _owner(self, row: int) -> int:
        return row % self.world

    def _local_rows(self) -> int:
        n = len(self.ids)
        return (n - self.rank + self.world - 1) // self.world if n > self.rank else 0

    def _local_topk(self, unit_queries: torch.Tensor, k: int):
        n_local = self._local_rows()
        s, i = self.engine.gallery_topk(unit_queries, self.matrix[:n_local], k=k)
        valid = i != 0x7FFFFFFF
        return s, torch.where(valid, i * self.world + self.rank, i)           # local row -> global row

    def load(self, ids, vectors, payloads=None) -> None:
        n = len(ids)
        mine = list(range(self.rank, n, self.world))
        self._grow((n + self.world - 1) // self.world + 1)
        if mine:
            v = vectors if isinstance(vectors, torch.Tensor) else torch.as_tensor(np.asarray(vectors, dtype=np.float32))
            unit = self._unit(v[mine])
            self.master[: len(mine)] = unit
            self.matrix[: len(mine)] = unit
        self.ids = list(ids)
        self.payloads = [dict(p or {}) for p in payloads] if payloads is not None else [{} for _ in ids]
        self._row_of = {pid: r for r, pid in enumerate(self.ids)}

    def upsert(self, point_id, vector, payload=None, momentum: float = 0.0) -> int:
        row = self._row_of.get(point_id)
        if row is None:
            row = len(self.ids)
            self.ids.append(point_id)
            self.payloads.append(dict(payload or {}))
            self._row_of[point_id] = row
            momentum = 0.0
        elif payload is not None:
            self.payloads[row] = dict(payload)
        if self._owner(row) == self.rank:                    # the write itself: owning rank only
            local = row // self.world
            self._grow(local + 1)
            unit = self._unit(np.asarray(vector, dtype=np.float32)[None, :])
            self.engine.gallery_update_row(self.matrix, local, unit, momentum, master=self.master)
        return row

    def vector(self, point_id):
        import torch.distributed as dist

        row = self._row_of.get(point_id)
        if row is None:
            return None
        owner = self._owner(row)
        buf = self.master[row // self.world].clone() if owner == self.rank else torch.empty(self.dim, dtype=torch.float32,
                                                                                            device=self.engine.device)
        if self.world > 1:
            src = dist.get_global_rank(self.group, owner) if self.group is not None else owner
            dist.broadcast(buf, src=src, group=self.group)
        return buf.cpu().numpy()

    def search_batch(self, queries, k: int = 5):
        k = int(k)
        if k < 1:
            raise ValueError(f"top_k={k}: must be >= 1")
        if k > _lib.TOPK_LIMIT:
            raise ValueError(f"top_k={k} exceeds the kernel limit ({_lib.TOPK_LIMIT}); the request is refused rather than truncated")
        q = np.asarray(queries, dtype=np.float32)
        n = len(self.ids)
        if n == 0:
            return [[] for _ in range(q.shape[0])]
        unit = self._unit(q)
        scores, idx = self._reid.search_all(unit, k=min(k, n))         # every rank holds the same queries: no query all-gather
        scores, idx = scores.cpu().numpy(), idx.cpu().numpy()
        return [[ScoredPoint(self.ids[i], float(s), self.payloads[i]) for s, i in zip(scores[r], idx[r]) if i < n and np.isfinite(s)]
                for r in range(q.shape[0])]
