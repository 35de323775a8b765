"""In-process stand-ins for the services the hot path talks to (Qdrant, NATS), with exactly the method
surface the reference uses (services/dinov3-pipeline/app/main.py:80-88,168-172,240-243;
services/tracking-service/app/reid/matcher.py:85-102,127-132,243-246,267-271,291-301;
shared/utils/nats_client.py:40-76).  Test infrastructure only.

FakeQdrant implements Distance.COSINE the way Qdrant documents it -- vectors are L2-normalised on insert,
the query is normalised, score = dot product, results in descending score -- with the tie order DEFINED
as insertion order (= GpuGallery row order).  The real server is absent (qdrant/qdrant:latest, un-pinned):
parity is unpinned at this boundary.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, List

import numpy as np


class FakeQdrant:
    def __init__(self, *args, **kwargs):
        self.collections: Dict[str, Dict[str, Any]] = {}

    # -- collection management -----------------------------------------------------------------
    def get_collections(self):
        return SimpleNamespace(collections=[SimpleNamespace(name=n) for n in self.collections])

    def create_collection(self, collection_name, vectors_config=None, **kw):
        self.collections[collection_name] = {"ids": [], "vectors": [], "payloads": [], "config": vectors_config}

    def get_collection(self, collection_name):
        return SimpleNamespace(points_count=len(self.collections[collection_name]["ids"]))

    # -- writes --------------------------------------------------------------------------------
    def upsert(self, collection_name, points, **kw):
        col = self.collections.setdefault(collection_name, {"ids": [], "vectors": [], "payloads": [], "config": None})
        for p in points:
            v = np.asarray(p.vector, dtype=np.float64)
            v = v / (np.linalg.norm(v) + 1e-30)       # COSINE collections store unit vectors
            if p.id in col["ids"]:
                r = col["ids"].index(p.id)
                col["vectors"][r], col["payloads"][r] = v, dict(p.payload or {})
            else:
                col["ids"].append(p.id)
                col["vectors"].append(v)
                col["payloads"].append(dict(p.payload or {}))

    def set_payload(self, collection_name, payload, points, **kw):
        col = self.collections[collection_name]
        for pid in points:
            col["payloads"][col["ids"].index(pid)].update(payload)

    # -- reads ---------------------------------------------------------------------------------
    def retrieve(self, collection_name, ids, with_vectors=False, **kw):
        col = self.collections[collection_name]
        out = []
        for pid in ids:
            if pid in col["ids"]:
                r = col["ids"].index(pid)
                out.append(SimpleNamespace(id=pid, payload=col["payloads"][r],
                                           vector=col["vectors"][r].tolist() if with_vectors else None))
        return out

    def scroll(self, collection_name, limit=10, with_vectors=False, with_payload=True, offset=None, **kw):
        """(points, next_page_offset) like qdrant_client: next_page_offset is None on the last page."""
        col = self.collections[collection_name]
        start = int(offset or 0)
        rows = list(zip(col["ids"], col["vectors"], col["payloads"]))
        pts = [SimpleNamespace(id=i, payload=p, vector=v.tolist() if with_vectors else None) for i, v, p in rows[start:start + limit]]
        return pts, (start + limit if start + limit < len(rows) else None)

    def _search(self, collection_name, query, limit) -> List[SimpleNamespace]:
        col = self.collections[collection_name]
        if not col["ids"]:
            return []
        q = np.asarray(query, dtype=np.float64)
        q = q / (np.linalg.norm(q) + 1e-30)
        scores = np.stack(col["vectors"]) @ q
        order = np.lexsort((np.arange(len(scores)), -scores))[:limit]
        return [SimpleNamespace(id=col["ids"][r], score=float(scores[r]), payload=col["payloads"][r]) for r in order]

    def search(self, collection_name, query_vector, limit=5, **kw):
        return self._search(collection_name, query_vector, limit)

    def query_points(self, collection_name, query, limit=5, with_payload=True, **kw):
        return SimpleNamespace(points=self._search(collection_name, query, limit))


class FakeNats:
    """NATSClient surface (shared/utils/nats_client.py): connect / publish(subject, dict) / subscribe / close."""

    def __init__(self):
        self.published: List[tuple] = []
        self.handlers: Dict[str, Any] = {}
        self.connected = False

    async def connect(self):
        self.connected = True

    async def publish(self, subject: str, data: dict):
        self.published.append((subject, data))
        cb = self.handlers.get(subject)
        if cb is not None:
            try:                       # nats_client.py:61-66: handler exceptions are printed and swallowed
                await cb(data)
            except Exception as e:     # pragma: no cover
                print(f"Error in message handler: {e}")

    async def subscribe(self, subject: str, callback, queue=None):
        self.handlers[subject] = callback

    async def close(self):
        self.connected = False
