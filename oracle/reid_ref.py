"""Oracle for K3b/K4: clip pooling, L2 normalisation, cosine scores and deterministic top-k (numpy).
Test infrastructure only.

* clip mean            services/dinov3-pipeline/app/main.py:204-208 (np.mean over frame embeddings, float64)
* canonical-frame mean services/tracking-service/app/main.py:294-302
* L2 normalise         services/tracking-service/app/reid/matcher.py:124  e / (||e|| + 1e-8)
* cosine + top-k       matcher.py:127-132 / dinov3 main.py:168-172 -> Qdrant Distance.COSINE (remote server,
                       un-pinned): restated as unit(q) . unit(g), descending; ties (unspecified by the
                       reference) are DEFINED as (score desc, index asc).  Parity unpinned at this boundary.
* thresholds           matcher.py:52-54,303-311
"""
from __future__ import annotations

import numpy as np

SIMILARITY_THRESHOLD_HIGH = 0.85    # matcher.py:52
SIMILARITY_THRESHOLD_MEDIUM = 0.75  # matcher.py:53
SIMILARITY_THRESHOLD_LOW = 0.65     # matcher.py:54


def clip_mean(frame_emb: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    return np.stack([np.mean(frame_emb[offsets[c]:offsets[c + 1]].astype(np.float64), axis=0)
                     for c in range(len(offsets) - 1)]).astype(np.float64)


def l2_normalise(e: np.ndarray) -> np.ndarray:
    return e / (np.linalg.norm(e, axis=-1, keepdims=True) + 1e-8)


def cosine_scores(unit_q: np.ndarray, gallery: np.ndarray) -> np.ndarray:
    """fp32 queries x gallery (as stored, e.g. bf16 values up-cast) accumulated in float64 -> float32."""
    return (unit_q.astype(np.float64) @ gallery.astype(np.float64).T).astype(np.float32)


def topk_rule(scores: np.ndarray, k: int, row_base: int = 0):
    """The defined total order applied to a score matrix [Q, N]: (score desc, index asc)."""
    q, n = scores.shape
    idx = np.empty((q, min(k, n)), dtype=np.int64)
    for r in range(q):
        order = np.lexsort((np.arange(n), -scores[r].astype(np.float64)))
        idx[r] = order[:k]
    top = np.take_along_axis(scores, idx, axis=1)
    return top, (idx + row_base).astype(np.int32)


def merge_rule(scores: np.ndarray, idx: np.ndarray, k: int):
    """Merge [lists, Q, k] candidate lists with the same order (cross-shard merge)."""
    lists, q, kk = scores.shape
    s = scores.transpose(1, 0, 2).reshape(q, lists * kk)
    i = idx.transpose(1, 0, 2).reshape(q, lists * kk)
    out_s = np.empty((q, k), dtype=np.float32)
    out_i = np.empty((q, k), dtype=np.int32)
    for r in range(q):
        order = np.lexsort((i[r], -s[r].astype(np.float64)))[:k]
        out_s[r], out_i[r] = s[r][order], i[r][order]
    return out_s, out_i


def score_to_confidence(score: float) -> str:
    if score >= SIMILARITY_THRESHOLD_HIGH:
        return "high"
    if score >= SIMILARITY_THRESHOLD_MEDIUM:
        return "medium"
    if score >= SIMILARITY_THRESHOLD_LOW:
        return "low"
    return "none"


def momentum_update(old_unit: np.ndarray, new_emb: np.ndarray, momentum: float = 0.9) -> np.ndarray:
    """matcher.py:257-301"""
    new_unit = new_emb / (np.linalg.norm(new_emb) + 1e-8)
    upd = momentum * old_unit + (1 - momentum) * new_unit
    return upd / (np.linalg.norm(upd) + 1e-8)


class MatcherOracle:
    """numpy restatement of CowReIDMatcher's read/decide/write cycle (matcher.py:104-301) over an in-memory
    COSINE gallery: match_embedding -> threshold decision -> momentum update or create.  Returns dicts shaped
    like tests/golden/reid_scenario.json steps."""

    def __init__(self, momentum: float = 0.9, auto_create: bool = True, store_dtype=np.float64):
        self.momentum, self.auto_create = momentum, auto_create
        self.vectors, self.cow_ids = [], []
        self.store_dtype = store_dtype

    def match_embedding(self, e: np.ndarray, top_k: int = 5):
        if not self.vectors:
            return None, []
        q = l2_normalise(np.asarray(e, dtype=np.float64))
        g = np.stack(self.vectors).astype(np.float64)
        s = g @ q
        order = np.lexsort((np.arange(len(s)), -s))[:top_k]
        cands = [{"cow_id": self.cow_ids[r], "similarity": float(s[r]), "confidence": score_to_confidence(s[r]), "row": int(r)}
                 for r in order]
        best = cands[0] if cands[0]["similarity"] >= SIMILARITY_THRESHOLD_LOW else None
        return best, cands

    def match_or_create(self, e: np.ndarray):
        best, cands = self.match_embedding(e)
        if best is not None and best["similarity"] >= SIMILARITY_THRESHOLD_MEDIUM:
            r = best["row"]
            self.vectors[r] = momentum_update(self.vectors[r].astype(np.float64), np.asarray(e, np.float64),
                                              self.momentum).astype(self.store_dtype)
            return {"cow_id": best["cow_id"], "similarity": best["similarity"], "confidence": best["confidence"], "is_new": False}
        if self.auto_create:
            self.vectors.append(l2_normalise(np.asarray(e, dtype=np.float64)).astype(self.store_dtype))
            self.cow_ids.append(f"COW-{len(self.cow_ids) + 1:04d}")
            return {"cow_id": self.cow_ids[-1], "similarity": 1.0, "confidence": "high", "is_new": True}
        return {"cow_id": "UNKNOWN", "similarity": cands[0]["similarity"] if cands else 0.0, "confidence": "low", "is_new": True}


def neighbor_evidence(similar_cases) -> float:
    """services/dinov3-pipeline/app/main.py:216-225"""
    if not similar_cases:
        return 0.5
    labels = [c["label"] for c in similar_cases if c["label"] is not None]
    if not labels:
        return 0.5
    return sum(1 for l in labels if l == 1) / len(labels)
