"""Golden transcript for the re-ID write path and for decisions NEAR the thresholds: the REFERENCE'S OWN ``CowReIDMatcher``
(services/tracking-service/app/reid/matcher.py, imported unmodified from /root/reference, its Qdrant client replaced by
oracle.fake_services.FakeQdrant) driven through a scripted sequence whose queries are built, step by step, at a chosen cosine
against what the reference HAS STORED at that moment -- 2e-3 on either side of 0.65 / 0.75 / 0.85 (matcher.py:52-54), i.e. ten
times the score error of a bf16 scan copy -- and, after every step, the vector the reference keeps in the durable store
(matcher.py:228-246 create, :281-301 momentum update).  Run in the authoring container only; outputs are committed.

    python -m oracle.make_golden_reid      ->  tests/golden/reid_tight.npz
"""
from __future__ import annotations

import asyncio
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import fake_services, reference_loader  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
DIM = 768
# (name, identity the query is aimed at (index into creation order, None = fresh direction), cosine against its STORED vector)
SCRIPT = [("create_a", None, None), ("create_b", None, None),
          ("a_just_high", 0, 0.852), ("a_just_below_high", 0, 0.848),
          ("b_just_medium", 1, 0.752), ("b_just_below_medium", 1, 0.748),      # 0.748 < 0.75 -> low -> creates COW-0003
          ("a_just_low", 0, 0.652), ("a_just_below_low", 0, 0.648),            # both create (0.65 <= s < 0.75 and s < 0.65)
          ("a_after_updates", 0, 0.93), ("b_after_update", 1, 0.99), ("a_again", 0, 0.751), ("a_final", 0, 0.8505)]


def main():
    assert reference_loader.available(), "/root/reference is required to generate golden vectors"
    rm = reference_loader.load_matcher_module()
    fq = fake_services.FakeQdrant()
    rm.QdrantClient = lambda url=None, **kw: fq
    matcher = rm.CowReIDMatcher(qdrant_url="fake://")
    asyncio.run(matcher.connect())
    rng = np.random.default_rng(97)
    created = []            # identity ids in creation order
    queries, stored_after, steps = [], [], []
    for name, target, cos in SCRIPT:
        if target is None:
            q = rng.standard_normal(DIM) * rng.uniform(0.5, 4.0)
        else:
            col = fq.collections[rm.CowReIDMatcher.COLLECTION_NAME]
            g = np.asarray(col["vectors"][col["ids"].index(created[target])], dtype=np.float64)
            n = rng.standard_normal(DIM)
            n -= (n @ g) * g
            n /= np.linalg.norm(n)
            q = (cos * g + np.sqrt(1.0 - cos * cos) * n) * rng.uniform(0.5, 4.0)
        m = matcher.match_or_create(np.asarray(q, dtype=np.float64), video_id=f"video-{name}", track_id=len(steps))
        if m.is_new_identity:
            created.append(str(m.identity_id))
        col = fq.collections[rm.CowReIDMatcher.COLLECTION_NAME]
        row = col["ids"].index(str(m.identity_id))
        queries.append(q)
        stored_after.append(np.asarray(col["vectors"][row], dtype=np.float64))
        steps.append({"name": name, "cow_id": m.cow_id, "similarity": float(m.similarity), "confidence": m.confidence,
                      "is_new": bool(m.is_new_identity), "total_sightings": int(col["payloads"][row]["total_sightings"])})
        print(steps[-1])
    np.savez_compressed(GOLDEN / "reid_tight.npz", queries=np.stack(queries), stored_after=np.stack(stored_after),
                        transcript=json.dumps(steps))


if __name__ == "__main__":
    main()
