"""Generates tests/golden/* by running the REFERENCE'S OWN code, imported unmodified from /root/reference
(oracle/reference_loader.py), on seeded synthetic inputs with seeded random-init weights.  Run in the
authoring container only (the GPU box has no /root/reference); the outputs are committed.

    python -m oracle.make_golden

Fixtures (inputs are regenerated from seeds by the tests, only reference OUTPUTS are stored):
  preprocess.npz       HF DINOv3ViTImageProcessor pixel_values for 4 frame shapes (what main.py:107 feeds the model)
  embed_vitb.npz       DINOv3Pipeline.extract_embedding (main.py:95-115) on 5 frames, ViT-B/16 seed 0
  embed_vitl.npz       same, ViT-L/16 seed 0, 2 frames
  clip_48x64_15fps.avi + video.json   extract_video_embeddings (main.py:117-163): sampled indices, times, embeddings
  process_video.json   process_video (main.py:188-282) results JSON + published NATS payloads, 3 videos, fake Qdrant/NATS
  reid_scenario.json   CowReIDMatcher.match_or_create transcript (matcher.py:151-201) over a scripted query sequence
  knn_graph.json       GraphBuilder.compute_knn_edges (gnn-pipeline/app/main.py:55-100) edge list on 42 clustered nodes
"""
from __future__ import annotations

import asyncio
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import common, fake_services, reference_loader  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

# (name, n, h, w, kind, seed) -- tests rebuild these with oracle.common generators
PREPROCESS_CASES = [("noise_1080p", 1, 1080, 1920, "noise", 21), ("smooth_720p", 1, 720, 1280, "smooth", 22),
                    ("noise_224", 1, 224, 224, "noise", 23), ("noise_270x482", 1, 270, 482, "noise", 24)]
EMBED_CASES = [("noise_224_a", 224, 224, "noise", 31), ("noise_224_b", 224, 224, "noise", 32),
               ("noise_270x482", 270, 482, "noise", 33), ("smooth_720p", 720, 1280, "smooth", 34),
               ("noise_1080p", 1080, 1920, "noise", 35)]


def frames_for(kind, n, h, w, seed):
    return common.noise_frames(n, h, w, seed) if kind == "noise" else common.smooth_frames(n, h, w, seed)


def reid_queries(dim=768, seed=41):
    """Scripted query sequence: every branch of match_or_create with wide margins around the thresholds."""
    rng = np.random.default_rng(seed)
    base = [rng.standard_normal(dim) for _ in range(3)]

    def mix(b, cos_target):
        n = rng.standard_normal(dim)
        n -= (n @ b) / (b @ b) * b
        n *= np.linalg.norm(b) / np.linalg.norm(n)
        return cos_target * b + np.sqrt(1 - cos_target ** 2) * n

    seq = [("new_a", base[0] * 3.0),                 # empty gallery -> create COW-0001
           ("same_a", mix(base[0], 0.95) * 0.5),     # high -> momentum update
           ("new_b", base[1]),                       # unrelated -> create COW-0002
           ("medium_a", mix(base[0], 0.80)),         # medium -> update
           ("low_b", mix(base[1], 0.70)),            # low (< 0.75) -> create COW-0003
           ("none_c", base[2] * 10.0),               # none -> create COW-0004
           ("again_a", mix(base[0], 0.90)),          # high after two momentum updates
           ("again_b", mix(base[1], 0.97))]
    return seq


def knn_embeddings(seed=61):
    """42 nodes in 6 well-separated clusters (neighbour sets are unambiguous under bf16 gallery rounding)."""
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((6, 768))
    return np.concatenate([c + 0.35 * rng.standard_normal((7, 768)) for c in centers])


def write_clip(path: Path, n=45, h=48, w=64, fps=15, seed=51):
    import cv2

    fr = common.smooth_frames(n, h, w, seed)
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"MJPG"), fps, (w, h))
    assert vw.isOpened()
    for f in fr:
        vw.write(f)
    vw.release()


def main():
    assert reference_loader.available(), "/root/reference is required to generate golden vectors"
    from transformers import DINOv3ViTImageProcessor

    GOLDEN.mkdir(parents=True, exist_ok=True)
    proc = DINOv3ViTImageProcessor()

    # ---- preprocess: exactly what main.py:98-107 computes before the model -----------------------------
    import cv2
    from PIL import Image

    out = {}
    for name, n, h, w, kind, seed in PREPROCESS_CASES:
        fr = frames_for(kind, n, h, w, seed)
        pil = Image.fromarray(cv2.cvtColor(fr[0], cv2.COLOR_BGR2RGB))
        out[name] = proc(images=pil, return_tensors="pt")["pixel_values"][0].numpy().astype(np.float32)
    np.savez_compressed(GOLDEN / "preprocess.npz", **out)

    # ---- extract_embedding, ViT-B/16 and ViT-L/16 ------------------------------------------------------
    model_b = common.hf_model()
    pipe = reference_loader.make_reference_pipeline(model_b, proc)
    embs = {}
    for name, h, w, kind, seed in EMBED_CASES:
        embs[name] = pipe.extract_embedding(frames_for(kind, 1, h, w, seed)[0]).astype(np.float32)
    embs["gray_224"] = pipe.extract_embedding(common.noise_frames(1, 224, 224, 36)[0, :, :, 0]).astype(np.float32)
    np.savez_compressed(GOLDEN / "embed_vitb.npz", **embs)

    model_l = common.hf_model(hidden=1024, mlp=4096, layers=24, heads=16)
    pipe_l = reference_loader.make_reference_pipeline(model_l, proc)
    embs = {name: pipe_l.extract_embedding(frames_for(kind, 1, h, w, seed)[0]).astype(np.float32)
            for name, h, w, kind, seed in EMBED_CASES[:2]}
    np.savez_compressed(GOLDEN / "embed_vitl.npz", **embs)
    del model_l, pipe_l

    # ---- extract_video_embeddings + process_video ------------------------------------------------------
    clip = GOLDEN / "clip_48x64_15fps.avi"
    write_clip(clip)
    data = pipe.extract_video_embeddings(clip)
    json.dump({"fps": data["fps"], "total_frames": data["total_frames"],
               "frames": [e["frame"] for e in data["embeddings"]], "times": [e["time"] for e in data["embeddings"]],
               "canonical": [e["frame"] for e in data["canonical_frames"]],
               "embeddings": [e["embedding"] for e in data["embeddings"]]}, open(GOLDEN / "video.json", "w"))

    qd, nats = fake_services.FakeQdrant(), fake_services.FakeNats()
    pipe.qdrant_client, pipe.nats_client, pipe.collection_name = qd, nats, "cow_embeddings"
    pipe.config = {"nats": {"subjects": {"pipeline_dinov3": "pipeline.dinov3", "video_preprocessed": "video.preprocessed"}}}
    qd.create_collection("cow_embeddings")
    transcript = []
    with tempfile.TemporaryDirectory() as td:
        pipe.results_dir = Path(td)
        clips = []
        for i, seed in enumerate((51, 52, 53)):
            p = Path(td) / f"v{i}.avi"
            write_clip(p, seed=seed)
            clips.append(p)
        for i, p in enumerate(clips):
            asyncio.run(pipe.process_video({"video_id": f"vid-{i}", "processed_path": str(p), "filename": p.name,
                                            "metadata": {"n": i}}))
            res = json.load(open(Path(td) / f"vid-{i}_dinov3.json"))
            subject, msg = nats.published[-1]
            msg = dict(msg)
            msg["results_path"] = Path(msg["results_path"]).name
            transcript.append({"results": res, "subject": subject, "message": msg})
            qd.set_payload("cow_embeddings", {"label": i % 2}, [f"vid-{i}"])   # "labelled later" -> neighbour evidence
        asyncio.run(pipe.process_video({"video_id": "missing", "processed_path": str(Path(td) / "nope.avi")}))
        assert len(nats.published) == 3
    json.dump({"clip_seeds": [51, 52, 53], "transcript": transcript}, open(GOLDEN / "process_video.json", "w"))

    # ---- re-ID matcher transcript ----------------------------------------------------------------------
    rm = reference_loader.load_matcher_module()
    fq = fake_services.FakeQdrant()
    rm.QdrantClient = lambda url=None, **kw: fq
    matcher = rm.CowReIDMatcher(qdrant_url="fake://")
    asyncio.run(matcher.connect())
    steps = []
    for name, q in reid_queries():
        m = matcher.match_or_create(np.asarray(q, dtype=np.float64), video_id=f"video-{name}", track_id=len(steps))
        steps.append({"name": name, "cow_id": m.cow_id, "similarity": float(m.similarity), "confidence": m.confidence,
                      "is_new": bool(m.is_new_identity)})
    best, cands = matcher.match_embedding(np.asarray(reid_queries()[0][1]))
    json.dump({"steps": steps, "final_candidates": [{"cow_id": c.cow_id, "similarity": float(c.similarity),
                                                     "confidence": c.confidence} for c in cands],
               "thresholds": [rm.CowReIDMatcher.SIMILARITY_THRESHOLD_HIGH, rm.CowReIDMatcher.SIMILARITY_THRESHOLD_MEDIUM,
                              rm.CowReIDMatcher.SIMILARITY_THRESHOLD_LOW],
               "confidence_probe": {str(s): matcher._score_to_confidence(s) for s in
                                    (0.0, 0.6499, 0.65, 0.7499, 0.75, 0.8499, 0.85, 1.0)},
               "statistics": matcher.get_statistics()}, open(GOLDEN / "reid_scenario.json", "w"), indent=1)
    # ---- kNN similarity graph (gnn-pipeline GraphBuilder.compute_knn_edges, extracted by AST and run unmodified) ----
    from oracle import knn_ref
    gb = knn_ref.reference_graph_builder()
    emb = knn_embeddings()
    ei, ew = gb(k_neighbors=5).compute_knn_edges(emb)
    ei2, ew2 = gb(k_neighbors=5).compute_knn_edges(emb[:4])       # N <= k: k shrinks to N - 1
    json.dump({"seed": 61, "edge_index": ei.tolist(), "edge_weights": ew.tolist(), "small_edge_index": ei2.tolist(),
               "small_edge_weights": ew2.tolist()}, open(GOLDEN / "knn_graph.json", "w"))
    for f in sorted(GOLDEN.iterdir()):
        print(f"{f.name}: {f.stat().st_size} bytes")


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
