"""End-to-end parity on a B200 through the public API: uint8 frames -> clip embeddings -> re-ID, against the
fp32 oracle and against the reference's own outputs in tests/golden/.  Gate (BASELINE.json): per-clip embedding
cosine >= 0.999 vs fp32; top-k indices bit-exact on the same scores."""
import asyncio
import json

import numpy as np
import pytest
import torch

from oracle import common, fake_services, preprocess_ref, reid_ref, vit_ref
from oracle.make_golden import EMBED_CASES, frames_for, reid_queries, write_clip

from conftest import engine_cached

pytestmark = pytest.mark.gpu

COS_GATE = 0.999          # BASELINE.json north_star tolerance
SUBJECTS = {"nats": {"subjects": {"pipeline_dinov3": "pipeline.dinov3", "video_preprocessed": "video.preprocessed"}},
            "qdrant": {"collection_name": "cow_embeddings"}}


@pytest.mark.parametrize("ln_fold,resid_split", [(1, 1), (1, 0), (0, 0)])
@pytest.mark.parametrize("cg", [1, 2])
def test_forward_tokens_vs_fp32_oracle(engine_b, cg, ln_fold, resid_split):
    """All three block formulations (LayerNorm folded into the GEMMs with the residual stream as two bf16 halves = default; folded
    with an fp32 stream; separate LayerNorm launches) and both tile shapes."""
    from vision_sam3_yolo_lameless_b200 import _lib
    from vision_sam3_yolo_lameless_b200.engine import set_cta_group
    eng, model = engine_b
    fr = common.noise_frames(5, 224, 224, seed=4)
    pv = preprocess_ref.preprocess(fr, bgr=True)
    ref_tok = vit_ref.vit_forward(model.state_dict(), torch.from_numpy(pv), heads=12, layers=12)
    set_cta_group(cg)
    _lib.set_tuning("ln_fold", ln_fold)
    _lib.set_tuning("resid_split", resid_split)
    try:
        patches = eng.preprocess(torch.from_numpy(fr).to(eng.device), bgr=True)
        emb, tok = eng.forward_patches(patches, 5, want_tokens=True)
    finally:
        set_cta_group(2)
        _lib.set_tuning("ln_fold", 1)
        _lib.set_tuning("resid_split", 1)
    tok, emb = tok.cpu(), emb.cpu()
    assert torch.isfinite(tok).all()
    assert ((tok - ref_tok).abs().max() / ref_tok.abs().max()).item() < 2e-2        # bf16 operands through 12 layers
    cos = common.cosine(emb.numpy(), ref_tok.mean(1).numpy())
    assert (cos >= 0.9999).all(), cos


def test_extract_embedding_vs_reference_golden(engine_b, golden, tmp_path):
    from vision_sam3_yolo_lameless_b200.extractor import DINOv3Pipeline
    eng, _ = engine_b
    pipe = DINOv3Pipeline(eng, config=SUBJECTS, results_dir=tmp_path)
    want = np.load(golden / "embed_vitb.npz")
    for name, h, w, kind, seed in EMBED_CASES:
        e = pipe.extract_embedding(frames_for(kind, 1, h, w, seed)[0])
        assert e.shape == (768,) and e.dtype == np.float32
        assert common.cosine(e, want[name]) >= COS_GATE, name
        assert np.abs(e - want[name]).max() < 2e-2, name
    g = pipe.extract_embedding(common.noise_frames(1, 224, 224, 36)[0, :, :, 0])
    assert common.cosine(g, want["gray_224"]) >= COS_GATE


def test_vit_l_vs_reference_golden(golden):
    eng, _ = engine_cached("l", max_frames=8)
    want = np.load(golden / "embed_vitl.npz")
    for name, h, w, kind, seed in EMBED_CASES[:2]:
        e = eng.embed_host_frames(frames_for(kind, 1, h, w, seed)).cpu().numpy()[0]
        assert e.shape == (1024,) and common.cosine(e, want[name]) >= COS_GATE, name


@pytest.mark.parametrize("size,t", [(518, 1029), (592, 1374)])
def test_large_grid_vs_fp32_oracle(size, t):
    """BASELINE config 3: 518x518 (32x32 patches, 6 px unused border) and 592x592 (1369 patches); no resize."""
    eng, model = engine_cached("b", layers=2, max_frames=2, resize=(size, size))
    assert eng.tokens == t
    fr = common.smooth_frames(2, size, size, seed=6)
    pv = preprocess_ref.preprocess(fr, bgr=True, size=(size, size))
    ref = vit_ref.frame_embeddings(model.state_dict(), torch.from_numpy(pv), heads=12, layers=2).numpy()
    got = eng.embed_host_frames(fr).cpu().numpy()
    assert (common.cosine(got, ref) >= 0.9999).all()


def test_clip_embeddings_meet_cosine_gate(engine_b):
    """Per-clip mean-pooled, L2-normalised embedding vs the fp32 oracle; host frames through the pipelined H2D path,
    chunked (max_frames=64 < 150 frames) and ragged clips."""
    eng, model = engine_b
    fr = np.concatenate([common.noise_frames(80, 224, 224, seed=8), common.smooth_frames(70, 224, 224, seed=9)])
    offs = np.array([0, 50, 80, 150], dtype=np.int32)
    emb = eng.embed_host_frames(fr)
    mean, unit = eng.pool_clips(emb, torch.from_numpy(offs))
    pv = preprocess_ref.preprocess(fr, bgr=True)
    ref = vit_ref.frame_embeddings(model.state_dict(), torch.from_numpy(pv), heads=12, layers=12).numpy()
    ref_unit = reid_ref.l2_normalise(reid_ref.clip_mean(ref, offs))
    assert (common.cosine(emb.cpu().numpy(), ref) >= COS_GATE).all()
    cos = common.cosine(unit.cpu().numpy(), ref_unit)
    assert (cos >= COS_GATE).all(), cos
    np.testing.assert_allclose(np.linalg.norm(unit.cpu().numpy(), axis=1), 1.0, atol=1e-5)
    # batch invariance: a frame's embedding does not depend on what else is in the batch (bitwise)
    again = eng.embed_host_frames(fr[40:47])
    assert torch.equal(again, emb[40:47])
    # pinned input takes the zero-staging path and gives the same bits
    pinned = torch.from_numpy(fr).pin_memory()
    assert torch.equal(eng.embed_host_frames(pinned), emb)


def test_real_clip_vs_reference_golden(engine_b, golden, tmp_path):
    """The clip the reference ships (1280 x 720 H.264, 125 frames at 25 fps) through DINOv3Pipeline.extract_video_embeddings: same
    sampled frames / times / canonical frames as the reference's own run (tests/golden/canonical_clips.npz), every frame embedding at
    cosine >= 0.999; K1 (the TMA variant: 720p qualifies) against the HF processor's pixel_values of the first decoded frame."""
    import cv2
    from vision_sam3_yolo_lameless_b200.extractor import DINOv3Pipeline
    eng, _ = engine_b
    want = np.load(golden / "canonical_clips.npz")
    name = str(want["clip_names"][0])
    key = name.split("-")[0]
    pipe = DINOv3Pipeline(eng, config=SUBJECTS, nats_client=fake_services.FakeNats(), qdrant_client=fake_services.FakeQdrant(),
                          results_dir=tmp_path)
    got = pipe.extract_video_embeddings(golden / name)
    assert [e["frame"] for e in got["embeddings"]] == want[f"{key}_frames"].tolist() == [0, 25, 50, 75, 100]
    assert [e["time"] for e in got["embeddings"]] == want[f"{key}_times"].tolist()
    assert [e["frame"] for e in got["canonical_frames"]] == want[f"{key}_canonical"].tolist()
    assert [got["total_frames"], got["fps"]] == want[f"{key}_meta"].tolist()
    cos = common.cosine(np.array([e["embedding"] for e in got["embeddings"]]), want[f"{key}_embeddings"])
    assert (cos >= COS_GATE).all(), cos
    cap = cv2.VideoCapture(str(golden / name))
    ok, frame = cap.read()
    cap.release()
    assert ok
    patches = eng.preprocess(torch.from_numpy(frame[None]).to(eng.device), bgr=True)
    ref = preprocess_ref.patchify(want["frame0_pixel_values"][None])
    assert (patches.float().cpu() - torch.from_numpy(ref)).abs().max().item() < 1.2e-2      # bf16 rounding of |x| <= 2.64


def test_process_video_vs_reference_transcript(engine_b, golden, tmp_path):
    from vision_sam3_yolo_lameless_b200.extractor import DINOv3Pipeline
    eng, _ = engine_b
    want = json.load(open(golden / "process_video.json"))
    for backend in ("qdrant", "gpu"):
        qd, nats = fake_services.FakeQdrant(), fake_services.FakeNats()
        out = tmp_path / backend
        pipe = DINOv3Pipeline(eng, config=SUBJECTS, nats_client=nats, qdrant_client=qd, results_dir=out, gallery_backend=backend)
        for i, (seed, t) in enumerate(zip(want["clip_seeds"], want["transcript"])):
            clip = tmp_path / f"{backend}{i}.avi"
            write_clip(clip, seed=seed)
            asyncio.run(pipe.process_video({"video_id": f"vid-{i}", "processed_path": str(clip), "filename": clip.name,
                                            "metadata": {"n": i}}))
            res, ref = json.load(open(out / f"vid-{i}_dinov3.json")), t["results"]
            assert list(res) == list(ref) and res["num_embeddings"] == ref["num_embeddings"]
            assert res["neighbor_evidence"] == ref["neighbor_evidence"]
            assert [c["video_id"] for c in res["similar_cases"]] == [c["video_id"] for c in ref["similar_cases"]]
            np.testing.assert_allclose([c["score"] for c in res["similar_cases"]], [c["score"] for c in ref["similar_cases"]], atol=5e-3)
            for a, b in zip(res["canonical_frames"], ref["canonical_frames"]):
                assert a["frame"] == b["frame"] and a["time"] == b["time"]
                assert common.cosine(np.array(a["embedding"]), np.array(b["embedding"])) >= COS_GATE
            assert nats.published[-1][0] == "pipeline.dinov3" and list(nats.published[-1][1]) == list(t["message"])
            qd.set_payload("cow_embeddings", {"label": i % 2}, [f"vid-{i}"])
            if pipe.gallery is not None:
                pipe.gallery.payloads[pipe.gallery._row_of[f"vid-{i}"]]["label"] = i % 2


def test_process_videos_coalesced_equals_sequential_on_gpu(engine_b, tmp_path):
    """Queued messages through process_videos (one GPU batch per frame size) = the same messages through process_video one by one:
    every kernel of the path is row-independent, so the stored vectors must be bit-identical, not merely close."""
    from vision_sam3_yolo_lameless_b200.extractor import DINOv3Pipeline
    eng, _ = engine_b
    msgs = []
    for i, (seed, hw) in enumerate([(81, (48, 64)), (82, (96, 128)), (83, (48, 64)), (84, (96, 128)), (85, (48, 64))]):
        clip = tmp_path / f"c{i}.avi"
        write_clip(clip, seed=seed, h=hw[0], w=hw[1])
        msgs.append({"video_id": f"vid-{i}", "processed_path": str(clip), "filename": clip.name})
    msgs.insert(1, {"video_id": "missing", "processed_path": str(tmp_path / "nope.avi")})
    stores = {}
    for mode in ("seq", "bat"):
        qd, nats = fake_services.FakeQdrant(), fake_services.FakeNats()
        pipe = DINOv3Pipeline(eng, config=SUBJECTS, nats_client=nats, qdrant_client=qd, results_dir=tmp_path / mode, gallery_backend="gpu")
        if mode == "seq":
            for m in msgs:
                asyncio.run(pipe.process_video(dict(m)))
        else:
            asyncio.run(pipe.process_videos([dict(m) for m in msgs]))
        stores[mode] = (qd, nats)
    ids = [f"vid-{i}" for i in range(5)]
    for a, b in zip(stores["seq"][0].retrieve("cow_embeddings", ids, with_vectors=True),
                    stores["bat"][0].retrieve("cow_embeddings", ids, with_vectors=True)):
        assert a.payload == b.payload
        np.testing.assert_array_equal(np.array(a.vector), np.array(b.vector))
    strip = lambda pub: [(s, {k: v for k, v in m.items() if k != "results_path"}) for s, m in pub]
    assert strip(stores["seq"][1].published) == strip(stores["bat"][1].published)
    assert [m["video_id"] for _, m in stores["bat"][1].published] == ids


def test_matcher_scenario_vs_reference_transcript(engine_b, golden):
    from vision_sam3_yolo_lameless_b200.reid import CowReIDMatcher
    eng, _ = engine_b
    want = json.load(open(golden / "reid_scenario.json"))
    qd = fake_services.FakeQdrant()
    m = CowReIDMatcher(engine=eng, qdrant_client=qd)
    asyncio.run(m.connect())
    for k, ((name, q), step) in enumerate(zip(reid_queries(), want["steps"])):
        got = m.match_or_create(np.asarray(q), video_id=f"video-{name}", track_id=k)
        assert (got.cow_id, got.confidence, got.is_new_identity) == (step["cow_id"], step["confidence"], step["is_new"]), name
        assert abs(got.similarity - step["similarity"]) < 4e-4, name        # bf16 rounding of the scan copy only (fp32 master rows)
    _, cands = m.match_embedding(np.asarray(reid_queries()[0][1]))
    assert [c.cow_id for c in cands] == [c["cow_id"] for c in want["final_candidates"]]
    np.testing.assert_allclose([c.similarity for c in cands], [c["similarity"] for c in want["final_candidates"]], atol=4e-4)


def test_matcher_near_thresholds_and_durable_store(engine_b, golden):
    """The reference matcher's own transcript 2e-3 on either side of 0.65 / 0.75 / 0.85 (tests/golden/reid_tight.npz): same decisions,
    similarities within the bf16 scan-copy error, and Qdrant receives the fp32 master vector (matcher.py:228-246,281-301)."""
    from conftest import replay_reid_tight
    worst = replay_reid_tight(engine_b[0], golden, sim_tol=4e-4)
    print(f"largest similarity difference vs the reference transcript: {worst:.2e}")


def test_sharded_matcher_on_two_gpus():
    """SURVEY 8(e) behind the matcher on real hardware: tests/multigpu_matcher.py under torchrun with 2 ranks over NCCL (skipped on a
    one-GPU box; profiles/r02*_multigpu_matcher.log holds the 2- and 8-GPU runs)."""
    import subprocess
    import sys
    from pathlib import Path

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on the box (run tests/multigpu_matcher.py under torchrun; see profiles/)")
    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(root / "tests" / "multigpu_matcher.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTIGPU_MATCHER_OK world=2" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_full_size_properties(engine_b):
    """BASELINE sizes, size-independent properties: (1) a 100k-row gallery scan returns, for queries that ARE gallery
    rows, that row first with score ~1; (2) permuting the clips permutes the result; (3) 8-way sharding + merge is
    idempotent w.r.t. the single scan; (4) a 1080p batch gives the same bits as the same frames one by one."""
    eng, _ = engine_b
    dev = eng.device
    gen = torch.Generator(device=dev).manual_seed(7)
    n, q = 100_000, 1024
    gal = torch.nn.functional.normalize(torch.randn(n, 768, device=dev, generator=gen), dim=1).to(torch.bfloat16)
    rows = torch.randperm(n, device=dev, generator=gen)[:q]
    qv = torch.nn.functional.normalize(gal[rows].float(), dim=1)
    s, i = eng.gallery_topk(qv, gal, k=5)
    assert torch.equal(i[:, 0].long(), rows) and (s[:, 0] > 0.998).all()
    assert (s[:, :-1] >= s[:, 1:]).all(), "scores must be sorted descending"
    perm = torch.randperm(q, device=dev, generator=gen)
    s2, i2 = eng.gallery_topk(qv[perm].contiguous(), gal, k=5)
    assert torch.equal(i2, i[perm]) and torch.equal(s2, s[perm])
    from vision_sam3_yolo_lameless_b200.sharded import shard_range
    parts = [eng.gallery_topk(qv, gal[lo:hi], k=5, row_base=lo) for lo, hi in (shard_range(n, r, 8) for r in range(8))]
    ms, mi = eng.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(mi, i) and torch.equal(ms, s)
    big = torch.randint(0, 256, (6, 1080, 1920, 3), device=dev, dtype=torch.uint8, generator=gen)
    whole = eng.embed_frames(big)
    single = torch.cat([eng.embed_frames(big[j:j + 1]) for j in range(6)])
    assert torch.equal(whole, single)


def test_knn_graph_vs_reference_golden(engine_b, golden):
    """the GNN service's kNN graph (gnn main.py:55-100) from one K4 launch (Q = N, self dropped)."""
    from oracle.make_golden import knn_embeddings
    from vision_sam3_yolo_lameless_b200.knn_graph import GraphBuilder
    eng, _ = engine_b
    want = json.load(open(golden / "knn_graph.json"))
    from conftest import assert_knn_equivalent
    emb = knn_embeddings()
    ei, ew = GraphBuilder(eng, k_neighbors=5).compute_knn_edges(emb)
    assert_knn_equivalent(ei, ew, want["edge_index"], want["edge_weights"], emb, 5)
    ei2, ew2 = GraphBuilder(eng).compute_knn_edges(emb[:4])
    assert_knn_equivalent(ei2, ew2, want["small_edge_index"], want["small_edge_weights"], emb[:4], 3)


def test_track_crop_embeddings_vs_reference_golden(engine_b, golden, tmp_path):
    """per-box embeddings through K1's region-of-interest mode match the REFERENCE's extract_embedding on the numpy
    crop (tests/golden/roi_crops.npz), and extract_track_embeddings averages a track's boxes like a clip."""
    from oracle.make_golden_roi import ROI_CASES, ROI_FRAMES
    from vision_sam3_yolo_lameless_b200.extractor import DINOv3Pipeline
    eng, _ = engine_b
    want = np.load(golden / "roi_crops.npz")
    kind, n, h, w, seed = ROI_FRAMES
    fr = frames_for(kind, n, h, w, seed)
    rois = [(f, x0, y0, x1, y1) for _, f, x0, y0, x1, y1 in ROI_CASES]
    emb = eng.embed_rois(torch.from_numpy(fr).to(eng.device), rois).cpu().numpy()
    for i, (name, *_rest) in enumerate(ROI_CASES):
        assert common.cosine(emb[i], want["emb_" + name]) >= COS_GATE, name
        assert np.abs(emb[i] - want["emb_" + name]).max() < 2e-2, name
    pipe = DINOv3Pipeline(eng, config=SUBJECTS, results_dir=tmp_path)
    recs = [{"frame": f, "track_id": 1 if i < 3 else 2, "bbox": [x0, y0, x1, y1]} for i, (f, x0, y0, x1, y1) in enumerate(rois)]
    tracks = pipe.extract_track_embeddings(fr, recs)
    assert common.cosine(tracks[1], emb[:3].mean(0)) > 0.999999 and common.cosine(tracks[2], emb[3:].mean(0)) > 0.999999
    # the crop path and the whole-image path agree on the same pixels
    assert common.cosine(emb[1], pipe.extract_embedding(fr[0])) > 0.99999
