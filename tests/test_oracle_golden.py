"""The CPU oracle (oracle/) against the reference's own outputs (tests/golden/, produced by
oracle/make_golden.py from the unmodified reference modules).  CPU only."""
import json

import numpy as np
import pytest
import torch

from oracle import common, pipeline_ref, preprocess_ref, reid_ref, vit_ref
from oracle.make_golden import EMBED_CASES, PREPROCESS_CASES, frames_for, reid_queries

from conftest import hf_model_cached


@pytest.mark.parametrize("case", PREPROCESS_CASES, ids=[c[0] for c in PREPROCESS_CASES])
def test_preprocess_restatement_matches_hf_processor(golden, case):
    name, n, h, w, kind, seed = case
    want = np.load(golden / "preprocess.npz")[name]
    got = preprocess_ref.preprocess(frames_for(kind, n, h, w, seed), bgr=True)[0]
    assert got.shape == want.shape == (3, 224, 224)
    np.testing.assert_allclose(got, want, atol=3e-5, rtol=0)


def test_aa_weights_properties():
    for n_in, n_out in [(1920, 224), (1080, 224), (224, 224), (100, 224), (482, 224)]:
        lo, cnt, w = preprocess_ref.aa_weights(n_in, n_out)
        np.testing.assert_allclose(w.sum(axis=1), 1.0, atol=1e-6)
        assert (lo >= 0).all() and (lo + cnt <= n_in).all() and (cnt >= 1).all()
    lo, cnt, w = preprocess_ref.aa_weights(224, 224)       # identity resize: one tap of weight 1 ... plus zero taps
    assert np.allclose(np.take_along_axis(w, (np.arange(224) - lo)[:, None], 1)[:, 0], 1.0)
    assert preprocess_ref.aa_weights(1920, 224)[1].max() <= 19


def test_patchify_is_conv_weight_order():
    pv = np.random.default_rng(0).standard_normal((2, 3, 32, 48)).astype(np.float32)
    w = np.random.default_rng(1).standard_normal((8, 3, 16, 16)).astype(np.float32)
    conv = torch.nn.functional.conv2d(torch.from_numpy(pv), torch.from_numpy(w), stride=16).flatten(2).transpose(1, 2).reshape(-1, 8)
    mm = preprocess_ref.patchify(pv) @ w.reshape(8, -1).T
    np.testing.assert_allclose(mm, conv.numpy(), atol=1e-3)


def test_vit_restatement_matches_reference_extract_embedding(golden, model_b):
    want = np.load(golden / "embed_vitb.npz")
    sd = model_b.state_dict()
    for name, h, w, kind, seed in EMBED_CASES[:4]:
        pv = preprocess_ref.preprocess(frames_for(kind, 1, h, w, seed), bgr=True)
        got = vit_ref.frame_embeddings(sd, torch.from_numpy(pv), heads=12, layers=12).numpy()[0]
        np.testing.assert_allclose(got, want[name], atol=2e-4, rtol=0, err_msg=name)
        assert common.cosine(got, want[name]) > 0.999999


def test_real_clip_oracle_matches_reference(golden, model_b):
    """Real pixels (the clip the reference ships, 1280 x 720 H.264): the oracle's decode loop + per-frame path against the
    reference's own extract_video_embeddings output, and the numpy antialias restatement against the HF processor on frame 0."""
    import cv2

    want = np.load(golden / "canonical_clips.npz")
    name = str(want["clip_names"][0])
    key = name.split("-")[0]
    pipe = pipeline_ref.ReferencePipelineCPU(model_b)
    got = pipe.extract_video_embeddings(golden / name)
    assert [e["frame"] for e in got["embeddings"]] == want[f"{key}_frames"].tolist() == [0, 25, 50, 75, 100]
    assert [e["time"] for e in got["embeddings"]] == want[f"{key}_times"].tolist()
    assert [e["frame"] for e in got["canonical_frames"]] == want[f"{key}_canonical"].tolist()
    assert [got["total_frames"], got["fps"]] == want[f"{key}_meta"].tolist() == [125, 25]
    np.testing.assert_allclose(np.array([e["embedding"] for e in got["embeddings"]]), want[f"{key}_embeddings"], atol=2e-4)
    assert pipeline_ref.sampled_frame_indices(125, 25.0) == [0, 25, 50, 75, 100]
    cap = cv2.VideoCapture(str(golden / name))
    ok, frame = cap.read()
    cap.release()
    assert ok and frame.shape == (720, 1280, 3)
    pv = preprocess_ref.preprocess(frame[None], bgr=True)[0]
    np.testing.assert_allclose(pv, want["frame0_pixel_values"], atol=3e-5, rtol=0)


def test_gray_image_passthrough(golden, model_b):
    """main.py:98-101: 2-D input skips cvtColor; PIL 'L' image -> HF processor replicates to 3 channels."""
    want = np.load(golden / "embed_vitb.npz")["gray_224"]
    g = common.noise_frames(1, 224, 224, 36)[0, :, :, 0]
    pv = preprocess_ref.preprocess(np.repeat(g[None, :, :, None], 3, axis=3), bgr=False)
    got = vit_ref.frame_embeddings(model_b.state_dict(), torch.from_numpy(pv), heads=12, layers=12).numpy()[0]
    np.testing.assert_allclose(got, want, atol=2e-4, rtol=0)


def test_vit_l_restatement(golden):
    want = np.load(golden / "embed_vitl.npz")
    m = hf_model_cached("l")
    name, h, w, kind, seed = EMBED_CASES[0]
    pv = preprocess_ref.preprocess(frames_for(kind, 1, h, w, seed), bgr=True)
    got = vit_ref.frame_embeddings(m.state_dict(), torch.from_numpy(pv), heads=16, layers=24).numpy()[0]
    np.testing.assert_allclose(got, want[name], atol=3e-4, rtol=0)


def test_pipeline_restatement_matches_reference_video(golden, model_b):
    want = json.load(open(golden / "video.json"))
    pipe = pipeline_ref.ReferencePipelineCPU(model_b)
    got = pipe.extract_video_embeddings(golden / "clip_48x64_15fps.avi")
    assert got["fps"] == want["fps"] and got["total_frames"] == want["total_frames"]
    assert [e["frame"] for e in got["embeddings"]] == want["frames"] == pipeline_ref.sampled_frame_indices(45, 15.0)
    assert [e["time"] for e in got["embeddings"]] == want["times"]
    assert [e["frame"] for e in got["canonical_frames"]] == want["canonical"]
    np.testing.assert_allclose(np.array([e["embedding"] for e in got["embeddings"]]), np.array(want["embeddings"]), atol=1e-5)


def test_fps_truncation_rule():
    assert pipeline_ref.sampled_frame_indices(150, 29.97) == [0, 29, 58, 87, 116, 145]
    assert pipeline_ref.sampled_frame_indices(150, 30) == [0, 30, 60, 90, 120]
    assert pipeline_ref.sampled_frame_indices(5, 0) == [0, 1, 2, 3, 4]


def test_reid_thresholds_and_scenario(golden):
    want = json.load(open(golden / "reid_scenario.json"))
    assert want["thresholds"] == [reid_ref.SIMILARITY_THRESHOLD_HIGH, reid_ref.SIMILARITY_THRESHOLD_MEDIUM,
                                  reid_ref.SIMILARITY_THRESHOLD_LOW]
    for s, label in want["confidence_probe"].items():
        assert reid_ref.score_to_confidence(float(s)) == label
    m = reid_ref.MatcherOracle()
    for (name, q), step in zip(reid_queries(), want["steps"]):
        got = m.match_or_create(q)
        assert (got["cow_id"], got["confidence"], got["is_new"]) == (step["cow_id"], step["confidence"], step["is_new"]), name
        assert abs(got["similarity"] - step["similarity"]) < 1e-6
    _, cands = m.match_embedding(reid_queries()[0][1])
    assert [c["cow_id"] for c in cands] == [c["cow_id"] for c in want["final_candidates"]]
    np.testing.assert_allclose([c["similarity"] for c in cands], [c["similarity"] for c in want["final_candidates"]], atol=1e-6)


def test_neighbor_evidence_rule(golden):
    for t in json.load(open(golden / "process_video.json"))["transcript"]:
        assert reid_ref.neighbor_evidence(t["results"]["similar_cases"]) == t["results"]["neighbor_evidence"]


def test_topk_rule_ties_and_merge():
    s = np.array([[0.5, 0.9, 0.9, 0.1, 0.9, -1.0]], dtype=np.float32)
    top, idx = reid_ref.topk_rule(s, 4, row_base=10)
    assert idx.tolist() == [[11, 12, 14, 10]]
    a_s, a_i = reid_ref.topk_rule(s[:, :3], 2, row_base=0)
    b_s, b_i = reid_ref.topk_rule(s[:, 3:], 2, row_base=3)
    ms, mi = reid_ref.merge_rule(np.stack([b_s, a_s]), np.stack([b_i, a_i]), 3)
    assert mi.tolist() == [[1, 2, 4]]
    cm = reid_ref.clip_mean(np.arange(12, dtype=np.float32).reshape(6, 2), np.array([0, 2, 6]))
    np.testing.assert_allclose(cm, [[1, 2], [7, 8]])


def test_knn_graph_restatement_matches_reference(golden):
    from oracle import knn_ref
    from oracle.make_golden import knn_embeddings
    want = json.load(open(golden / "knn_graph.json"))
    emb = knn_embeddings()
    ei, ew = knn_ref.compute_knn_edges(emb, 5)
    assert ei.tolist() == want["edge_index"]
    np.testing.assert_allclose(ew, want["edge_weights"], atol=1e-12)
    ei2, ew2 = knn_ref.compute_knn_edges(emb[:4], 5)
    assert ei2.tolist() == want["small_edge_index"] and ei2.shape == (2, 12)


def test_crop_path_matches_reference_on_crops(golden, model_b):
    """Per-track crops: the oracle applied to frame[y0:y1, x0:x1] reproduces the reference's extract_embedding on
    that crop (tests/golden/roi_crops.npz, oracle/make_golden_roi.py) and the HF processor's pixel_values."""
    from oracle.make_golden_roi import ROI_CASES, ROI_FRAMES
    want = np.load(golden / "roi_crops.npz")
    kind, n, h, w, seed = ROI_FRAMES
    fr = frames_for(kind, n, h, w, seed)
    sd = model_b.state_dict()
    for name, f, x0, y0, x1, y1 in ROI_CASES:
        pv = preprocess_ref.preprocess(np.ascontiguousarray(fr[f:f + 1, y0:y1, x0:x1]), bgr=True)
        if "pix_" + name in want:
            np.testing.assert_allclose(pv[0], want["pix_" + name].astype(np.float32), atol=2e-3, rtol=0)   # stored as fp16
        got = vit_ref.frame_embeddings(sd, torch.from_numpy(pv), heads=12, layers=12).numpy()[0]
        np.testing.assert_allclose(got, want["emb_" + name], atol=2e-4, rtol=0, err_msg=name)
