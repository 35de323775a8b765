"""pytest configuration.  `-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbols (no compute).
`-m gpu`: parity of the CUDA path against the oracle / golden vectors, called through the C-ABI."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) and the built libcre_b200.so")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never pass (or skip) silently on a box without CUDA: when they are selected there (`-m gpu`, or no `-m` at
    all), each one is made to FAIL with the reason; `-m "not gpu"` deselects them before this hook sees them."""
    import torch

    if torch.cuda.is_available():
        return

    def _no_gpu(*_a, **_k):
        pytest.fail("GPU test selected on a box without CUDA (run with -m 'not gpu' here; there is no CPU fallback)", pytrace=False)

    for item in items:
        if item.get_closest_marker("gpu") is not None:
            item.obj = _no_gpu


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


_MODELS = {}


def hf_model_cached(kind="b", layers=None):
    """Seeded random-init HF models are expensive to build; share them across tests."""
    from oracle import common

    key = (kind, layers)
    if key not in _MODELS:
        if kind == "b":
            _MODELS[key] = common.hf_model(layers=12 if layers is None else layers)
        else:
            _MODELS[key] = common.hf_model(hidden=1024, mlp=4096, layers=24 if layers is None else layers, heads=16)
    return _MODELS[key]


@pytest.fixture(scope="session")
def model_b():
    return hf_model_cached("b")


_ENGINES = {}


def engine_cached(kind="b", layers=None, max_frames=64, resize=(224, 224)):
    from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine, VitConfig

    key = (kind, layers, max_frames, resize)
    if key not in _ENGINES:
        m = hf_model_cached(kind, layers)
        _ENGINES[key] = (ClipEmbedEngine(VitConfig.from_hf(m.config), m.state_dict(), max_frames=max_frames, resize=resize), m)
    return _ENGINES[key]


@pytest.fixture(scope="session")
def engine_b():
    """ViT-B/16, 12 layers, seed 0 -- the model the golden vectors were produced with."""
    return engine_cached("b")


@pytest.fixture(scope="session")
def engine_small():
    """ViT-B/16 width, 1 layer: enough for kernel-level tests that only need a context."""
    return engine_cached("b", layers=1, max_frames=16)[0]


def assert_knn_equivalent(ei, ew, want_ei, want_ew, emb, k, tol=6e-3):
    """Edge lists agree up to swaps among near-tied neighbours (the GPU gallery is bf16: scores move by <= ~4e-3):
    same sources, per-node weights equal within tol position by position, and every chosen neighbour's TRUE cosine is
    within tol of the reference's k-th best."""
    import numpy as np
    ei, want_ei = np.asarray(ei), np.asarray(want_ei)
    assert ei.shape == want_ei.shape and (ei[0] == want_ei[0]).all()
    np.testing.assert_allclose(ew, want_ew, atol=tol)
    e = emb / (np.linalg.norm(emb, axis=1, keepdims=True) + 1e-8)
    sim = e @ e.T
    for i in range(len(emb)):
        sel = ei[1][ei[0] == i]
        ref = want_ei[1][want_ei[0] == i]
        assert len(set(sel.tolist())) == len(sel) and i not in sel
        assert (sim[i, sel] >= sim[i, ref].min() - tol).all(), i
        assert (np.diff(np.asarray(ew)[ei[0] == i]) >= -1e-6).all(), "neighbours must come in ascending similarity"


def replay_reid_tight(engine, golden, sim_tol):
    """Replay tests/golden/reid_tight.npz (oracle/make_golden_reid.py: the reference matcher driven 2e-3 on either side of every
    threshold) through CowReIDMatcher on `engine`: same decisions and sighting counts, similarities within sim_tol, and the durable
    store receives what the reference stores -- the full-precision normalised / momentum-blended vector (matcher.py:228-246,281-301),
    not a bf16 rounding of it.  Returns the largest similarity difference seen."""
    import asyncio
    import json

    import numpy as np
    from oracle import fake_services
    from vision_sam3_yolo_lameless_b200.reid import CowReIDMatcher

    data = np.load(golden / "reid_tight.npz")
    steps = json.loads(str(data["transcript"]))
    qd = fake_services.FakeQdrant()
    m = CowReIDMatcher(qdrant_url="fake://", engine=engine, qdrant_client=qd)
    asyncio.run(m.connect())
    worst = 0.0
    for k, (q, want_vec, step) in enumerate(zip(data["queries"], data["stored_after"], steps)):
        got = m.match_or_create(np.asarray(q), video_id=f"video-{step['name']}", track_id=k)
        assert (got.cow_id, got.confidence, got.is_new_identity) == (step["cow_id"], step["confidence"], step["is_new"]), step["name"]
        worst = max(worst, abs(got.similarity - step["similarity"]))
        col = qd.collections[CowReIDMatcher.COLLECTION_NAME]
        row = col["ids"].index(str(got.identity_id))
        assert col["payloads"][row]["total_sightings"] == step["total_sightings"], step["name"]
        np.testing.assert_allclose(col["vectors"][row], want_vec, atol=2e-6, err_msg=step["name"])    # fp32 master vs the reference's fp64
    assert worst < sim_tol, worst
    return worst
