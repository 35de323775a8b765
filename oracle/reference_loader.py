"""Imports the reference's own hot-path modules UNMODIFIED from /root/reference with its absent
service dependencies (qdrant_client, nats) stubbed, and builds a ``DINOv3Pipeline`` without running its
network-bound ``__init__`` (HF hub download at services/dinov3-pipeline/app/main.py:34-35, Qdrant
connect :39-40).  Only usable where /root/reference exists (the authoring container): used by
``make_golden.py`` to produce tests/golden/*.npz and by CPU tests that are skipped when it is absent.
Test infrastructure only.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "services/dinov3-pipeline/app/main.py").exists()


def _stub_modules() -> None:
    def mod(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
        return sys.modules[name]

    class _Any:
        def __init__(self, *a, **k):
            self.args, self.kwargs = a, k
            for key, val in k.items():
                setattr(self, key, val)

    class _Distance:
        COSINE = "Cosine"

    mod("qdrant_client", QdrantClient=_Any)
    mod("qdrant_client.models", Distance=_Distance, VectorParams=_Any, PointStruct=_Any,
        Filter=_Any, FieldCondition=_Any, MatchValue=_Any)
    mod("qdrant_client.http", models=sys.modules["qdrant_client.models"])
    mod("qdrant_client.http.models", Distance=_Distance, VectorParams=_Any, PointStruct=_Any)
    mod("nats")
    mod("nats.aio")
    mod("nats.aio.client", Client=_Any)


def _load(path: Path, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_dinov3_module():
    _stub_modules()
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    return _load(REFERENCE_ROOT / "services/dinov3-pipeline/app/main.py", "ref_dinov3_main")


def load_matcher_module():
    _stub_modules()
    return _load(REFERENCE_ROOT / "services/tracking-service/app/reid/matcher.py", "ref_reid_matcher")


def make_reference_pipeline(model, processor):
    """object.__new__(DINOv3Pipeline) with a model/processor injected; extract_embedding and
    extract_video_embeddings then run exactly the reference's code."""
    import torch

    ref = load_dinov3_module()
    pipe = object.__new__(ref.DINOv3Pipeline)
    pipe.device = torch.device("cpu")
    pipe.processor = processor
    pipe.model = model.eval()
    return pipe
