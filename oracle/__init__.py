"""CPU oracle for the clip-embedding + re-ID hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vision_sam3_yolo_lameless_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` do, and there only as the checker / the timed CPU baseline, never as the product path.

Pinning status
--------------
* Embedding path (preprocess + ViT + pooling): the arithmetic lives in HuggingFace ``transformers``
  (reference pins ``transformers>=4.35.0``, services/dinov3-pipeline/environment.yml:11; this image has
  5.5.0) and torch/torchvision aten kernels.  The restatement here is pinned against outputs of the
  reference's OWN ``DINOv3Pipeline.extract_embedding`` / ``extract_video_embeddings`` imported unmodified
  from /root/reference (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
* Re-ID similarity + top-k: computed by a remote Qdrant server (``qdrant/qdrant:latest``, un-pinned,
  docker-compose.yml:34), source absent from the reference and no reference test holds its outputs:
  **parity unpinned** at that boundary.  The oracle restates Qdrant's documented COSINE semantics
  (normalise both sides, dot product, descending score) and defines the tie order (score desc,
  index asc); threshold logic is pinned against the reference's ``CowReIDMatcher`` constants.
"""
