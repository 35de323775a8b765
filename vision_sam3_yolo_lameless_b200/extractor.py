"""Drop-in replacement for the reference's embedding extractor
(``services/dinov3-pipeline/app/main.py`` class ``DINOv3Pipeline``, lines 21-296).

Same method names, argument meaning, return shapes, file / NATS / Qdrant outputs and error behaviour;
the arithmetic (BGR->RGB, HF processor, ViT forward, token mean, clip mean, similarity search) runs in
libcre_b200 kernels through :class:`ClipEmbedEngine`.  Differences, all deliberate and opt-in except (1):

1. frames of one video are embedded as ONE batch (the reference runs batch-1 forwards with a host
   sync per frame, main.py:107-113); per-frame results are identical up to bf16 tolerance.
2. ``gallery_backend="gpu"`` serves ``search_similar`` from a device mirror of the collection
   (GpuGallery) with write-through to Qdrant; default ``"qdrant"`` keeps the reference behaviour.
3. ``emit_embedding=True`` adds an ``"embedding"`` key to the results JSON (the reference omits it, which
   makes the tracking service fall back to canonical frames, tracking main.py:292-304).
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .engine import ClipEmbedEngine, VitConfig
from .gallery import GpuGallery

DEFAULT_CONFIG = {  # services/dinov3-pipeline/app/main.py:57-68 fallback config
    "qdrant": {"url": os.getenv("QDRANT_URL", "http://localhost:6333"), "collection_name": "cow_embeddings"},
    "nats": {"subjects": {"video_preprocessed": "video.preprocessed", "pipeline_dinov3": "pipeline.dinov3"}},
}


class _PointStruct:
    """Stand-in with qdrant_client.models.PointStruct's attribute surface (id, vector, payload), used when
    qdrant_client is not importable (tests, benchmarks).  The real class is used when available."""

    def __init__(self, id, vector, payload):
        self.id, self.vector, self.payload = id, vector, payload


def _point_struct(**kw):
    try:
        from qdrant_client.models import PointStruct  # type: ignore
        return PointStruct(**kw)
    except Exception:
        return _PointStruct(**kw)


class DINOv3Pipeline:
    """DINOv3 embedding extraction and VectorDB storage (B200-native)."""

    def __init__(self, engine: ClipEmbedEngine, config: Optional[dict] = None, nats_client=None, qdrant_client=None,
                 results_dir: Optional[Path] = None, gallery_backend: str = "qdrant", emit_embedding: bool = False,
                 frame_interval: Optional[int] = None):
        if gallery_backend not in ("qdrant", "gpu"):
            raise ValueError("gallery_backend must be 'qdrant' or 'gpu'")
        self.config = config if config is not None else DEFAULT_CONFIG
        self.engine = engine
        self.device = engine.device
        self.model = engine       # attribute kept for callers that poke at .model / .processor
        self.processor = None
        self.nats_client = nats_client
        self.qdrant_client = qdrant_client
        self.collection_name = self.config.get("qdrant", {}).get("collection_name", "cow_embeddings")
        self.gallery_backend = gallery_backend
        self.emit_embedding = emit_embedding
        self.frame_interval = frame_interval
        self.gallery: Optional[GpuGallery] = GpuGallery(engine, engine.cfg.hidden) if gallery_backend == "gpu" else None
        self.processed_dir = Path("/app/data/processed")
        self.results_dir = Path(results_dir) if results_dir is not None else Path("/app/data/results/dinov3")
        self.results_dir.mkdir(parents=True, exist_ok=True)
        self._ensure_collection()
        self._mirror_collection()

    # -- main.py:70-93 ---------------------------------------------------------------------------
    def _ensure_collection(self):
        if self.qdrant_client is None:
            return
        try:
            names = [c.name for c in self.qdrant_client.get_collections().collections]
            if self.collection_name not in names:
                try:
                    from qdrant_client.models import Distance, VectorParams  # type: ignore
                    vc = VectorParams(size=self.engine.cfg.hidden, distance=Distance.COSINE)
                except Exception:
                    vc = {"size": self.engine.cfg.hidden, "distance": "Cosine"}
                self.qdrant_client.create_collection(collection_name=self.collection_name, vectors_config=vc)
                print(f"Created Qdrant collection: {self.collection_name}")
        except Exception as e:
            print(f"Error ensuring collection: {e}")

    def _mirror_collection(self):
        """gallery_backend == 'gpu': load what the durable store already holds (service restart) into the device gallery, so that
        search_similar sees the same points the reference's Qdrant search would.  If the mirror cannot be loaded the device gallery is
        dropped and searches go to Qdrant, as with gallery_backend == 'qdrant' -- never a silently empty gallery."""
        if self.gallery is None or self.qdrant_client is None:
            return
        try:
            names = [c.name for c in self.qdrant_client.get_collections().collections]
            if self.collection_name in names:
                n = self.gallery.load_from_qdrant(self.qdrant_client, self.collection_name)
                print(f"Mirrored {n} points of {self.collection_name} into the GPU gallery")
        except Exception as e:
            print(f"Error mirroring {self.collection_name} into the GPU gallery ({e}); searching Qdrant instead")
            self.gallery = None
            self.gallery_backend = "qdrant"

    # -- batched entry points (new; the reference embeds one frame per call) ------------------------
    def embed_frames(self, frames: np.ndarray, bgr: bool = True) -> np.ndarray:
        """HOST uint8 [n, H, W, 3] (cv2 BGR by default) -> float32 [n, D] per-frame embeddings (token mean).
        H2D copies are pipelined with the kernels inside the engine."""
        return self.engine.embed_host_frames(frames, bgr=bgr).cpu().numpy()

    def embed_clips(self, frames, clip_offsets, bgr: bool = True, top_k: int = 0, sharded=None):
        """Many clips in one call: HOST frames uint8 [F, H, W, 3] (numpy or pinned CPU tensor), clip c = frames
        [clip_offsets[c], clip_offsets[c+1]).  Returns (clip_mean [Q, D], clip_unit [Q, D]) as numpy, plus
        (scores [Q, k], ids) against the GPU gallery when top_k > 0 and gallery_backend == 'gpu'.
        ``sharded`` (a :class:`ShardedReID`, one process per GPU, every rank calling with the same number of clips): the search runs
        against the ROW-SHARDED gallery instead -- all-gather of the queries, per-shard scan, all-gather + merge of the candidates
        (sharded.py) -- and this rank's rows of the global result are returned."""
        emb = self.engine.embed_host_frames(frames, bgr=bgr)
        mean, unit = self.engine.pool_clips(emb, torch.as_tensor(np.asarray(clip_offsets, dtype=np.int32)))
        if top_k > 0 and sharded is not None:
            scores, idx = sharded.search(unit, k=top_k)
            q, r = unit.shape[0], sharded.rank
            return mean.cpu().numpy(), unit.cpu().numpy(), scores[r * q:(r + 1) * q].cpu().numpy(), idx[r * q:(r + 1) * q].cpu().numpy()
        if top_k > 0 and self.gallery is not None and len(self.gallery) > 0:
            scores, idx = self.engine.gallery_topk(unit, self.gallery.matrix[: len(self.gallery)], k=top_k)
            return mean.cpu().numpy(), unit.cpu().numpy(), scores.cpu().numpy(), idx.cpu().numpy()
        return mean.cpu().numpy(), unit.cpu().numpy()

    # -- per-track crops (new; tracking main.py:332-334 "in production, you'd extract per-track embeddings") --------
    @staticmethod
    def track_boxes(frame_tracks, height: int, width: int, frame_index=None) -> Dict[int, List[Tuple[int, int, int, int, int]]]:
        """``frame_tracks`` = the tracking service's per-frame records ({"frame", "track_id", "bbox": [x1, y1, x2, y2]},
        tracking main.py:180-187) -> {track_id: [(row, x0, y0, x1, y1), ...]} with integer pixel boxes (floor / ceil, clamped
        to the frame, empty boxes dropped).  ``frame_index`` maps a video frame number to its row in the frame array
        (default: identity)."""
        out: Dict[int, List[Tuple[int, int, int, int, int]]] = {}
        for rec in frame_tracks:
            f = int(rec["frame"])
            row = f if frame_index is None else frame_index.get(f)
            if row is None:
                continue
            x1, y1, x2, y2 = (float(v) for v in rec["bbox"][:4])
            x0, y0 = max(0, int(np.floor(x1))), max(0, int(np.floor(y1)))
            xe, ye = min(width, int(np.ceil(x2))), min(height, int(np.ceil(y2)))
            if xe > x0 and ye > y0:
                out.setdefault(int(rec["track_id"]), []).append((row, x0, y0, xe, ye))
        return out

    def extract_track_embeddings(self, frames, frame_tracks, frame_index=None, bgr: bool = True) -> Dict[int, np.ndarray]:
        """One embedding per track: every box of the track is cropped from its frame, embedded exactly as
        ``extract_embedding(frame[y0:y1, x0:x1])`` would (K1 in region-of-interest mode: the frames are uploaded once, no
        crop is materialised), and the track's embeddings are averaged like a clip (main.py:204-208).
        frames: uint8 [n, H, W, 3] host array or CUDA tensor; returns {track_id: float32 [D]}."""
        dev_frames = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames))
        dev_frames = dev_frames.to(self.engine.device)
        n, h, w, _ = dev_frames.shape
        boxes = self.track_boxes(frame_tracks, h, w, frame_index)
        boxes = {t: [b for b in bs if 0 <= b[0] < n] for t, bs in boxes.items()}
        boxes = {t: bs for t, bs in boxes.items() if bs}
        if not boxes:
            return {}
        ids = sorted(boxes)
        rois = np.array([b for t in ids for b in boxes[t]], dtype=np.int64)
        offsets = np.cumsum([0] + [len(boxes[t]) for t in ids]).astype(np.int32)
        emb = self.engine.embed_rois(dev_frames, rois, bgr=bgr)
        mean, _ = self.engine.pool_clips(emb, torch.as_tensor(offsets))
        mean = mean.cpu().numpy()
        return {t: mean[i] for i, t in enumerate(ids)}

    # -- main.py:95-115 ---------------------------------------------------------------------------
    def extract_embedding(self, image: np.ndarray) -> np.ndarray:
        """Extract DINOv3 embedding from one image (H, W, 3) uint8 BGR -> (D,) float32."""
        if image.ndim == 3 and image.shape[2] == 3:
            return self.embed_frames(image[None], bgr=True)[0]
        # the reference passes non-3-channel input through unconverted (main.py:98-101); PIL then turns a
        # 2-D array into a grey image that the HF processor replicates to RGB
        if image.ndim == 2:
            return self.embed_frames(np.repeat(image[None, :, :, None], 3, axis=3), bgr=False)[0]
        raise ValueError(f"unsupported image shape {image.shape}")

    # -- main.py:117-163 --------------------------------------------------------------------------
    def _decode_sampled(self, video_path: Path):
        """The decode loop of extract_video_embeddings (main.py:119-146) without the per-frame forward: returns
        (staging uint8 [>= n, H, W, 3] or None, sampled frame numbers, fps, total_frames)."""
        import cv2

        cap = cv2.VideoCapture(str(video_path))
        if not cap.isOpened():
            raise Exception(f"Failed to open video: {video_path}")
        fps = int(cap.get(cv2.CAP_PROP_FPS))
        total_frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        frame_interval = self.frame_interval if self.frame_interval is not None else max(1, fps)
        # Same sampling rule as the reference (every frame is visited, frame_count % interval == 0 is kept), but a skipped frame
        # is only grab()bed -- no BGR conversion, no copy -- and a kept frame is retrieve()d straight into one pinned staging
        # array, so the sampled frames go host -> device with no intermediate list / np.stack.
        picked_idx: List[int] = []
        staging = None          # uint8 [capacity, H, W, 3]; pinned when CUDA is present
        frame_count = 0
        while True:
            if not cap.grab():
                break
            if frame_count % frame_interval == 0:
                n = len(picked_idx)
                if staging is not None and n < staging.shape[0]:
                    ret, frame = cap.retrieve(staging[n].numpy())
                else:
                    ret, frame = cap.retrieve()
                if not ret:
                    break
                if staging is None or n >= staging.shape[0] or frame.shape != tuple(staging.shape[1:]):
                    want = max(2 * n, (max(total_frames, frame_count + 1) + frame_interval - 1) // frame_interval + 1)
                    grown = torch.empty((want,) + frame.shape, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
                    if staging is not None and n:
                        grown[:n].copy_(staging[:n])
                    staging = grown
                if frame.ctypes.data != staging[n].data_ptr():      # cv2 allocated its own array (first frame, growth)
                    staging[n].copy_(torch.from_numpy(frame))
                picked_idx.append(frame_count)
            frame_count += 1
        cap.release()
        return staging, picked_idx, fps, total_frames

    @staticmethod
    def _embedding_records(picked_idx, embs, fps, total_frames) -> Dict[str, Any]:
        """The dict extract_video_embeddings returns (main.py:140-163), from the sampled frame numbers and their embeddings."""
        embeddings = [{"frame": idx, "time": idx / fps if fps > 0 else 0, "embedding": e.tolist()} for idx, e in zip(picked_idx, embs)]
        canonical_frames = []
        if embeddings:
            canonical_frames = [embeddings[0], embeddings[len(embeddings) // 2], embeddings[-1]]
        return {"embeddings": embeddings, "canonical_frames": canonical_frames, "total_frames": total_frames, "fps": fps}

    def extract_video_embeddings(self, video_path: Path) -> Dict[str, Any]:
        """Extract embeddings from video frames (1 frame per second, like the reference)."""
        staging, picked_idx, fps, total_frames = self._decode_sampled(video_path)
        embs = []
        if picked_idx:
            embs = self.engine.embed_host_frames(staging[: len(picked_idx)], bgr=True).cpu().numpy()
        return self._embedding_records(picked_idx, embs, fps, total_frames)

    # -- main.py:165-186 --------------------------------------------------------------------------
    def search_similar(self, query_embedding: np.ndarray, top_k: int = 5) -> List[Dict]:
        """Search for similar embeddings in the VectorDB; [] on any error."""
        try:
            if self.gallery is not None:
                results = self.gallery.search(query_embedding, top_k)
            else:
                results = self.qdrant_client.search(collection_name=self.collection_name,
                                                    query_vector=query_embedding.tolist(), limit=top_k)
            similar_cases = []
            for result in results:
                similar_cases.append({
                    "video_id": result.payload.get("video_id", "unknown"),
                    "score": float(result.score),
                    "label": result.payload.get("label", None),
                    "metadata": result.payload.get("metadata", {}),
                })
            return similar_cases
        except Exception as e:
            print(f"Error searching similar: {e}")
            return []

    # -- main.py:188-282 --------------------------------------------------------------------------
    async def process_video(self, video_data: dict):
        """Process a preprocessed video (NATS `video.preprocessed` handler).  Never raises for a bad video."""
        video_id = video_data["video_id"]
        processed_path = Path(video_data["processed_path"])
        print(f"DINOv3 pipeline processing video {video_id}")
        if not processed_path.exists():
            print(f"Processed video not found: {processed_path}")
            return
        try:
            embedding_data = self.extract_video_embeddings(processed_path)
            await self._finish_video(video_id, video_data, embedding_data)
        except Exception as e:
            print(f"Error in DINOv3 pipeline for {video_id}: {e}")
            import traceback
            traceback.print_exc()

    async def process_videos(self, messages: List[dict]) -> None:
        """Coalesced form of ``process_video`` for messages that are already queued (the admin UI's batch reprocess publishes one
        `video.preprocessed` per video in a loop, admin backend routers/pipeline.py:312-354; the reference then works through them one
        blocking handler call at a time).  The sampled frames of ALL messages go to the GPU as one batch per frame size; everything
        with side effects -- similarity search, VectorDB upsert, results JSON, `pipeline.dinov3` publish -- then runs per message IN
        ARRIVAL ORDER, so message k still sees the upserts of messages < k and every output is what k sequential ``process_video``
        calls produce.  A bad message (missing file, unreadable video) is reported and skipped exactly as there."""
        import traceback

        decoded = []                                  # (message, staging, picked_idx, fps, total_frames) in arrival order
        for video_data in messages:
            try:
                video_id = video_data["video_id"]
                processed_path = Path(video_data["processed_path"])
            except KeyError as e:                     # the reference raises inside the callback; nats_client.py:65-66 prints and goes on
                print(f"Error in message handler: {e}")
                continue
            print(f"DINOv3 pipeline processing video {video_id}")
            if not processed_path.exists():
                print(f"Processed video not found: {processed_path}")
                continue
            try:
                decoded.append((video_data,) + tuple(self._decode_sampled(processed_path)))
            except Exception as e:
                print(f"Error in DINOv3 pipeline for {video_id}: {e}")
                traceback.print_exc()
        # one engine call per frame size (K1 takes one H x W per launch sequence)
        by_shape: Dict[tuple, List[int]] = {}
        for i, (_, staging, picked, _, _) in enumerate(decoded):
            if picked:
                by_shape.setdefault(tuple(staging.shape[1:]), []).append(i)
        embs: Dict[int, Any] = {}
        failed: Dict[int, Exception] = {}
        for shape, members in by_shape.items():
            try:
                out = self.engine.embed_host_frames([decoded[i][1][: len(decoded[i][2])] for i in members], bgr=True).cpu().numpy()
                start = 0
                for i in members:
                    embs[i] = out[start:start + len(decoded[i][2])]
                    start += len(decoded[i][2])
            except Exception as e:                    # the whole group shares the failure, reported per message below
                for i in members:
                    failed[i] = e
        for i, (video_data, _, picked, fps, total_frames) in enumerate(decoded):
            video_id = video_data["video_id"]
            try:
                if i in failed:
                    raise failed[i]
                await self._finish_video(video_id, video_data, self._embedding_records(picked, embs.get(i, []), fps, total_frames))
            except Exception as e:
                print(f"Error in DINOv3 pipeline for {video_id}: {e}")
                traceback.print_exc()

    async def _finish_video(self, video_id: str, video_data: dict, embedding_data: Dict[str, Any]) -> None:
        """main.py:203-277: clip mean, neighbour evidence, VectorDB upsert, results JSON, `pipeline.dinov3` publish."""
        if embedding_data["embeddings"]:
            avg_embedding = np.mean([np.array(e["embedding"]) for e in embedding_data["embeddings"]], axis=0)
        else:
            print(f"No embeddings extracted for {video_id}")
            return
        similar_cases = self.search_similar(avg_embedding, top_k=5)
        if similar_cases:
            labels = [case["label"] for case in similar_cases if case["label"] is not None]
            if labels:
                lame_count = sum(1 for label in labels if label == 1)
                neighbor_evidence = lame_count / len(labels)
            else:
                neighbor_evidence = 0.5
        else:
            neighbor_evidence = 0.5
        payload = {
            "video_id": video_id,
            "filename": video_data.get("filename", ""),
            "uploaded_at": video_data.get("uploaded_at", ""),
            "label": None,
            "metadata": video_data.get("metadata", {}),
        }
        try:
            if self.gallery is not None:
                self.gallery.upsert(video_id, avg_embedding, payload)
            if self.qdrant_client is not None:
                point = _point_struct(id=video_id, vector=avg_embedding.tolist(), payload=payload)
                self.qdrant_client.upsert(collection_name=self.collection_name, points=[point])
            print(f"Stored embedding in VectorDB for {video_id}")
        except Exception as e:
            print(f"Error storing in VectorDB: {e}")
        results = {
            "video_id": video_id,
            "embedding_dim": len(avg_embedding),
            "num_embeddings": len(embedding_data["embeddings"]),
            "similar_cases": similar_cases,
            "neighbor_evidence": neighbor_evidence,
            "canonical_frames": embedding_data["canonical_frames"],
        }
        if self.emit_embedding:
            results["embedding"] = avg_embedding.tolist()
        results_file = self.results_dir / f"{video_id}_dinov3.json"
        with open(results_file, "w") as f:
            json.dump(results, f, indent=2)
        pipeline_result = {
            "video_id": video_id,
            "pipeline": "dinov3",
            "results_path": str(results_file),
            "neighbor_evidence": neighbor_evidence,
            "similar_cases": similar_cases,
            "embedding_dim": len(avg_embedding),
        }
        await self.nats_client.publish(self.config["nats"]["subjects"]["pipeline_dinov3"], pipeline_result)
        print(f"DINOv3 pipeline completed for {video_id}")

    # -- main.py:284-296 --------------------------------------------------------------------------
    async def start(self, coalesce: bool = False, max_batch: int = 64):
        """Subscribe to `video.preprocessed` and serve forever (main.py:284-296).  ``coalesce=True`` queues incoming messages and
        hands everything that has accumulated (up to ``max_batch``) to ``process_videos`` -- same outputs, one GPU batch."""
        import asyncio

        await self.nats_client.connect()
        subject = self.config["nats"]["subjects"]["video_preprocessed"]
        print(f"DINOv3 pipeline subscribed to {subject}")
        if not coalesce:
            await self.nats_client.subscribe(subject, self.process_video)
            print("DINOv3 pipeline service started. Waiting for videos...")
            await asyncio.Event().wait()
            return
        queue: "asyncio.Queue[dict]" = asyncio.Queue()

        async def enqueue(video_data: dict):
            video_data["video_id"], video_data["processed_path"]     # same KeyError as the reference handler, inside the callback
            await queue.put(video_data)

        await self.nats_client.subscribe(subject, enqueue)
        print("DINOv3 pipeline service started (coalescing). Waiting for videos...")
        while True:
            await self.drain_once(queue, max_batch)

    async def drain_once(self, queue, max_batch: int = 64) -> int:
        """Wait for one message, take everything else that is already queued (up to max_batch) and process it as one batch."""
        batch = [await queue.get()]
        while len(batch) < max_batch and not queue.empty():
            batch.append(queue.get_nowait())
        await self.process_videos(batch)
        return len(batch)


def build_pipeline_from_hf(model, **kw) -> DINOv3Pipeline:
    """Convenience: HF DINOv3ViTModel (already loaded by the caller) -> B200 pipeline."""
    max_frames = kw.pop("max_frames", 256)
    resize = kw.pop("resize", (224, 224))
    cfg = VitConfig.from_hf(model.config)
    engine = ClipEmbedEngine(cfg, model.state_dict(), max_frames=max_frames, resize=resize)
    return DINOv3Pipeline(engine, **kw)
