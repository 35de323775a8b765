"""Drop-in replacement for the reference's re-ID matcher
(``services/tracking-service/app/reid/matcher.py`` class ``CowReIDMatcher``, lines 37-372) and the
``pipeline.dinov3`` handler part of the tracking service
(``services/tracking-service/app/main.py:268-381`` ``process_dinov3_results`` / ``_perform_reid``).

Same constructor arguments, thresholds, method names, return types and error behaviour.  The read
path (normalise -> cosine top-k, matcher.py:123-132) runs on the GPU-resident gallery
(:class:`GpuGallery`, kernels K3b/K4); the write path (create / momentum update, matcher.py:203-301)
updates the device rows with one kernel -- the fp32 master row is what gets blended and what is written
through to Qdrant when a client is attached (matcher.py:267-301); the bf16 row is only the scan copy.
Threshold / decision logic (matcher.py:144-201,303-311) stays on the host, unchanged.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Tuple
from uuid import UUID, uuid4

import numpy as np

from .engine import ClipEmbedEngine
from .gallery import GpuGallery, ShardedGpuGallery


@dataclass
class CowIdentity:
    """Represents a known cow identity (matcher.py:16-24)."""
    identity_id: UUID
    cow_id: str
    tag_number: Optional[str] = None
    total_sightings: int = 0
    embedding: Optional[np.ndarray] = None
    embedding_dim: int = 768


@dataclass
class ReIDMatch:
    """Result of a Re-ID query (matcher.py:27-34)."""
    identity_id: UUID
    cow_id: str
    similarity: float
    confidence: str
    is_new_identity: bool = False


class CowReIDMatcher:
    COLLECTION_NAME = "cow_identities"
    SIMILARITY_THRESHOLD_HIGH = 0.85
    SIMILARITY_THRESHOLD_MEDIUM = 0.75
    SIMILARITY_THRESHOLD_LOW = 0.65

    def __init__(self, qdrant_url: str = "http://qdrant:6333", embedding_dim: int = 768,
                 auto_create_identities: bool = True, embedding_momentum: float = 0.9,
                 engine: Optional[ClipEmbedEngine] = None, qdrant_client=None, sharded: bool = False, group=None):
        """``sharded=True`` (one process per GPU, torch.distributed initialised): the gallery is row-sharded over the ranks of
        ``group`` (:class:`ShardedGpuGallery`); every rank must then drive the matcher with the same calls in the same order."""
        self.sharded = sharded
        self.group = group
        self.qdrant_url = qdrant_url
        self.embedding_dim = embedding_dim
        self.auto_create_identities = auto_create_identities
        self.embedding_momentum = embedding_momentum
        self.engine = engine
        self.qdrant_client = qdrant_client  # optional durable store (write-through)
        self.client: Optional[GpuGallery] = None  # named `client` like the reference: None == not connected
        self.identity_counter = 0

    async def connect(self):
        """Create the device gallery; if a Qdrant client is attached, make sure the collection exists
        and mirror its points (matcher.py:80-102)."""
        if self.engine is None:
            raise RuntimeError("CowReIDMatcher needs a ClipEmbedEngine (no CPU fallback)")
        self.client = (ShardedGpuGallery(self.engine, self.embedding_dim, group=self.group) if self.sharded
                       else GpuGallery(self.engine, self.embedding_dim))
        if self.qdrant_client is not None:
            names = [c.name for c in self.qdrant_client.get_collections().collections]
            if self.COLLECTION_NAME not in names:
                self.qdrant_client.create_collection(collection_name=self.COLLECTION_NAME,
                                                     vectors_config={"size": self.embedding_dim, "distance": "Cosine"})
                print(f"Created Qdrant collection: {self.COLLECTION_NAME}")
            else:
                print(f"Using existing Qdrant collection: {self.COLLECTION_NAME}")
                self.client.load_from_qdrant(self.qdrant_client, self.COLLECTION_NAME)     # paged scroll, one bulk load
        self.identity_counter = len(self.client)

    # -- matcher.py:104-149 ----------------------------------------------------------------------
    def match_embedding(self, embedding: np.ndarray, top_k: int = 5) -> Tuple[Optional[ReIDMatch], List[ReIDMatch]]:
        if self.client is None:
            raise RuntimeError("Not connected to Qdrant. Call connect() first.")
        points = self.client.search(np.asarray(embedding, dtype=np.float32), top_k)  # normalises on the device
        candidates = []
        for point in points:
            candidates.append(ReIDMatch(identity_id=UUID(point.payload["identity_id"]), cow_id=point.payload["cow_id"],
                                        similarity=point.score, confidence=self._score_to_confidence(point.score)))
        best_match = None
        if candidates and candidates[0].similarity >= self.SIMILARITY_THRESHOLD_LOW:
            best_match = candidates[0]
        return best_match, candidates

    def match_embeddings(self, embeddings: np.ndarray, top_k: int = 5) -> List[Tuple[Optional[ReIDMatch], List[ReIDMatch]]]:
        """Batched read path: all queries scored against one gallery snapshot in a single K4 launch."""
        if self.client is None:
            raise RuntimeError("Not connected to Qdrant. Call connect() first.")
        out = []
        for points in self.client.search_batch(np.asarray(embeddings, dtype=np.float32), top_k):
            cands = [ReIDMatch(identity_id=UUID(p.payload["identity_id"]), cow_id=p.payload["cow_id"], similarity=p.score,
                               confidence=self._score_to_confidence(p.score)) for p in points]
            best = cands[0] if cands and cands[0].similarity >= self.SIMILARITY_THRESHOLD_LOW else None
            out.append((best, cands))
        return out

    # -- matcher.py:151-201 ----------------------------------------------------------------------
    def match_or_create(self, embedding: np.ndarray, video_id: str, track_id: int, metadata: Optional[Dict] = None) -> ReIDMatch:
        best_match, candidates = self.match_embedding(embedding)
        if best_match is not None and best_match.similarity >= self.SIMILARITY_THRESHOLD_MEDIUM:
            self._update_identity_embedding(best_match.identity_id, embedding)
            return best_match
        if self.auto_create_identities:
            new_identity = self.create_identity(
                embedding=embedding, tag_number=None,
                metadata={"first_video": video_id, "first_track": track_id, **(metadata or {})})
            return ReIDMatch(identity_id=new_identity.identity_id, cow_id=new_identity.cow_id, similarity=1.0,
                             confidence="high", is_new_identity=True)
        return ReIDMatch(identity_id=uuid4(), cow_id="UNKNOWN",
                         similarity=candidates[0].similarity if candidates else 0.0, confidence="low", is_new_identity=True)

    # -- matcher.py:203-255 ----------------------------------------------------------------------
    def create_identity(self, embedding: np.ndarray, tag_number: Optional[str] = None, metadata: Optional[Dict] = None) -> CowIdentity:
        if self.client is None:
            raise RuntimeError("Not connected to Qdrant. Call connect() first.")
        self.identity_counter += 1
        identity_id = uuid4()
        cow_id = f"COW-{self.identity_counter:04d}"
        payload = {"identity_id": str(identity_id), "cow_id": cow_id, "tag_number": tag_number, "total_sightings": 1,
                   **(metadata or {})}
        self.client.upsert(str(identity_id), embedding, payload)           # normalised on the device
        stored = self.client.vector(str(identity_id))
        self._write_through(str(identity_id), stored, payload)
        return CowIdentity(identity_id=identity_id, cow_id=cow_id, tag_number=tag_number, total_sightings=1,
                           embedding=stored, embedding_dim=len(stored))

    # -- matcher.py:257-301 ----------------------------------------------------------------------
    def _update_identity_embedding(self, identity_id: UUID, new_embedding: np.ndarray):
        if self.client is None:
            return
        row = self.client._row_of.get(str(identity_id))
        if row is None:
            return
        payload = dict(self.client.payloads[row])
        payload["total_sightings"] = payload.get("total_sightings", 0) + 1
        self.client.upsert(str(identity_id), new_embedding, payload, momentum=self.embedding_momentum)
        self._write_through(str(identity_id), self.client.vector(str(identity_id)), payload)

    def _write_through(self, point_id: str, vector: np.ndarray, payload: Dict) -> None:
        if self.qdrant_client is None:
            return
        from .extractor import _point_struct
        self.qdrant_client.upsert(collection_name=self.COLLECTION_NAME,
                                  points=[_point_struct(id=point_id, vector=vector.tolist(), payload=payload)])

    # -- matcher.py:303-311 ----------------------------------------------------------------------
    def _score_to_confidence(self, score: float) -> str:
        if score >= self.SIMILARITY_THRESHOLD_HIGH:
            return "high"
        elif score >= self.SIMILARITY_THRESHOLD_MEDIUM:
            return "medium"
        elif score >= self.SIMILARITY_THRESHOLD_LOW:
            return "low"
        return "none"

    # -- matcher.py:313-372 ----------------------------------------------------------------------
    def get_identity(self, identity_id: UUID) -> Optional[CowIdentity]:
        if self.client is None:
            return None
        row = self.client._row_of.get(str(identity_id))
        if row is None:
            return None
        payload = self.client.payloads[row]
        return CowIdentity(identity_id=UUID(payload["identity_id"]), cow_id=payload["cow_id"],
                           tag_number=payload.get("tag_number"), total_sightings=payload.get("total_sightings", 0),
                           embedding=self.client.vector(str(identity_id)), embedding_dim=self.embedding_dim)

    def get_all_identities(self, limit: int = 100) -> List[CowIdentity]:
        if self.client is None:
            return []
        return [CowIdentity(identity_id=UUID(p["identity_id"]), cow_id=p["cow_id"], tag_number=p.get("tag_number"),
                            total_sightings=p.get("total_sightings", 0), embedding_dim=self.embedding_dim)
                for p in self.client.payloads[:limit]]

    def get_statistics(self) -> dict:
        if self.client is None:
            return {"status": "disconnected"}
        return {"status": "connected", "collection": self.COLLECTION_NAME, "total_identities": len(self.client),
                "embedding_dim": self.embedding_dim, "similarity_threshold": self.SIMILARITY_THRESHOLD_MEDIUM}


class TrackingReIDHandler:
    """The re-ID half of the reference TrackingService (tracking main.py:268-381): consumes
    ``pipeline.dinov3`` messages, runs re-ID for the video's pending tracks, rewrites the tracking JSON and
    publishes ``tracking.reid.match``.  Postgres writes are delegated to an optional ``save_track`` coroutine
    (the reference's ``_save_track_to_db`` returns early without a session, tracking main.py:385-386)."""

    def __init__(self, reid_matcher: CowReIDMatcher, nats_client, results_dir: Path, subjects: Optional[dict] = None,
                 save_track=None):
        self.reid_matcher = reid_matcher
        self.nats_client = nats_client
        self.results_dir = Path(results_dir)
        self.subjects = subjects or {}
        self.save_track = save_track
        self.video_embeddings: Dict[str, np.ndarray] = {}
        self.pending_tracks: Dict[str, List[Dict]] = {}
        # opt-in: {video_id: {track_id: embedding}} from DINOv3Pipeline.extract_track_embeddings; a track found
        # here is re-identified on ITS crops' embedding instead of the whole-video embedding the reference reuses for every track
        self.track_embeddings: Dict[str, Dict[int, np.ndarray]] = {}

    # -- tracking main.py:268-320 ------------------------------------------------------------------
    async def process_dinov3_results(self, message: dict):
        video_id = message.get("video_id")
        if not video_id:
            return
        print(f"Tracking service processing DINOv3 results for {video_id}")
        try:
            results_path = message.get("results_path")
            embedding = None
            if results_path:
                results_file = Path(results_path)
                if results_file.exists():
                    with open(results_file) as f:
                        dinov3_data = json.load(f)
                    if "embedding" in dinov3_data:
                        embedding = np.array(dinov3_data["embedding"])
                    elif "canonical_frames" in dinov3_data and dinov3_data["canonical_frames"]:
                        frame_embeddings = [np.array(frame["embedding"]) for frame in dinov3_data["canonical_frames"]
                                            if "embedding" in frame]
                        if frame_embeddings:
                            embedding = np.mean(frame_embeddings, axis=0)
                    elif "video_embedding" in dinov3_data:
                        embedding = np.array(dinov3_data["video_embedding"])
            if embedding is None or len(embedding) == 0:
                print(f"  No embedding found for {video_id}")
                return
            self.video_embeddings[video_id] = embedding
            if video_id in self.pending_tracks:
                await self._perform_reid(video_id, embedding)
        except Exception as e:
            print(f"  Error processing DINOv3 for Re-ID: {e}")
            import traceback
            traceback.print_exc()

    # -- tracking main.py:322-381 ------------------------------------------------------------------
    async def _perform_reid(self, video_id: str, embedding: np.ndarray):
        tracks = self.pending_tracks.get(video_id, [])
        if not tracks:
            return
        print(f"Performing Re-ID for {len(tracks)} tracks in {video_id}")
        reid_results = []
        per_track = self.track_embeddings.pop(video_id, {})
        for track in tracks:
            match = self.reid_matcher.match_or_create(
                embedding=per_track.get(track["track_id"], embedding), video_id=video_id, track_id=track["track_id"],
                metadata={"start_frame": track["start_frame"], "end_frame": track["end_frame"]})
            reid_results.append({
                "track_id": track["track_id"], "cow_id": match.cow_id, "identity_id": str(match.identity_id),
                "similarity": match.similarity, "confidence": match.confidence, "is_new": match.is_new_identity})
            if self.save_track is not None:
                await self.save_track(video_id, track, match)
        results_file = self.results_dir / f"{video_id}_tracking.json"
        if results_file.exists():
            with open(results_file) as f:
                results = json.load(f)
            results["reid_results"] = reid_results
            results["reid_complete"] = True
            with open(results_file, "w") as f:
                json.dump(results, f, indent=2)
        await self.nats_client.publish(
            "tracking.reid.match",
            {"video_id": video_id, "matches": reid_results,
             "new_identities": sum(1 for r in reid_results if r["is_new"])})
        del self.pending_tracks[video_id]
        print(f"Re-ID complete for {video_id}: {len(reid_results)} tracks processed")
