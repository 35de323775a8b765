"""B200-native clip-embedding + re-ID hot path of UBC-AWP/vision-sam3-yolo-lameless.

Host side mirrors the reference's Python interfaces (``DINOv3Pipeline``, ``CowReIDMatcher``, the
``pipeline.dinov3`` handler); all arithmetic runs in ``libcre_b200.so`` (hand-written sm_100a CUDA behind the
C-ABI in ``include/cre.h``).  There is no CPU / PyTorch fallback: importing works anywhere, computing needs
the built library and a B200.
"""
from . import _lib
from .engine import ClipEmbedEngine, VitConfig, pack_weights, set_cta_group
from .extractor import DINOv3Pipeline, build_pipeline_from_hf
from .gallery import GpuGallery, ScoredPoint
from .knn_graph import GraphBuilder
from .reid import CowIdentity, CowReIDMatcher, ReIDMatch, TrackingReIDHandler

__all__ = ["_lib", "ClipEmbedEngine", "VitConfig", "pack_weights", "set_cta_group", "DINOv3Pipeline",
           "build_pipeline_from_hf", "GpuGallery", "ScoredPoint", "CowIdentity", "CowReIDMatcher", "ReIDMatch",
           "TrackingReIDHandler", "GraphBuilder"]
