"""Device-side engine: owns the packed weights, the workspace and the cre_ctx, and sequences the
C-ABI calls for   uint8 frames -> patches -> ViT -> frame embeddings -> clip embeddings -> re-ID top-k.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed only; every arithmetic
step is a libcre_b200 kernel.  Reference call sites replaced (relative to the reference tree):
services/dinov3-pipeline/app/main.py:95-115 (extract_embedding), :204-208 (clip mean),
services/tracking-service/app/reid/matcher.py:124-132 (normalise + cosine top-k).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # HF DINOv3ViTImageProcessor defaults
IMAGENET_STD = (0.229, 0.224, 0.225)


@dataclass(frozen=True)
class VitConfig:
    """The DINOv3ViTConfig fields the forward pass depends on (HF:configuration_dinov3_vit.py:74-101)."""

    hidden: int = 768
    layers: int = 12
    heads: int = 12
    mlp: int = 3072
    patch: int = 16
    registers: int = 4
    rope_theta: float = 100.0
    ln_eps: float = 1e-5

    @staticmethod
    def vit_b16() -> "VitConfig":
        return VitConfig()

    @staticmethod
    def vit_l16() -> "VitConfig":
        return VitConfig(hidden=1024, layers=24, heads=16, mlp=4096)

    @staticmethod
    def from_hf(cfg) -> "VitConfig":
        if getattr(cfg, "model_type", "dinov3_vit") != "dinov3_vit":
            raise ValueError(f"only dinov3_vit checkpoints are supported, got model_type={cfg.model_type!r}")
        if getattr(cfg, "use_gated_mlp", False):
            raise ValueError("gated-MLP DINOv3 variants are not supported")
        if cfg.hidden_act != "gelu":
            raise ValueError(f"hidden_act={cfg.hidden_act!r} unsupported (exact-erf gelu only)")
        if cfg.hidden_size // cfg.num_attention_heads != 64:
            raise ValueError("head_dim must be 64")
        return VitConfig(
            hidden=cfg.hidden_size,
            layers=cfg.num_hidden_layers,
            heads=cfg.num_attention_heads,
            mlp=cfg.intermediate_size,
            patch=cfg.patch_size,
            registers=cfg.num_register_tokens,
            rope_theta=float(cfg.rope_theta),
            ln_eps=float(cfg.layer_norm_eps),
        )

    def c_struct(self) -> _lib.ModelCfg:
        return _lib.ModelCfg(self.hidden, self.layers, self.heads, self.mlp, self.patch, self.registers,
                             self.rope_theta, self.ln_eps)

    @property
    def prefix(self) -> int:
        return 1 + self.registers

    def flops_per_frame(self, grid_h: int, grid_w: int) -> float:
        """Algorithmic FLOPs (2MNK per GEMM, 4 T^2 D per layer of attention), SURVEY.md section 8(d)."""
        p = grid_h * grid_w
        t = p + self.prefix
        d, f = self.hidden, self.mlp
        return 2.0 * p * d * 3 * self.patch * self.patch + self.layers * (8.0 * t * d * d + 4.0 * t * d * f + 4.0 * t * t * d)


def _get(sd: Mapping[str, torch.Tensor], *names: str) -> torch.Tensor:
    for n in names:
        if n in sd:
            return sd[n]
    raise KeyError(f"none of {names} in state_dict")


def pack_weights(cfg: VitConfig, state_dict: Mapping[str, torch.Tensor]) -> torch.Tensor:
    """HF DINOv3ViTModel.state_dict() -> one uint8 host blob in the layout cre_weight_offset() defines
    (bf16 matrices [out, in], fp32 vectors; q/k/v fused into one [3D, D] matrix, zero k bias)."""
    lib = _lib.load()
    cs = cfg.c_struct()
    total = _lib.check_size(lib.cre_packed_weights_bytes(C.byref(cs)), "cre_packed_weights_bytes")
    blob = torch.zeros(total, dtype=torch.uint8)

    def put(layer: int, kind: int, t: torch.Tensor) -> None:
        off = _lib.check_size(lib.cre_weight_offset(C.byref(cs), layer, kind), "cre_weight_offset")
        n = _lib.check_size(lib.cre_weight_elems(C.byref(cs), layer, kind), "cre_weight_elems")
        t = t.detach().to("cpu").contiguous().reshape(-1)
        if t.numel() != n:
            raise ValueError(f"weight kind {kind} layer {layer}: {t.numel()} elements, expected {n}")
        if kind in _lib.BF16_KINDS:
            raw = t.to(torch.bfloat16).view(torch.uint8)
        else:
            raw = t.to(torch.float32).view(torch.uint8)
        blob[off:off + raw.numel()] = raw

    sd = state_dict
    d = cfg.hidden
    put(-1, _lib.W_PATCH, _get(sd, "embeddings.patch_embeddings.weight").reshape(d, -1))
    put(-1, _lib.B_PATCH, _get(sd, "embeddings.patch_embeddings.bias"))
    put(-1, _lib.PREFIX, torch.cat([_get(sd, "embeddings.cls_token").reshape(1, d),
                                    _get(sd, "embeddings.register_tokens").reshape(-1, d)], dim=0))
    put(-1, _lib.LN_F_G, _get(sd, "norm.weight"))
    put(-1, _lib.LN_F_B, _get(sd, "norm.bias"))
    for i in range(cfg.layers):
        def g(name: str) -> torch.Tensor:
            return _get(sd, f"model.layer.{i}.{name}", f"layer.{i}.{name}")

        def gopt(name: str, n: int) -> torch.Tensor:
            for k in (f"model.layer.{i}.{name}", f"layer.{i}.{name}"):
                if k in sd:
                    return sd[k]
            return torch.zeros(n)

        put(i, _lib.LN1_G, g("norm1.weight"))
        put(i, _lib.LN1_B, g("norm1.bias"))
        put(i, _lib.W_QKV, torch.cat([g("attention.q_proj.weight"), g("attention.k_proj.weight"),
                                      g("attention.v_proj.weight")], dim=0))
        put(i, _lib.B_QKV, torch.cat([gopt("attention.q_proj.bias", d), gopt("attention.k_proj.bias", d),
                                      gopt("attention.v_proj.bias", d)], dim=0))
        put(i, _lib.W_O, g("attention.o_proj.weight"))
        put(i, _lib.B_O, gopt("attention.o_proj.bias", d))
        put(i, _lib.LS1, g("layer_scale1.lambda1"))
        put(i, _lib.LN2_G, g("norm2.weight"))
        put(i, _lib.LN2_B, g("norm2.bias"))
        put(i, _lib.W_UP, g("mlp.up_proj.weight"))
        put(i, _lib.B_UP, gopt("mlp.up_proj.bias", cfg.mlp))
        put(i, _lib.W_DOWN, g("mlp.down_proj.weight"))
        put(i, _lib.B_DOWN, gopt("mlp.down_proj.bias", d))
        put(i, _lib.LS2, g("layer_scale2.lambda1"))
    return blob


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class ClipEmbedEngine:
    """One engine per process / GPU.  Not thread-safe (one cre_ctx, one workspace)."""

    def __init__(self, cfg: VitConfig, state_dict: Mapping[str, torch.Tensor], device: Optional[int] = None,
                 max_frames: int = 256, resize: Tuple[int, int] = (224, 224),
                 mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD):
        if not torch.cuda.is_available():
            raise _lib.CreError("ClipEmbedEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.resize = (int(resize[0]), int(resize[1]))
        self.grid = (self.resize[0] // cfg.patch, self.resize[1] // cfg.patch)
        self.tokens = self.grid[0] * self.grid[1] + cfg.prefix
        self.max_frames = int(max_frames)
        self._mean = (C.c_float * 3)(*mean)
        self._std = (C.c_float * 3)(*std)
        self._cs = cfg.c_struct()
        blob = pack_weights(cfg, state_dict)
        with torch.cuda.device(self.device):
            self.weights = blob.to(self.device)
            ctx = C.c_void_p()
            _lib.check(self.lib.cre_create(C.byref(self._cs), self.weights.data_ptr(), self.device_index, C.byref(ctx)),
                       "cre_create")
            self._ctx = ctx
            self._ws_bytes = _lib.check_size(
                self.lib.cre_workspace_bytes(C.byref(self._cs), self.max_frames, self.grid[0], self.grid[1]),
                "cre_workspace_bytes")
            self.workspace = torch.empty(self._ws_bytes, dtype=torch.uint8, device=self.device)
            self.patches = torch.empty((self.max_frames * self.grid[0] * self.grid[1], 3 * cfg.patch * cfg.patch),
                                       dtype=torch.bfloat16, device=self.device)
        self._gallery_scratch: Optional[torch.Tensor] = None
        self._staging = None

    # ------------------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None:
            self.lib.cre_destroy(self._ctx)
            self._ctx = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------------------------------
    def preprocess(self, frames: torch.Tensor, bgr: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 [n, H, W, 3] on the device -> bf16 patch rows [n * P, 768] (K1)."""
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_cuda:
            raise ValueError("frames must be a CUDA uint8 tensor [n, H, W, 3]")
        if frames.stride(3) != 1 or frames.stride(2) != 3:
            frames = frames.contiguous()
        n, h, w, _ = frames.shape
        p = self.grid[0] * self.grid[1]
        if out is None:
            if n > self.max_frames:
                raise ValueError(f"{n} frames > max_frames={self.max_frames}")
            out = self.patches[: n * p]
        _lib.check(self.lib.cre_preprocess_patchify(
            self._ctx, frames.data_ptr(), n, h, w, frames.stride(1), frames.stride(0), 1 if bgr else 0,
            self.resize[0], self.resize[1], self._mean, self._std, out.data_ptr(), self._stream()),
            "cre_preprocess_patchify")
        return out

    def preprocess_rois(self, frames: torch.Tensor, rois, bgr: bool = True) -> torch.Tensor:
        """uint8 [n, H, W, 3] on the device + boxes [(frame, x0, y0, x1, y1), ...] -> bf16 patch rows [n_rois * P, 768]: every box is
        cropped and resized exactly as the image frames[frame][y0:y1, x0:x1] would be (K1 in region-of-interest mode)."""
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_cuda:
            raise ValueError("frames must be a CUDA uint8 tensor [n, H, W, 3]")
        if frames.stride(3) != 1 or frames.stride(2) != 3:
            frames = frames.contiguous()
        n, h, w, _ = frames.shape
        boxes = np.asarray(rois, dtype=np.int64).reshape(-1, 5)
        if boxes.shape[0] == 0:
            raise ValueError("no regions of interest")
        if boxes.shape[0] > self.max_frames:
            raise ValueError(f"{boxes.shape[0]} regions > max_frames={self.max_frames}")
        f, x0, y0, x1, y1 = boxes.T
        if ((f < 0) | (f >= n) | (x0 < 0) | (y0 < 0) | (x1 > w) | (y1 > h) | (x1 <= x0) | (y1 <= y0)).any():
            raise ValueError("regions must be non-empty boxes inside their frame: (frame, x0, y0, x1, y1), x1 / y1 exclusive")
        rois_dev = torch.from_numpy(boxes.astype(np.int32)).to(self.device)
        need = _lib.check_size(self.lib.cre_roi_scratch_bytes(boxes.shape[0], h, w, self.resize[0], self.resize[1]), "cre_roi_scratch_bytes")
        if getattr(self, "_roi_scratch", None) is None or self._roi_scratch.numel() < need:
            self._roi_scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = self.patches[: boxes.shape[0] * self.grid[0] * self.grid[1]]
        _lib.check(self.lib.cre_preprocess_patchify_roi(
            self._ctx, frames.data_ptr(), n, h, w, frames.stride(1), frames.stride(0), 1 if bgr else 0, rois_dev.data_ptr(),
            boxes.shape[0], self.resize[0], self.resize[1], self._mean, self._std, self._roi_scratch.data_ptr(),
            self._roi_scratch.numel(), out.data_ptr(), self._stream()), "cre_preprocess_patchify_roi")
        return out

    def embed_rois(self, frames: torch.Tensor, rois, bgr: bool = True) -> torch.Tensor:
        """Per-box embeddings f32 [n_rois, D] (crop -> HF-processor resize -> ViT -> token mean), chunked by max_frames."""
        boxes = np.asarray(rois, dtype=np.int64).reshape(-1, 5)
        out = torch.empty((boxes.shape[0], self.cfg.hidden), dtype=torch.float32, device=self.device)
        for s in range(0, boxes.shape[0], self.max_frames):
            chunk = boxes[s:s + self.max_frames]
            patches = self.preprocess_rois(frames, chunk, bgr=bgr)
            self.forward_patches(patches, chunk.shape[0], out=out[s:s + chunk.shape[0]])
        return out

    def forward_patches(self, patches: torch.Tensor, n: int, want_tokens: bool = False,
                        out: Optional[torch.Tensor] = None):
        """bf16 patch rows -> f32 frame embeddings [n, D] (final LayerNorm + mean over all tokens)."""
        if n > self.max_frames:
            raise ValueError(f"{n} frames > max_frames={self.max_frames}")
        emb = out if out is not None else torch.empty((n, self.cfg.hidden), dtype=torch.float32, device=self.device)
        if emb.shape != (n, self.cfg.hidden) or emb.dtype != torch.float32 or not emb.is_contiguous():
            raise ValueError("out must be a contiguous f32 [n, hidden] tensor")
        tokens = (torch.empty((n, self.tokens, self.cfg.hidden), dtype=torch.float32, device=self.device)
                  if want_tokens else None)
        _lib.check(self.lib.cre_vit_forward(
            self._ctx, patches.data_ptr(), n, self.grid[0], self.grid[1], self.workspace.data_ptr(), self._ws_bytes,
            emb.data_ptr(), _ptr(tokens), self._stream()), "cre_vit_forward")
        return (emb, tokens) if want_tokens else emb

    def embed_frames(self, frames: torch.Tensor, bgr: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 [n, H, W, 3] (device) -> f32 [n, D]; processed in chunks of max_frames on the current stream."""
        n = frames.shape[0]
        if out is None:
            out = torch.empty((n, self.cfg.hidden), dtype=torch.float32, device=self.device)
        for s in range(0, n, self.max_frames):
            chunk = frames[s:s + self.max_frames]
            patches = self.preprocess(chunk, bgr=bgr)
            self.forward_patches(patches, chunk.shape[0], out=out[s:s + chunk.shape[0]])
        return out

    def embed_host_frames(self, frames, bgr: bool = True, out: Optional[torch.Tensor] = None, copy_only: bool = False) -> torch.Tensor:
        """HOST uint8 frames -> DEVICE f32 [n, D].  ``frames``: one [n, H, W, 3] numpy array / CPU tensor, or a list of
        them (e.g. one per clip; all the same H x W), embedded as if concatenated.

        Batches of max_frames are copied host->device on a side stream while the previous batch is being
        embedded (two device frame buffers, events both ways).  Pinned CPU tensors are copied in place;
        pageable input is staged through two pinned buffers first.  Returns after queueing; the caller
        synchronises by reading the result (``.cpu()``) or on the current stream.  A pinned staging slot is never
        rewritten before the copy that last read it has finished -- also across calls (back-to-back calls with pageable input).
        ``copy_only=True`` (bench.py's H2D probe) issues exactly the same copies, events and stream waits but launches no kernel."""
        srcs = list(frames) if isinstance(frames, (list, tuple)) else [frames]
        srcs = [torch.from_numpy(np.ascontiguousarray(f)) if isinstance(f, np.ndarray) else f for f in srcs]
        if len(srcs) == 1 and srcs[0].is_cuda:
            return self.embed_frames(srcs[0], bgr=bgr, out=out)
        for f in srcs:
            if f.is_cuda or f.dtype != torch.uint8 or f.dim() != 4 or f.shape[-1] != 3 or f.shape[1:] != srcs[0].shape[1:]:
                raise ValueError("frames must be host uint8 [n, H, W, 3] arrays of one common H x W")
        srcs = [f.contiguous() for f in srcs if f.shape[0] > 0]
        n = sum(f.shape[0] for f in srcs)
        if out is None:
            out = torch.empty((n, self.cfg.hidden), dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        h, w = int(srcs[0].shape[1]), int(srcs[0].shape[2])
        per = h * w * 3
        cf = min(self.max_frames, n)
        st = self._staging
        if st is None or st["dev"][0].numel() < cf * per:
            st = {"dev": [torch.empty(cf * per, dtype=torch.uint8, device=self.device) for _ in range(2)],
                  "pin": [None, None],
                  "copied": [torch.cuda.Event() for _ in range(2)], "freed": [torch.cuda.Event() for _ in range(2)],
                  "stream": torch.cuda.Stream(device=self.device)}
            self._staging = st
        all_pinned = all(f.is_pinned() for f in srcs)
        if not all_pinned and (st["pin"][0] is None or st["pin"][0].numel() < cf * per):
            st["pin"] = [torch.empty(cf * per, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        starts = np.cumsum([0] + [f.shape[0] for f in srcs])
        compute = torch.cuda.current_stream(self.device)
        st["stream"].wait_stream(compute)
        for i, s in enumerate(range(0, n, cf)):
            b = i & 1
            m = min(cf, n - s)
            dev = st["dev"][b]
            if not all_pinned:
                # this pinned slot's previous H2D -- of this call (i >= 2) or of an EARLIER call that returned after queueing --
                # must have finished before the host overwrites it; an event that was never recorded is complete
                st["copied"][b].synchronize()
            with torch.cuda.stream(st["stream"]):
                if i >= 2:
                    st["stream"].wait_event(st["freed"][b])
                k = int(np.searchsorted(starts, s, side="right")) - 1
                pos = s
                while pos < s + m:                     # one copy per source segment overlapping this batch
                    lo = pos - int(starts[k])
                    cnt = min(srcs[k].shape[0] - lo, s + m - pos)
                    src = srcs[k][lo:lo + cnt].reshape(-1)
                    if not all_pinned:
                        stage = st["pin"][b][(pos - s) * per:(pos - s + cnt) * per]
                        stage.copy_(src)
                        src = stage
                    dev[(pos - s) * per:(pos - s + cnt) * per].copy_(src, non_blocking=True)
                    pos += cnt
                    k += 1
                st["copied"][b].record(st["stream"])
            compute.wait_event(st["copied"][b])
            if not copy_only:
                self.embed_frames(dev[: m * per].view(m, h, w, 3), bgr=bgr, out=out[s:s + m])
            st["freed"][b].record(compute)
        return out

    def pool_clips(self, frame_emb: torch.Tensor, clip_offsets: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """f32 [F, D] + int32 offsets [Q + 1] -> (raw clip mean [Q, D], unit-norm [Q, D])  (K3b)."""
        q = clip_offsets.numel() - 1
        d = frame_emb.shape[1]
        offs = clip_offsets.to(device=self.device, dtype=torch.int32).contiguous()
        mean = torch.empty((q, d), dtype=torch.float32, device=self.device)
        unit = torch.empty((q, d), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cre_pool_clips(frame_emb.contiguous().data_ptr(), offs.data_ptr(), q, d, mean.data_ptr(),
                                           unit.data_ptr(), self._stream()), "cre_pool_clips")
        return mean, unit

    def gallery_topk(self, queries: torch.Tensor, gallery: torch.Tensor, k: int = 5, row_base: int = 0,
                     dump_scores: bool = False):
        """f32 unit queries [Q, D] x bf16 unit gallery [N, D] -> (scores [Q, k], idx [Q, k]) (K4)."""
        if gallery.dtype != torch.bfloat16 or not gallery.is_contiguous():
            raise ValueError("gallery must be a contiguous bf16 tensor [N, D]")
        queries = queries.to(torch.float32).contiguous()
        q, d = queries.shape
        rows = gallery.shape[0]
        need = _lib.check_size(self.lib.cre_gallery_scratch_bytes(q, d, k), "cre_gallery_scratch_bytes")
        if self._gallery_scratch is None or self._gallery_scratch.numel() < need:
            self._gallery_scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        scores = torch.empty((q, k), dtype=torch.float32, device=self.device)
        idx = torch.empty((q, k), dtype=torch.int32, device=self.device)
        dump = torch.empty((q, rows), dtype=torch.float32, device=self.device) if dump_scores else None
        _lib.check(self.lib.cre_gallery_topk(
            self._ctx, queries.data_ptr(), q, d, gallery.data_ptr() if rows else None, rows, row_base, k,
            self._gallery_scratch.data_ptr(), self._gallery_scratch.numel(), scores.data_ptr(), idx.data_ptr(),
            _ptr(dump), self._stream()), "cre_gallery_topk")
        return (scores, idx, dump) if dump_scores else (scores, idx)

    def merge_topk(self, scores: torch.Tensor, idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """[lists, Q, k] candidate lists -> [Q, k] under (score desc, index asc)."""
        lists, q, k = scores.shape
        scores = scores.contiguous()
        idx = idx.to(torch.int32).contiguous()
        out_s = torch.empty((q, k), dtype=torch.float32, device=self.device)
        out_i = torch.empty((q, k), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.cre_merge_topk(scores.data_ptr(), idx.data_ptr(), lists, q, k, out_s.data_ptr(),
                                           out_i.data_ptr(), self._stream()), "cre_merge_topk")
        return out_s, out_i

    def gallery_update_row(self, gallery: torch.Tensor, row: int, unit_query: torch.Tensor, momentum: float,
                           master: Optional[torch.Tensor] = None) -> None:
        """One gallery row <- normalise(momentum * old + (1 - momentum) * unit_query); ``master`` f32 [N, D] is the full-precision
        copy (read for ``old``, written), ``gallery`` bf16 [N, D] the scan copy."""
        if master is not None and (master.dtype != torch.float32 or master.shape != gallery.shape or not master.is_contiguous()):
            raise ValueError("master must be a contiguous f32 tensor of the gallery's shape")
        _lib.check(self.lib.cre_gallery_update_row(gallery.data_ptr(), _ptr(master), gallery.shape[1], int(row),
                                                   unit_query.to(torch.float32).contiguous().data_ptr(),
                                                   float(momentum), self._stream()), "cre_gallery_update_row")

    # ---- building blocks (parity tests) ------------------------------------------------------
    def gemm(self, a: torch.Tensor, b: torch.Tensor, epilogue: int = _lib.EPI_F32, bias=None, scale=None,
             out: Optional[torch.Tensor] = None, cta_group: int = 1) -> torch.Tensor:
        m, k = a.shape
        n = b.shape[0]
        if out is None:
            dt = torch.float32 if epilogue in (_lib.EPI_F32, _lib.EPI_RESID) else torch.bfloat16
            out = torch.zeros((m, n), dtype=dt, device=self.device)
        _lib.check(self.lib.cre_gemm_bf16(self._ctx, a.data_ptr(), b.data_ptr(), m, n, k, epilogue, _ptr(bias),
                                          _ptr(scale), out.data_ptr(), cta_group, self._stream()), "cre_gemm_bf16")
        return out

    def row_stats(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """f32 [rows, dim] -> (bf16 [rows, dim] = x - row mean, f32 statistics rows [rows, 2 * dim / 128 + 4])."""
        rows, dim = x.shape
        xb = torch.empty((rows, dim), dtype=torch.bfloat16, device=self.device)
        stats = torch.zeros((rows, 2 * (dim // 128) + 4), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cre_row_stats(x.data_ptr(), rows, dim, xb.data_ptr(), stats.data_ptr(), self._stream()), "cre_row_stats")
        return xb, stats

    def row_stats_split(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """f32 [rows, dim] -> (hi = bf16(x - row mean), lo = bf16(x - row mean - hi), statistics rows): the split residual stream."""
        rows, dim = x.shape
        hi = torch.empty((rows, dim), dtype=torch.bfloat16, device=self.device)
        lo = torch.empty_like(hi)
        stats = torch.zeros((rows, 2 * (dim // 128) + 4), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cre_row_stats_split(x.data_ptr(), rows, dim, hi.data_ptr(), lo.data_ptr(), stats.data_ptr(), self._stream()),
                   "cre_row_stats_split")
        return hi, lo, stats

    def fold_ln_weights(self, w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: Optional[torch.Tensor] = None,
                        scaled_rows: int = 0, row_scale: float = 1.0):
        """bf16 W [n, k] + LayerNorm (gamma, beta) [k] + bias [n] -> (bf16 W * gamma, c1 = its row sums, c2 = bias + W beta)."""
        n, k = w.shape
        wf = torch.empty_like(w)
        c1 = torch.empty(n, dtype=torch.float32, device=self.device)
        c2 = torch.empty(n, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cre_fold_ln_weights(w.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(bias), n, k, int(scaled_rows), float(row_scale),
                                                wf.data_ptr(),
                                                c1.data_ptr(), c2.data_ptr(), self._stream()), "cre_fold_ln_weights")
        return wf, c1, c2

    def gemm_ln(self, a: torch.Tensor, b: torch.Tensor, epilogue: int, stats_in: torch.Tensor, ln_dim: int, bias=None, c1=None,
                scale=None, out: Optional[torch.Tensor] = None, eps: float = 1e-5, cta_group: int = 2):
        """LayerNorm-folded GEMM building block (include/cre.h cre_gemm_ln).  EPI_BF16 / EPI_GELU: returns bf16 [m, n].
        EPI_RESID_LN: `out` f32 [m, n] is updated in place; returns (out, bf16 out - pivot, new statistics rows).
        EPI_RESID_SP: `out` = (hi, lo) bf16 [m, n] halves of the residual stream, updated in place; returns (hi, lo, new rows)."""
        m, k = a.shape
        n = b.shape[0]
        if epilogue in (_lib.EPI_RESID_LN, _lib.EPI_RESID_LN3):
            xb = torch.empty((m, n), dtype=torch.bfloat16, device=self.device)
            stats_out = torch.zeros_like(stats_in)
            _lib.check(self.lib.cre_gemm_ln(self._ctx, a.data_ptr(), b.data_ptr(), m, n, k, epilogue, _ptr(bias), None, _ptr(scale),
                                            stats_in.data_ptr(), ln_dim, eps, out.data_ptr(), xb.data_ptr(), stats_out.data_ptr(),
                                            cta_group, self._stream()), "cre_gemm_ln")
            return out, xb, stats_out
        if epilogue in (_lib.EPI_RESID_SP, _lib.EPI_RESID_SP3):
            hi, lo = out                                       # both halves are updated in place
            stats_out = torch.zeros_like(stats_in)
            _lib.check(self.lib.cre_gemm_ln(self._ctx, a.data_ptr(), b.data_ptr(), m, n, k, epilogue, _ptr(bias), None, _ptr(scale),
                                            stats_in.data_ptr(), ln_dim, eps, lo.data_ptr(), hi.data_ptr(), stats_out.data_ptr(),
                                            cta_group, self._stream()), "cre_gemm_ln")
            return hi, lo, stats_out
        if out is None:
            out = torch.zeros((m, n), dtype=torch.bfloat16, device=self.device)
        _lib.check(self.lib.cre_gemm_ln(self._ctx, a.data_ptr(), b.data_ptr(), m, n, k, epilogue, _ptr(bias), _ptr(c1), None,
                                        stats_in.data_ptr(), ln_dim, eps, out.data_ptr(), None, None, cta_group, self._stream()),
                   "cre_gemm_ln")
        return out

    def layernorm(self, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
        rows, dim = x.shape
        out = torch.empty((rows, dim), dtype=torch.bfloat16, device=self.device)
        _lib.check(self.lib.cre_layernorm_bf16(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rows, dim, eps,
                                               out.data_ptr(), self._stream()), "cre_layernorm_bf16")
        return out

    def attention(self, qkv: torch.Tensor, n: int, t: int, heads: int) -> torch.Tensor:
        """qkv bf16 [n*t, 3*heads*64] (q pre-scaled, rotary applied) -> bf16 [n*t, heads*64]."""
        d = heads * 64
        out = torch.empty((n * t, d), dtype=torch.bfloat16, device=self.device)
        need = _lib.check_size(self.lib.cre_attention_scratch_bytes(n, heads), "cre_attention_scratch_bytes")
        scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.cre_attention(self._ctx, qkv.data_ptr(), qkv.shape[1], d, 2 * d, n, t, heads, out.data_ptr(),
                                          scratch.data_ptr(), need, self._stream()), "cre_attention")
        return out


def set_cta_group(cg: int) -> None:
    _lib.check(_lib.load().cre_set_cta_group(int(cg)), "cre_set_cta_group")
