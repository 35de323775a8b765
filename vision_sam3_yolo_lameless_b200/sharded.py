"""Multi-GPU layout of the hot path: one process per GPU, clips sharded over ranks, gallery sharded by row.

The reference has no multi-GPU path (single process pinned to CUDA_VISIBLE_DEVICES=0,
docker-compose.gpu.yml:126-127); this is the partitioning BASELINE.json's north_star prescribes:

* embedding: clip c is owned by rank ``c % world`` -- pure data parallel, weights replicated, no collective;
* re-ID: rank r holds gallery rows ``[row_lo(r), row_hi(r))``.  Two small exchanges per batch:
    1. all-gather of the per-rank unit query embeddings  [Q_local, D] f32  -> every rank has all Q queries
    2. each rank scores all Q queries against its shard (K4, cre_gallery_topk, global indices via row_base)
    3. all-gather of the per-shard candidates (score f32, index i32) [Q, k]  -> cre_merge_topk under the total
       order (score desc, index asc); the result is identical on every rank.

``torch.distributed`` is plumbing only (NCCL on GPUs; the CPU tests drive the same code over gloo with the two
kernel calls replaced by injected callables).  No arithmetic happens here.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n rows: the first n % world ranks get one extra row."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def bind_host_to_gpu(device_index: int) -> Optional[str]:
    """Pin the calling process to the CPU cores NVML reports as local to ``cuda:device_index`` (its NUMA node).

    With one process per GPU on a two-socket box, pinned staging buffers otherwise land on whatever node the process was started
    on, and half of the host->device copies cross the socket interconnect: the end-to-end path (1080p frames, 6.2 MB each) is
    PCIe-bound, so this is what decides its N-GPU scaling.  Call it BEFORE allocating pinned memory (first touch decides the
    node).  Returns a description of what was done, or None when NVML / the affinity call is unavailable (never raises)."""
    try:
        import os

        import pynvml

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        cpus = sorted(os.sched_getaffinity(0))
        return f"cuda:{device_index} -> {len(cpus)} local cpus [{cpus[0]}..{cpus[-1]}]"
    except Exception:        # noqa: BLE001 -- best effort: no NVML, container without the capability, single-node box
        return None


def clips_of_rank(num_clips: int, rank: int, world: int) -> List[int]:
    """Round-robin clip ownership (clip c -> rank c % world)."""
    return list(range(rank, num_clips, world))


class ShardedReID:
    """Row-sharded cosine top-k.  ``local_topk(queries [Q, D] f32, k) -> (scores [Q, k] f32, idx [Q, k] i32)`` scores
    against THIS rank's shard and returns GLOBAL row indices; ``merge(scores [R, Q, k], idx [R, Q, k]) -> ([Q, k], [Q, k])``.
    Both default to the engine's kernels and must be supplied explicitly when there is no engine (CPU tests)."""

    def __init__(self, engine=None, gallery_shard: Optional[torch.Tensor] = None, row_base: int = 0,
                 group: Optional[dist.ProcessGroup] = None,
                 local_topk: Optional[Callable] = None, merge: Optional[Callable] = None):
        if engine is None and (local_topk is None or merge is None):
            raise RuntimeError("ShardedReID needs a ClipEmbedEngine (GPU kernels); there is no CPU fallback")
        self.engine = engine
        self.gallery_shard = gallery_shard
        self.row_base = int(row_base)
        self.group = group
        self._local_topk = local_topk
        self._merge = merge

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if dist.is_initialized() else 0

    # -- step 1 -----------------------------------------------------------------------------------
    def gather_queries(self, unit_queries_local: torch.Tensor, counts: Optional[List[int]] = None) -> torch.Tensor:
        """[Q_local, D] per rank -> [Q, D] in RANK-MAJOR order (rank 0's queries first).  ``counts`` = queries per
        rank when they differ (ragged tail); omitted = every rank holds the same number."""
        world = self.world
        if world == 1:
            return unit_queries_local
        q_local, d = unit_queries_local.shape
        if counts is None:
            out = unit_queries_local.new_empty((world * q_local, d))
            dist.all_gather_into_tensor(out, unit_queries_local.contiguous(), group=self.group)
            return out
        q_max = max(counts)
        padded = unit_queries_local.new_zeros((q_max, d))
        padded[:q_local] = unit_queries_local
        out = unit_queries_local.new_empty((world * q_max, d))
        dist.all_gather_into_tensor(out, padded, group=self.group)
        return torch.cat([out[r * q_max: r * q_max + counts[r]] for r in range(world)], dim=0)

    # -- steps 2 + 3 ------------------------------------------------------------------------------
    def search_all(self, unit_queries_all: torch.Tensor, k: int = 5) -> Tuple[torch.Tensor, torch.Tensor]:
        """All queries (already gathered) -> global top-k (scores [Q, k], idx [Q, k]); same result on every rank."""
        if self._local_topk is not None:
            s, i = self._local_topk(unit_queries_all, k)
        else:
            s, i = self.engine.gallery_topk(unit_queries_all, self.gallery_shard, k=k, row_base=self.row_base)
        world = self.world
        if world == 1:
            return s, i
        q = unit_queries_all.shape[0]
        all_s = s.new_empty((world * q, k))      # rank-major concatenation == [world, q, k]
        all_i = i.new_empty((world * q, k))
        dist.all_gather_into_tensor(all_s, s.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_i, i.contiguous(), group=self.group)
        all_s, all_i = all_s.view(world, q, k), all_i.view(world, q, k)
        if self._merge is not None:
            return self._merge(all_s, all_i)
        return self.engine.merge_topk(all_s, all_i)

    def search(self, unit_queries_local: torch.Tensor, k: int = 5, counts: Optional[List[int]] = None):
        return self.search_all(self.gather_queries(unit_queries_local, counts), k)
