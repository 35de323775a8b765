"""Multi-rank plumbing (sharded.py) over gloo, world_size 2 and 3, on CPU: all-gather of queries, per-shard top-k
(oracle callables stand in for the two kernels), all-gather + merge of candidates.  The result must equal the
single-shard top-k over the whole gallery, ties included, on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import reid_ref
from vision_sam3_yolo_lameless_b200.sharded import ShardedReID, clips_of_rank, shard_range


def test_shard_helpers():
    for n, w in [(100000, 8), (7, 3), (2, 4), (0, 2)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert clips_of_rank(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((clips_of_rank(10, r, 4) for r in range(4)), [])) == list(range(10))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ShardedReID(engine=None)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(q_total=11, n=501, d=64, seed=5):
    rng = np.random.default_rng(seed)
    g = reid_ref.l2_normalise(rng.standard_normal((n, d))).astype(np.float32)
    g[400] = g[3]
    g[77] = g[3]                       # exact ties across shards
    q = reid_ref.l2_normalise(rng.standard_normal((q_total, d))).astype(np.float32)
    q[0] = g[3]
    return q, g


def _worker(rank, world, port, ragged, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q, g = _problem()
        lo, hi = shard_range(len(g), rank, world)
        shard = g[lo:hi]

        def local_topk(queries, k):
            s = reid_ref.cosine_scores(queries.numpy(), shard)
            kk = min(k, s.shape[1])
            top, idx = reid_ref.topk_rule(s, kk, row_base=lo)
            ps = np.full((s.shape[0], k), -np.inf, np.float32)
            pi = np.full((s.shape[0], k), 0x7FFFFFFF, np.int32)
            ps[:, :kk], pi[:, :kk] = top, idx
            return torch.from_numpy(ps), torch.from_numpy(pi)

        def merge(all_s, all_i):
            s, i = reid_ref.merge_rule(all_s.numpy(), all_i.numpy(), all_s.shape[2])
            return torch.from_numpy(s), torch.from_numpy(i)

        sr = ShardedReID(local_topk=local_topk, merge=merge)
        assert sr.world == world and sr.rank == rank
        if ragged:
            counts = [len(range(r, len(q), world)) for r in range(world)]     # clips_of_rank ownership
            mine = q[clips_of_rank(len(q), rank, world)]
            order = sum((clips_of_rank(len(q), r, world) for r in range(world)), [])
            s, i = sr.search(torch.from_numpy(mine), k=5, counts=counts)
        else:
            per = len(q) // world
            mine = q[rank * per:(rank + 1) * per]
            order = list(range(per * world))
            s, i = sr.search(torch.from_numpy(mine), k=5)
        np.save(os.path.join(out_dir, f"s{rank}.npy"), s.numpy())
        np.save(os.path.join(out_dir, f"i{rank}.npy"), i.numpy())
        np.save(os.path.join(out_dir, f"o{rank}.npy"), np.array(order))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ragged", [(2, False), (2, True), (3, True)])
def test_sharded_topk_equals_single_shard(tmp_path, world, ragged):
    mp.spawn(_worker, args=(world, _free_port(), ragged, str(tmp_path)), nprocs=world, join=True)
    q, g = _problem()
    ref_s, ref_i = reid_ref.topk_rule(reid_ref.cosine_scores(q, g), 5)
    outs = [(np.load(tmp_path / f"s{r}.npy"), np.load(tmp_path / f"i{r}.npy"), np.load(tmp_path / f"o{r}.npy")) for r in range(world)]
    for s, i, order in outs:
        assert (i == outs[0][1]).all() and (s == outs[0][0]).all(), "every rank must hold the same result"
        assert (i == ref_i[order]).all(), "sharded top-k indices differ from the single-shard rule"
        np.testing.assert_array_equal(s, ref_s[order])
    assert outs[0][1][list(outs[0][2]).index(0)].tolist()[:3] == [3, 77, 400]
