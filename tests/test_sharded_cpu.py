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


def _matcher_worker(rank, world, port, out_dir):
    """Every rank drives a row-sharded CowReIDMatcher with the same calls (SPMD); arithmetic by the oracle-backed StubEngine."""
    import asyncio
    import json
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import GOLDEN, replay_reid_tight
        from oracle import common, fake_services
        from oracle.make_golden import reid_queries
        from stub_engine import StubEngine
        from vision_sam3_yolo_lameless_b200.gallery import ShardedGpuGallery
        from vision_sam3_yolo_lameless_b200.reid import CowReIDMatcher

        stub = StubEngine(common.hf_model(layers=1))
        want = json.load(open(GOLDEN / "reid_scenario.json"))
        qd = fake_services.FakeQdrant()
        m = CowReIDMatcher(qdrant_url="fake://", engine=stub, qdrant_client=qd, sharded=True)
        asyncio.run(m.connect())
        assert isinstance(m.client, ShardedGpuGallery) and m.client.world == world
        log = []
        for k, ((name, q), step) in enumerate(zip(reid_queries(), want["steps"])):
            got = m.match_or_create(np.asarray(q), video_id=f"video-{name}", track_id=k)
            assert (got.cow_id, got.confidence, got.is_new_identity) == (step["cow_id"], step["confidence"], step["is_new"]), name
            assert abs(got.similarity - step["similarity"]) < 4e-4, name
            log.append([got.cow_id, got.similarity])
        _, cands = m.match_embedding(np.asarray(reid_queries()[0][1]))
        assert [c.cow_id for c in cands] == [c["cow_id"] for c in want["final_candidates"]]
        # rows really are spread: global row r on rank r % world
        assert m.client._local_rows() == len(range(rank, len(m.client), world))
        # durable store: every rank that holds a client wrote every identity, at full precision (owner's master row, broadcast)
        assert len(qd.collections["cow_identities"]["ids"]) == 4
        # a second sharded matcher mirrors the store; more than 8 hits come back complete and ordered
        m2 = CowReIDMatcher(qdrant_url="fake://", engine=stub, qdrant_client=qd, sharded=True)
        asyncio.run(m2.connect())
        assert m2.identity_counter == 4 and m2.match_embedding(np.asarray(reid_queries()[0][1]))[0].cow_id == "COW-0001"
        # the near-threshold transcript with its durable-store check, sharded
        class _Shim(CowReIDMatcher):
            def __init__(self, **kw):
                super().__init__(sharded=True, **kw)
        import vision_sam3_yolo_lameless_b200.reid as reid_mod
        orig = reid_mod.CowReIDMatcher
        reid_mod.CowReIDMatcher = _Shim
        try:
            worst = replay_reid_tight(stub, GOLDEN, sim_tol=4e-4)
        finally:
            reid_mod.CowReIDMatcher = orig
        json.dump({"log": log, "worst": worst}, open(os.path.join(out_dir, f"m{rank}.json"), "w"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_matcher_reproduces_reference_transcript(tmp_path, world):
    """SURVEY 8(e) behind the reference-facing matcher: the gallery row-sharded over `world` ranks gives, on EVERY rank, the reference's
    own match_or_create transcript (tests/golden/reid_scenario.json, reid_tight.npz) -- reads are collectives, writes run on the owner."""
    import json
    mp.spawn(_matcher_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    logs = [json.load(open(tmp_path / f"m{r}.json")) for r in range(world)]
    assert all(l["log"] == logs[0]["log"] for l in logs), "every rank must report the same matches"
