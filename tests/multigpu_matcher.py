"""Row-sharded re-ID on real GPUs over NCCL, one process per GPU (run under torchrun; tests/test_gpu_pipeline.py spawns it when the
box has >= 2 GPUs, tools/gpu_multi.sh runs it on N):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_matcher.py

Every rank drives a sharded CowReIDMatcher through the reference's own transcripts (tests/golden/reid_scenario.json, reid_tight.npz)
and checks a sharded 30 000-row search (k = 5 and k = 20, planted ties) against a single-device scan of the same gallery.
Prints one line `MULTIGPU_MATCHER_OK world=N` from rank 0 on success; any assertion fails the launch."""
import asyncio
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from conftest import GOLDEN, engine_cached, replay_reid_tight
    from oracle import fake_services
    from oracle.make_golden import reid_queries
    import vision_sam3_yolo_lameless_b200.reid as reid_mod
    from vision_sam3_yolo_lameless_b200.gallery import GpuGallery, ShardedGpuGallery

    eng = engine_cached("b", layers=1, max_frames=16)[0]
    want = json.load(open(GOLDEN / "reid_scenario.json"))
    m = reid_mod.CowReIDMatcher(qdrant_url="fake://", engine=eng, qdrant_client=fake_services.FakeQdrant(), sharded=True)
    asyncio.run(m.connect())
    assert isinstance(m.client, ShardedGpuGallery) and m.client.world == world
    for k, ((name, q), step) in enumerate(zip(reid_queries(), want["steps"])):
        got = m.match_or_create(np.asarray(q), video_id=f"video-{name}", track_id=k)
        assert (got.cow_id, got.confidence, got.is_new_identity) == (step["cow_id"], step["confidence"], step["is_new"]), (rank, name)
        assert abs(got.similarity - step["similarity"]) < 4e-4, (rank, name, got.similarity)
    _, cands = m.match_embedding(np.asarray(reid_queries()[0][1]))
    assert [c.cow_id for c in cands] == [c["cow_id"] for c in want["final_candidates"]]

    class _Shim(reid_mod.CowReIDMatcher):
        def __init__(self, **kw):
            super().__init__(sharded=True, **kw)

    orig, reid_mod.CowReIDMatcher = reid_mod.CowReIDMatcher, _Shim
    try:
        worst = replay_reid_tight(eng, GOLDEN, sim_tol=4e-4)
    finally:
        reid_mod.CowReIDMatcher = orig

    # a larger gallery: sharded search == single-device search, bit for bit (scores and global rows), ties included
    rng = np.random.default_rng(17)
    n = 30011
    vecs = rng.standard_normal((n, 768)).astype(np.float32)
    vecs[29000] = vecs[17]
    vecs[12345] = vecs[17]
    ids = [f"p{i}" for i in range(n)]
    whole, shard = GpuGallery(eng, 768), ShardedGpuGallery(eng, 768)
    whole.load(ids, vecs)
    shard.load(ids, vecs)
    queries = np.concatenate([vecs[17:18] + 0.01 * rng.standard_normal((1, 768)).astype(np.float32), rng.standard_normal((6, 768)).astype(np.float32)])
    for k in (5, 20):
        a, b = whole.search_batch(queries, k), shard.search_batch(queries, k)
        assert [[p.id for p in r] for r in a] == [[p.id for p in r] for r in b], (rank, k)
        assert [[p.score for p in r] for r in a] == [[p.score for p in r] for r in b], (rank, k)
        assert [p.id for p in b[0]][:3] == ["p17", "p12345", "p29000"]
    one = shard.search(queries[0], 5)                       # the serving form (one query) on every shard
    assert [p.id for p in one] == [p.id for p in whole.search(queries[0], 5)]
    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU_MATCHER_OK world={world} worst_similarity_diff={worst:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
