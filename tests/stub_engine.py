"""Test double for ClipEmbedEngine on boxes without a GPU: same method surface, arithmetic by the CPU oracle.
It exists so the HOST logic of the product (extractor / gallery / matcher / handler / sharding) can be tested
under `-m "not gpu"`; it lives in tests/ and is never importable from the package."""
import numpy as np
import torch

from oracle import preprocess_ref, reid_ref, vit_ref
from vision_sam3_yolo_lameless_b200.engine import VitConfig


class StubEngine:
    def __init__(self, model, max_frames=64, resize=(224, 224)):
        self.cfg = VitConfig.from_hf(model.config)
        self.sd = model.state_dict()
        self.device = torch.device("cpu")
        self.max_frames, self.resize = max_frames, resize
        self.calls = []

    def embed_host_frames(self, frames, bgr=True, out=None):
        if isinstance(frames, (list, tuple)):      # one array per clip, embedded as if concatenated (engine.embed_host_frames)
            frames = np.concatenate([np.asarray(f) for f in frames], axis=0)
        frames = np.asarray(frames)
        self.calls.append(("embed", frames.shape))
        pv = preprocess_ref.preprocess(frames, bgr=bgr, size=self.resize)
        return vit_ref.frame_embeddings(self.sd, torch.from_numpy(pv), heads=self.cfg.heads, layers=self.cfg.layers)

    embed_frames = embed_host_frames

    def embed_rois(self, frames, rois, bgr=True):
        frames = frames.numpy() if isinstance(frames, torch.Tensor) else np.asarray(frames)
        self.calls.append(("rois", len(rois)))
        out = [self.embed_host_frames(frames[f:f + 1, y0:y1, x0:x1], bgr=bgr)[0] for f, x0, y0, x1, y1 in np.asarray(rois)]
        return torch.stack(out)

    def pool_clips(self, frame_emb, clip_offsets):
        offs = np.asarray(clip_offsets)
        mean = reid_ref.clip_mean(frame_emb.numpy(), offs)
        return torch.from_numpy(mean.astype(np.float32)), torch.from_numpy(reid_ref.l2_normalise(mean).astype(np.float32))

    def gallery_topk(self, queries, gallery, k=5, row_base=0, dump_scores=False):
        scores = reid_ref.cosine_scores(queries.numpy(), gallery.float().numpy())
        kk = min(k, scores.shape[1])
        top, idx = reid_ref.topk_rule(scores, kk, row_base)
        s = np.full((scores.shape[0], k), -np.inf, dtype=np.float32)
        i = np.full((scores.shape[0], k), 0x7FFFFFFF, dtype=np.int32)
        s[:, :kk], i[:, :kk] = top, idx
        out = (torch.from_numpy(s), torch.from_numpy(i))
        return out + (torch.from_numpy(scores),) if dump_scores else out

    def merge_topk(self, scores, idx):
        s, i = reid_ref.merge_rule(scores.numpy(), idx.numpy(), scores.shape[2])
        return torch.from_numpy(s), torch.from_numpy(i)

    def gallery_update_row(self, gallery, row, unit_query, momentum, master=None):
        old = (master if master is not None else gallery)[row].float().numpy().astype(np.float64)
        uq = unit_query.reshape(-1).numpy().astype(np.float64)
        v = momentum * old + (1.0 - momentum) * uq if momentum != 0.0 else uq
        v = v / (np.linalg.norm(v) + 1e-8)
        gallery[row] = torch.from_numpy(v).to(gallery.dtype)
        if master is not None:
            master[row] = torch.from_numpy(v).to(master.dtype)
