"""SURVEY 8f-3 / BASELINE config 5: batched scoring of (candidate video, user history) pairs, forward-only.

Replaces the loop of inference/save_logits_for_all_leave_SegMM.py:97-148: the reference moves every feature tensor
to the GPU per batch, L1-normalises, calls the model with mode="inference" and then converts every row of logits to a
Python list keyed by f"{uid}-{pid}-{time_ms}" (`logit.cpu().detach().tolist()` per row: one sync per interaction).
Here the embedding table is resident in HBM, a batch is int32 row ids (gather + pad + mask + L1-normalise fused in
mmi_gather_l1norm_fwd), the logits of ALL batches are packed into one [n, 40] device tensor and come back to the host
with a single copy; `to_dict` rebuilds the reference's mapping when a caller wants the JSON the script wrote.
"""
from __future__ import annotations

import torch

from . import _lib
from .train import DeviceGather

PHOTO_MAX = 40


class InferenceScorer:
    def __init__(self, model, table: torch.Tensor, max_batch: int = 1024):
        if not table.is_cuda:
            raise _lib.MMIError("InferenceScorer needs the embedding table in device memory; there is no CPU fallback")
        self.model = model
        self.engine = model.engine()
        self.gather = DeviceGather(table, self.engine.act_dtype)
        self.max_batch = int(max_batch)

    @torch.no_grad()
    def score(self, usr_idx: torch.Tensor, vid_idx: torch.Tensor, usr_id=None, vid_id=None, out: torch.Tensor | None = None):
        """logits [n, 40] fp32 (incl. the learnable position bias when the model has one) for n interactions given as
        device int32 row ids usr_idx [n, Lt], vid_idx [n, 40] (-1 = pad).  Runs in chunks of `max_batch`; nothing is
        synchronised."""
        n = usr_idx.shape[0]
        dev = usr_idx.device
        if out is None:
            out = torch.empty(n, PHOTO_MAX, dtype=torch.float32, device=dev)
        for a in range(0, n, self.max_batch):
            b = min(n, a + self.max_batch)
            usr, um = self.gather(usr_idx[a:b], "usr")
            vid, vm = self.gather(vid_idx[a:b], "vid")
            uid = usr_id[a:b] if usr_id is not None else torch.zeros(b - a, dtype=torch.int64, device=dev)
            pid = vid_id[a:b] if vid_id is not None else torch.zeros(b - a, dtype=torch.int64, device=dev)
            res = self.model(usr_image=usr, usr_id=uid, usr_mask=um, vid_image=vid, vid_id=pid, vid_mask=vm, gt=None, mode="inference")
            out[a:b].copy_(res["logits"])
        return out

    def score_host(self, usr_idx_h: torch.Tensor, vid_idx_h: torch.Tensor, usr_id_h=None, vid_id_h=None) -> torch.Tensor:
        """End-to-end variant: pinned host row ids in, packed host logits [n, 40] out (one D2H copy + sync)."""
        dev = self.gather.table.device
        u = usr_idx_h.to(dev, non_blocking=True)
        v = vid_idx_h.to(dev, non_blocking=True)
        uid = usr_id_h.to(dev, non_blocking=True) if usr_id_h is not None else None
        pid = vid_id_h.to(dev, non_blocking=True) if vid_id_h is not None else None
        return self.score(u, v, uid, pid).cpu()

    @staticmethod
    def to_dict(logits_host: torch.Tensor, user_id, photo_id, time_ms) -> dict:
        """{"uid-pid-time_ms": [40 floats]} exactly as inference/save_logits_for_all_leave_SegMM.py:129-132 builds it (later
        duplicates of a key overwrite earlier ones, like the reference's dict assignment)."""
        rows = logits_host.tolist()
        return {f"{int(u)}-{int(p)}-{int(t)}": r for u, p, t, r in zip(user_id, photo_id, time_ms, rows)}
