"""The fused training step behind the reference driver's loop body
(main_for_seq_leave_earlystop_SegMM.py:270-300): device-side gather + pad + mask +
L1-normalise -> forward -> focal loss -> hand-written backward -> bucketed gradient
all-reduce -> (optional global-norm clip) + AdamW, all on one stream with no host synchronisation."""
from __future__ import annotations

import torch

from . import _lib, ops
from .dp import GradBuckets, exchange_rows

PHOTO_MAX = 40


class DeviceGather:
    """a-1..a-3: the embedding table lives in HBM; a batch is described by int32 row ids
    (-1 = pad) instead of the reference's per-segment string-keyed dict lookups
    (utils/dataloader_SegMM.py:301-350)."""

    def __init__(self, table: torch.Tensor, out_dtype=torch.float32):
        if not table.is_cuda:
            raise _lib.MMIError("DeviceGather needs the table in device memory; there is no CPU fallback")
        self.table = table.contiguous()
        self.out_dtype = out_dtype
        self._buf = {}
        self.valid_hint = {}     # profiling only: tag -> number of valid ids per call (host-known; see ops.gather_l1norm)

    def __call__(self, idx: torch.Tensor, tag: str, normalise: bool = True):
        B, L = idx.shape
        key = (tag, B, L)
        if key not in self._buf:
            self._buf[key] = (torch.empty(B, L, self.table.shape[1], device=idx.device, dtype=self.out_dtype),
                              torch.empty(B, L, device=idx.device, dtype=torch.uint8))
        out, mask = self._buf[key]
        ops.gather_l1norm(self.table, idx, out, mask, normalise, n_valid=self.valid_hint.get(tag))
        return out, mask.view(torch.bool)


class TrainStep:
    def __init__(self, model, table: torch.Tensor, lr=1e-3, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 max_norm=None, process_group=None, global_batch=None, bucket_bytes=25 << 20, dropout=None, sparse_tables=None):
        """max_norm: None / 0 (default) applies no gradient clipping -- what the reference driver effectively does: its
        `clip_grad_norm_(param_dict, 10.0)` (main...SegMM.py:298) receives the generator AdamW already consumed (:224-225),
        so nothing is clipped; a float opts in to real global-norm clipping (the norm is reported either way).
        dropout: None follows the module (model.dropout_p(): the backbone's p under train(), 0 under eval()); a float
        overrides it.  Every rank and every forward call draws its own masks (seed mixes torch.initial_seed() and the rank)."""
        self.model = model
        self.engine = model.engine()
        eng = self.engine
        self.dropout = dropout
        self.gather = DeviceGather(table, eng.act_dtype)
        self.lr, self.wd, self.betas, self.eps, self.max_norm = lr, weight_decay, betas, eps, max_norm
        self.exp_avg = torch.zeros_like(eng.flat)
        self.exp_avg_sq = torch.zeros_like(eng.flat)
        self.norm = torch.zeros(2, device=eng.device)
        self.ws = torch.empty(int(_lib.load().mmi_clip_adamw_workspace(eng.n_flat)), device=eng.device)
        self.step_no = 0
        # sparse_tables: embedding-table gradients of ID towers travel as (ids, rows) between the ranks instead of through a
        # dense all-reduce (None: on whenever there is more than one rank and the model has ID tables)
        import torch.distributed as dist
        world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        tables = eng.table_groups()
        self.sparse_tables = bool(tables) and (world > 1 if sparse_tables is None else bool(sparse_tables))
        eng.sparse_tables = self.sparse_tables
        self.buckets = GradBuckets(eng.flat_grad, process_group, bucket_bytes, skip=tables if self.sparse_tables else ())
        eng.drop_seed = (int(torch.initial_seed()) + 0x9E3779B97F4A7C15 * self.buckets.rank) & 0xFFFFFFFFFFFFFFFF
        self.global_batch = global_batch
        self.loss_cfg = model.loss_cfg()

    def step(self, usr_idx: torch.Tensor, vid_idx: torch.Tensor, gt: torch.Tensor, usr_id=None, vid_id=None, micro_batch: int = 0):
        """One training step on device-resident int32 indices [B,Lt], [B,40] and int64 labels
        [B,40] (rewritten in place like the reference).  Returns the device scalars
        [focal, mse, mse2, loss, ...]; nothing is synchronised.

        micro_batch > 0 runs the batch in slices of that many interactions with gradient accumulation (the saved
        activations of one slice are all that lives in HBM: config 3 / 4 histories do not fit otherwise); the gradient
        all-reduce overlaps the backward of the LAST slice.  Exact for losses that are sums over interactions / B_global
        (focal, hazard, interestCE/KL); losses whose mean runs over a data-dependent count (interestBPR, surviveCE) or
        over pairs of rows (huber, the mse diagnostics) are evaluated per slice and averaged."""
        eng = self.engine
        B = usr_idx.shape[0]
        gb = self.global_batch or B * self.buckets.world
        eng.ensure_bound()
        if self.buckets.flat.data_ptr() != eng.flat_grad.data_ptr() or self.exp_avg.numel() != eng.flat.numel():
            # the engine re-bound its flat buffers (.to(), load_state_dict(assign=True), a replaced .data): reducing the old
            # gradient buffer would silently leave every rank on its local gradients.  The layout is unchanged, so the Adam
            # moments stay valid; only the bucket views move.
            if self.exp_avg.numel() != eng.flat.numel():
                raise _lib.MMIError("the engine's parameter layout changed under a live TrainStep; build a new TrainStep")
            self.buckets = GradBuckets(eng.flat_grad, self.buckets.group, self.buckets.bucket_elems * 4,
                                       skip=eng.table_groups() if self.sparse_tables else ())
        eng.bind_grads()           # Parameters' .grad alias the flat gradient buffer
        eng.flat_grad.zero_()
        eng.drop_p = self.model.dropout_p() if self.dropout is None else float(self.dropout)
        mb = int(micro_batch) if micro_batch and micro_batch < B else B
        n_slices = (B + mb - 1) // mb
        total = None
        for si, a in enumerate(range(0, B, mb)):
            b = min(B, a + mb)
            usr, um = self.gather(usr_idx[a:b], "usr")
            vid, vm = self.gather(vid_idx[a:b], "vid")
            logits = eng.forward(usr, um, vid, vm, usr_id=None if usr_id is None else usr_id[a:b],
                                 vid_id=None if vid_id is None else vid_id[a:b], refresh=si == 0)
            # focal is a sum / B_global; interestBPR is a mean over this slice's rows, averaged over slices and ranks (DDP semantics)
            scal, _ = eng.loss(logits, gt[a:b], self.model.exposure_prob, 1.0 / gb, self.loss_cfg,
                               bpr_scale=1.0 / (self.buckets.world * n_slices))
            last = si == n_slices - 1
            if last:
                self.buckets.begin()
            eng.table_rows = []
            eng.backward(None, on_ready=self.buckets.ready if last else None)
            for key, tw_cols, ids_, rows_ in eng.table_rows:       # row-sparse exchange of this slice's table gradients
                all_ids, all_rows = exchange_rows(ids_, rows_, self.buckets.group)
                gtab = eng.g(key)
                ops.scatter_rows_add(all_ids, all_rows, tw_cols, gtab.numel() // tw_cols, gtab)
            if n_slices > 1:
                total = scal.clone() if total is None else total.add_(scal)
        if n_slices > 1:
            total[1:3].div_(n_slices)          # mse / mse2 diagnostics: mean of the per-slice values
            scal = total
        self.buckets.finish()
        self.step_no += 1
        ops.clip_adamw(eng.flat, eng.flat_grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps,
                       self.wd, self.max_norm if self.max_norm else 0.0, self.step_no, self.norm, None, self.ws)
        return scal

    def check_indices(self, *idx):
        """Debug aid (one sync): raises IndexError when a row id is past the table -- the reference's feat_memmap[line_id]
        does; the gather kernel treats such an id as padding."""
        n = self.gather.table.shape[0]
        for t in idx:
            if t.numel() and int(t.max()) >= n:
                raise IndexError(f"row id {int(t.max())} is past the embedding table ({n} rows)")

    def step_host(self, usr_idx_h, vid_idx_h, gt_h, dev_bufs, micro_batch: int = 0, usr_id=None, vid_id=None):
        """End-to-end variant: pinned host index/label buffers -> H2D inside the call, loss read
        back to the host (one D2H + sync), as the reference loop's `loss.item()` does."""
        u, v, g = dev_bufs
        u.copy_(usr_idx_h, non_blocking=True)
        v.copy_(vid_idx_h, non_blocking=True)
        g.copy_(gt_h, non_blocking=True)
        scal = self.step(u, v, g, usr_id=usr_id, vid_id=vid_id, micro_batch=micro_batch)
        return float(scal[3].item())
