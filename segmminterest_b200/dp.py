"""Batch-sharded data parallelism: one process per GPU, model and
embedding table replicated, each rank gathers and runs its own interactions, and the only
exchange is a SUM all-reduce of the flat gradient buffer, cut into buckets that are handed to
NCCL (NVLink 5 / NVSwitch) as soon as the hand-written backward has finished them, so the
transfer overlaps the rest of backward.  The reference has no distributed path at all
(DistributedDataParallel is imported and never used, main...SegMM.py:13)."""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradBuckets:
    def __init__(self, flat_grad: torch.Tensor, group=None, bucket_bytes: int = 25 << 20, skip=()):
        """skip: [(offset, numel), ...] ranges of the flat buffer that are NOT all-reduced (embedding tables whose
        gradients travel row-sparse, see exchange_rows)."""
        self.flat = flat_grad
        self.group = group
        self.skip = sorted((int(o), int(o) + int(n)) for o, n in skip)
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._hi = flat_grad.numel()      # everything at or above _hi has been handed to NCCL
        self._ready_lo = flat_grad.numel()
        self._works = []
        self.n_collectives = 0

    def begin(self):
        self._hi = self.flat.numel()
        self._ready_lo = self._hi
        self._works = []

    def ready(self, lo: int):
        """Backward finished every gradient at flat offsets >= lo."""
        if self.world == 1:
            return
        self._ready_lo = min(self._ready_lo, lo)
        if self._hi - self._ready_lo >= self.bucket_elems:
            self._launch(self._ready_lo, self._hi)

    def _launch(self, lo, hi):
        if hi <= lo:
            return
        # all_reduce(async_op=True) orders the NCCL kernel after everything already queued on the
        # current (compute) stream and runs it on the process group's own stream.
        a = lo
        for s0, s1 in self.skip + [(hi, hi)]:            # the pieces of [lo, hi) outside the skipped ranges
            b = min(max(s0, a), hi)
            if b > a:
                w = dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self._works.append(w)
                self.n_collectives += 1
            a = max(a, min(s1, hi))
        self._hi = lo

    def finish(self):
        """Flushes the tail bucket and makes the compute stream wait for all reductions."""
        if self.world == 1:
            return
        self._launch(0, self._hi)
        for w in self._works:
            w.wait()
        self._works = []


def exchange_rows(ids: torch.Tensor, rows: torch.Tensor, group=None):
    """Row-sparse gradient exchange for an embedding table: every rank contributes its B (id, gradient row)
    pairs; returns the concatenation over ranks ([world * B] ids, [world * B, tw] rows).  Payload per rank B x (8 + 4 tw)
    bytes instead of a dense n_rows x tw x 4 all-reduce (1 MB against 361 MB for the reference's video table at B = 1024)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ids, rows
    all_ids = torch.empty(world * ids.numel(), dtype=ids.dtype, device=ids.device)
    all_rows = torch.empty(world * rows.shape[0], rows.shape[1], dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(all_ids, ids.contiguous(), group=group)
    dist.all_gather_into_tensor(all_rows, rows.contiguous(), group=group)
    return all_ids, all_rows


def shard_rows(n_rows: int, rank: int, world: int):
    """Rank r owns rows [r*n/world, (r+1)*n/world) of a global batch."""
    per = n_rows // world
    return rank * per, (rank + 1) * per if rank < world - 1 else n_rows
