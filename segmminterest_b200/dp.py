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
    def __init__(self, flat_grad: torch.Tensor, group=None, bucket_bytes: int = 25 << 20):
        self.flat = flat_grad
        self.group = group
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
        w = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._works.append(w)
        self.n_collectives += 1
        self._hi = lo

    def finish(self):
        """Flushes the tail bucket and makes the compute stream wait for all reductions."""
        if self.world == 1:
            return
        self._launch(0, self._hi)
        for w in self._works:
            w.wait()
        self._works = []


def shard_rows(n_rows: int, rank: int, world: int):
    """Rank r owns rows [r*n/world, (r+1)*n/world) of a global batch."""
    per = n_rows // world
    return rank * per, (rank + 1) * per if rank < world - 1 else n_rows
