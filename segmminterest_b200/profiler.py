"""CUDA-event timing of kernel categories on the launching stream (bench.py roofline leg).
Disabled by default: zero overhead in the product path."""
from __future__ import annotations

import contextlib

import torch


class KernelTimer:
    def __init__(self):
        self.enabled = False
        self.detail = False   # per-shape categories (tools/profile_step.py --table)
        self.records = []   # (category, start_event, end_event, work)  work = flops or bytes

    @contextlib.contextmanager
    def region(self, category: str, work: float = 0.0):
        if not self.enabled:
            yield
            return
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        try:
            yield
        finally:
            e.record()
            self.records.append((category, s, e, work))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for cat, s, e, work in self.records:
            ms = s.elapsed_time(e)
            d = out.setdefault(cat, {"ms": 0.0, "launches": 0, "work": 0.0})
            d["ms"] += ms
            d["launches"] += 1
            d["work"] += work
        return out

    def reset(self):
        self.records = []


TIMER = KernelTimer()
