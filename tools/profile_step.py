#!/usr/bin/env python
"""One warm-up + N training steps of the bench workload, for ncu captures and for a
per-launch CUDA-event table (python tools/profile_step.py --table)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import model_args  # noqa: E402
from segmminterest_b200 import ops, synth  # noqa: E402
from segmminterest_b200.model import build_model  # noqa: E402
from segmminterest_b200.profiler import TIMER  # noqa: E402
from segmminterest_b200.train import TrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--table", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1, help="0 = eval()-mode arithmetic")
    a = ap.parse_args()
    wl = synth.WORKLOADS[a.workload]
    B = a.batch or wl.batch
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    model = build_model(model_args(a.precision), din=wl.din, max_usr_len=wl.lt).to(dev).train(a.dropout > 0)
    g = torch.Generator(device=dev).manual_seed(1234)
    table = torch.randn(wl.n_rows, wl.din, generator=g, device=dev)
    ts = TrainStep(model, table, global_batch=B, dropout=a.dropout)
    u, v, gt = synth.make_indices(B, wl.lt, wl.segs_per_video, wl.n_rows, seed=2025)
    u, v, gt = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev), torch.from_numpy(gt).to(dev)
    for _ in range(a.warmup):
        ts.step(u, v, gt)
    torch.cuda.synchronize()
    if a.table:
        TIMER.enabled = True
        TIMER.detail = True
    for _ in range(a.steps):
        ts.step(u, v, gt)
    torch.cuda.synchronize()
    if a.table:
        rows = {}
        for cat, s, e, work in TIMER.records:
            ms = s.elapsed_time(e)
            r = rows.setdefault(cat, [0.0, 0, 0.0])
            r[0] += ms; r[1] += 1; r[2] += work
        tot = sum(r[0] for r in rows.values())
        print(f"{'kernel':60s} {'ms':>9s} {'n':>4s} {'%':>6s} {'TF/s|GB/s':>10s}")
        for cat, (ms, n, work) in sorted(rows.items(), key=lambda kv: -kv[1][0]):
            rate = work / (ms * 1e-3) / 1e12 if work else 0.0
            print(f"{cat:60s} {ms / a.steps:9.3f} {n // a.steps:4d} {100 * ms / tot:6.1f} {rate:10.1f}")
        print(f"total {tot / a.steps:.3f} ms/step")


if __name__ == "__main__":
    main()
