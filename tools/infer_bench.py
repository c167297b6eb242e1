#!/usr/bin/env python
"""BASELINE config 5: forward-only batched scoring of (candidate video, user history) pairs at the c2 shapes
(history 50 videos x 10 segments, Din 640), through InferenceScorer.  Prints one JSON line: scored pairs/s with the row
ids resident in HBM and end to end from pinned host row ids (one D2H copy of the packed logits).

    python tools/infer_bench.py [--pairs 65536] [--batch 1024] [--precision bf16]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import model_args  # noqa: E402
from segmminterest_b200 import InferenceScorer, ops, synth  # noqa: E402
from segmminterest_b200.model import build_model  # noqa: E402
from segmminterest_b200.profiler import TIMER  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=65536)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    wl = synth.WORKLOADS["c2"]
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    model = build_model(model_args(a.precision), din=wl.din, max_usr_len=wl.lt).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1234)
    table = torch.randn(wl.n_rows, wl.din, generator=g, device=dev)
    scorer = InferenceScorer(model, table, max_batch=a.batch)
    n = a.pairs
    u, v, _ = synth.make_indices(n, wl.lt, wl.segs_per_video, wl.n_rows, seed=2025)
    hu, hv = torch.from_numpy(u).pin_memory(), torch.from_numpy(v).pin_memory()
    du, dv = hu.to(dev), hv.to(dev)
    out = torch.empty(n, 40, dtype=torch.float32, device=dev)
    scorer.score(du[: 3 * a.batch], dv[: 3 * a.batch])            # warm-up (3 batches)
    torch.cuda.synchronize()
    n0 = ops.LaunchCounter.n
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    scorer.score(du, dv, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ops.LaunchCounter.n - n0
    e0.record()
    host = scorer.score_host(hu, hv)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    TIMER.enabled = True
    TIMER.reset()
    scorer.score(du[: 4 * a.batch], dv[: 4 * a.batch])
    summ = TIMER.summary()
    TIMER.enabled = False
    tot = sum(x["ms"] for x in summ.values())
    line = {"metric": "inference_interactions_per_s", "value": n / (ms * 1e-3), "unit": "interactions/s", "n_gpus": 1,
            "pairs": n, "batch": a.batch, "ms_total": ms, "dtype": a.precision, "data": "synthetic",
            "config": {"workload": "c5_inference_scoring_c2_shapes", "hist_len": wl.lt, "cand_pad": 40, "din": wl.din},
            "e2e": {"value": n / (ms_e2e * 1e-3), "unit": "interactions/s", "h2d_bytes": int(hu.numel() * 4 + hv.numel() * 4),
                    "d2h_bytes": int(host.numel() * 4)},
            "gpu_launches": launches, "finite": bool(torch.isfinite(host).all()),
            "kernel_share": {k: round(x["ms"] / tot, 3) for k, x in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
