#!/usr/bin/env python
"""Times the tcgen05 attention kernels alone at a workload's shapes (CUDA events, L2 flushed by
size: the q/k/v buffers of one launch exceed 126 MB), for kernel iteration and ncu captures.

    python tools/attn_bench.py [--workload c2] [--batch 1024] [--iters 5] [--masked]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from segmminterest_b200 import ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--heads", type=int, default=16)
    ap.add_argument("--sides", default="usr,vid")
    ap.add_argument("--dropout", action="store_true", help="logits dropout 0.1 live (the train()-mode kernels)")
    ap.add_argument("--bias", action="store_true", help="with the fused bias-gradient sums")
    ap.add_argument("--fused", action="store_true", help="also time the per-key-block fused kernel")
    ap.add_argument("--no-two", action="store_true", help="skip the dq + dk/dv kernel pair")
    ap.add_argument("--valid-cand", type=int, default=-1, help="valid candidate segments (default: workload's S)")
    a = ap.parse_args()
    wl = synth.WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    B, H, dh, Lv, Lt = a.batch, a.heads, 32, 40, wl.lt
    d = H * dh
    nv = wl.segs_per_video if a.valid_cand < 0 else a.valid_cand
    torch.manual_seed(0)
    mv = (torch.arange(Lv, device=dev)[None] < nv).expand(B, Lv).contiguous().view(torch.uint8)
    mt = torch.ones(B, Lt, dtype=torch.uint8, device=dev)
    qkv = {"vid": torch.randn(B * Lv, 6 * d, device=dev).mul_(0.5).bfloat16(), "usr": torch.randn(B * Lt, 6 * d, device=dev).mul_(0.5).bfloat16()}
    dqkv = {k: torch.zeros_like(v) for k, v in qkv.items()}
    L = {"vid": Lv, "usr": Lt}
    mask = {"vid": mv, "usr": mt}
    esz = 2

    def col(t, s, j):
        return (t[s].data_ptr() + j * d * esz, 6 * d)

    print(f"{'kernel':28s} {'ms':>8s} {'TF/s':>8s} {'Gscore/s':>9s}")
    total = 0.0
    for s in a.sides.split(","):
        Lq = L[s]
        out = torch.empty(B * Lq, d, device=dev, dtype=torch.bfloat16)
        dout = torch.randn(B * Lq, d, device=dev).mul_(0.3).bfloat16()
        lse = torch.empty(B, H, Lq, device=dev)
        delta = torch.empty(B, H, Lq, device=dev)
        if s == "vid":
            idx = [(("vid", 0), ("vid", 1), ("vid", 2)), (("vid", 3), ("usr", 0), ("usr", 1))]
        else:
            idx = [(("usr", 2), ("vid", 4), ("vid", 5)), (("usr", 3), ("usr", 4), ("usr", 5))]
        blocks = [dict(q=col(qkv, *q), k=col(qkv, *k), v=col(qkv, *v), mask_k=mask[k[0]], Lk=L[k[0]]) for q, k, v in idx]
        grads = [dict(dq=col(dqkv, *q), dk=col(dqkv, *k), dv=col(dqkv, *v)) for q, k, v in idx]
        if a.bias:                       # fused bias-gradient column sums, as the engine asks for them
            dbs = [torch.zeros(d, device=dev) for _ in range(6)]
            for i_, g_ in enumerate(grads):
                g_.update(dbq=dbs[3 * i_].data_ptr(), dbk=dbs[3 * i_ + 1].data_ptr(), dbv=dbs[3 * i_ + 2].data_ptr())
        site = None
        if a.dropout:
            from segmminterest_b200.dropout import DropSite, quantise
            site = DropSite(0x1234ABCD, *quantise(0.1))
        side = ops.AttnSide(ops.BF16, ops.IMPL_TC, B, H, dh, Lq, mask[s], out, d, lse, blocks, drop=site)
        side.set_bwd(dout, d, delta, grads)
        acc = [torch.zeros(B * Lq, d, device=dev) for _ in range(2)]
        cnt = [torch.zeros(B * H, device=dev, dtype=torch.int32) for _ in range(2)]
        side.set_fused(acc, cnt)
        calls = [("fwd", side.fwd, 1.0, None)]
        if not a.no_two:
            calls += [("bwd_dq", side.bwd_dq, 1.5, None), ("bwd_dkv0", lambda: side.bwd_dkv(0), 2.0, 0), ("bwd_dkv1", lambda: side.bwd_dkv(1), 2.0, 1)]
        if a.fused:
            calls += [("bwd_fused0", lambda: side.bwd_fused(0), 2.5, 0), ("bwd_fused1", lambda: side.bwd_fused(1), 2.5, 1)]
        calls += [("bwd_all", side.bwd_all, 2.5, None)]
        for name, fn, mult, which in calls:
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            lk = sum(b["Lk"] for b in blocks) if which is None else blocks[which]["Lk"]
            scores = float(B) * H * Lq * lk
            total += ms
            print(f"{s + '.' + name:28s} {ms:8.3f} {mult * 4 * scores * dh / ms / 1e9:8.1f} {scores / ms / 1e6:9.1f}")
    print(f"sum {total:.3f} ms  (one full layer's attention, fwd + bwd)")


if __name__ == "__main__":
    main()
