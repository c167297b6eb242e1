#!/usr/bin/env python
"""Relative error (vs fp64) of the fp32 GEMM paths: FFMA kernel vs the split-bf16 tensor-core path, as a function of K.

    python tools/split_gemm_error.py > profiles/rNN_split_gemm_error.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from segmminterest_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    return float((a.double() - b).norm() / b.norm())


print(f"{'product':34s} {'FFMA':>10s} {'split-bf16 tcgen05':>20s}   signed mean of (C - ref)/|ref| for the split path")
for layout, (M, N, K) in [("NT", (4096, 512, 512)), ("NT", (4096, 512, 640)), ("NT", (4096, 512, 3072)), ("NT", (4096, 3072, 512)),
                          ("NN", (4096, 512, 3072)), ("TN", (512, 512, 4096)), ("TN", (3072, 512, 65536)), ("TN", (512, 512, 512000))]:
    torch.manual_seed(0)
    A = torch.randn(M, K, device=dev)
    Bm = torch.randn(N, K, device=dev)
    ref = (A.double() @ Bm.double().T)
    out = {}
    for mode in (False, True):
        ops.FP32_TC["on"] = mode
        if layout == "NT":
            C = torch.empty(M, N, device=dev)
            ops.gemm(ops.GEMM_NT, ops.IMPL_SIMT, A, K, Bm, K, C, N, M, N, K)
        elif layout == "NN":
            C = torch.empty(M, N, device=dev)
            ops.gemm(ops.GEMM_NN, ops.IMPL_SIMT, A, K, Bm.T.contiguous(), N, C, N, M, N, K)
        else:
            C = torch.zeros(M, N, device=dev)
            ops.gemm(ops.GEMM_TN, ops.IMPL_SIMT, A.T.contiguous(), M, Bm.T.contiguous(), N, C, N, M, N, K, accumulate=True,
                     split_k=max(1, min(32, K // 4096)))
        out[mode] = C
    ops.FP32_TC["on"] = False
    bias = float(((out[True].double() - ref) / ref.abs().clamp_min(1e-3)).mean())
    print(f"{layout} M={M} N={N} K={K:<8d}      {rel(out[False], ref):10.2e} {rel(out[True], ref):20.2e}   {bias:+.2e}")
