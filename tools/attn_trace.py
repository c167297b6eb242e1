#!/usr/bin/env python
"""Debug: clock64 trace of one forward-attention CTA (library built with MMI_NVCC_EXTRA=-DMMI_ATTN_TRACE)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from segmminterest_b200 import _lib, ops  # noqa: E402

B, H, dh, Lv, Lt = 512, 16, 32, 40, 500
d = H * dh
dev = torch.device("cuda:0")
mv = (torch.arange(Lv, device=dev)[None] < 10).expand(B, Lv).contiguous().view(torch.uint8)
mt = torch.ones(B, Lt, dtype=torch.uint8, device=dev)
qv = torch.randn(B * Lv, 6 * d, device=dev).mul_(0.5).bfloat16()
qu = torch.randn(B * Lt, 6 * d, device=dev).mul_(0.5).bfloat16()
out = torch.empty(B * Lt, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, Lt, device=dev)
col = lambda t, j: (t.data_ptr() + j * d * 2, 6 * d)
blocks = [dict(q=col(qu, 2), k=col(qv, 4), v=col(qv, 5), mask_k=mv, Lk=Lv), dict(q=col(qu, 3), k=col(qu, 4), v=col(qu, 5), mask_k=mt, Lk=Lt)]
side = ops.AttnSide(ops.BF16, ops.IMPL_TC, B, H, dh, Lt, mt, out, d, lse, blocks)
side.fwd(); side.fwd()
torch.cuda.synchronize()
buf = (C.c_longlong * 8192)()
lib = _lib.load()
lib.mmi_debug_trace.argtypes = [C.c_void_p, C.c_int]
print("rc", lib.mmi_debug_trace(buf, 8192))
t = list(buf)
T = 17
base = t[4090]
print("cta: tmem_setup done @", t[4091] - base, " last tile done @", t[4093] - base, " end @", t[4092] - base)
print("softmax warp 4 per tile: start@, wait a_ready, ld+arrive, max+exp, write+arrive")
for j in range(T):
    s = t[j * 8: j * 8 + 5]
    print(f"  j={j:2d} start@{s[0]-base:6d}  wait {s[1]-s[0]:5d}  ld {s[2]-s[1]:5d}  compute {s[3]-s[2]:5d}  write {s[4]-s[3]:5d}")
