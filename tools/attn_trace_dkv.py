#!/usr/bin/env python
"""Debug: clock64 trace of one dK/dV attention CTA (library built with MMI_NVCC_EXTRA=-DMMI_ATTN_TRACE).
Usage: python tools/attn_trace_dkv.py [Lq] [Lk] [dropout 0|1]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from segmminterest_b200 import _lib, ops  # noqa: E402
from segmminterest_b200.dropout import DropSite, quantise  # noqa: E402

Lq = int(sys.argv[1]) if len(sys.argv) > 1 else 500
Lk = int(sys.argv[2]) if len(sys.argv) > 2 else 500
drop = len(sys.argv) > 3 and sys.argv[3] == "1"
B, H, dh = 512, 16, 32
d = H * dh
dev = torch.device("cuda:0")
mq = torch.ones(B, Lq, dtype=torch.uint8, device=dev)
mk = torch.ones(B, Lk, dtype=torch.uint8, device=dev)
q = torch.randn(B * Lq, d, device=dev).mul_(0.5).bfloat16()
k = torch.randn(B * Lk, d, device=dev).mul_(0.5).bfloat16()
v = torch.randn(B * Lk, d, device=dev).mul_(0.5).bfloat16()
out = torch.empty(B * Lq, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, Lq, device=dev)
delta = torch.empty(B, H, Lq, device=dev)
dO = torch.randn(B * Lq, d, device=dev).mul_(0.5).bfloat16()
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
dbk, dbv = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
thr8, scale = quantise(0.1)
site = DropSite(0x1234ABCD, thr8, scale) if drop else None
side = ops.AttnSide(ops.BF16, ops.IMPL_TC, B, H, dh, Lq, mq, out, d, lse, [dict(q=(q.data_ptr(), d), k=(k.data_ptr(), d), v=(v.data_ptr(), d), mask_k=mk, Lk=Lk)], drop=site)
side.fwd()
side.set_bwd(dO, d, delta, [dict(dq=(dq.data_ptr(), d), dk=(dk.data_ptr(), d), dv=(dv.data_ptr(), d), dbk=dbk.data_ptr(), dbv=dbv.data_ptr())])
side.bwd_dq()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
side.bwd_dkv(0)
e0.record(); side.bwd_dkv(0); e1.record()
torch.cuda.synchronize()
print(f"dkv Lq={Lq} Lk={Lk} dropout={drop}: {e0.elapsed_time(e1):.3f} ms per launch (B={B})")
buf = (C.c_longlong * 8192)()
lib = _lib.load()
lib.mmi_debug_trace.argtypes = [C.c_void_p, C.c_int]
print("rc", lib.mmi_debug_trace(buf, 8192))
t = list(buf)
T = (Lq + 31) // 32
base = t[4090]
print(f"cycles from CTA entry: tmem_setup done @{t[4091]-base}  accumulators done @{t[4093]-base}  epilogue stored @{t[4094]-base}  teardown @{t[4092]-base}")
print("softmax warp 2 per tile: start@, wait kv_full, wait a_ready, ld+compute, p_free+write+arrive")
for j in range(T):
    s = t[j * 8: j * 8 + 5]
    print(f"  i={j:2d} start@{s[0]-base:6d}  kv_full {s[1]-s[0]:5d}  a_ready {s[2]-s[1]:5d}  compute {s[3]-s[2]:5d}  write {s[4]-s[3]:5d}")
