#!/usr/bin/env python
"""Debug: clock64 trace of one CTA of the all-keys attention backward (library built with -DMMI_ATTN_TRACE:
tools/build_variant.sh trace attention_tc.cu -DMMI_ATTN_TRACE; run with MMI_LIB_PATH=.../libmmi_trace.so).
Usage: python tools/attn_trace_all.py [Lq] [dropout 0|1]   (keys: 40 candidate + 500 history, the c2 shapes)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from segmminterest_b200 import _lib, ops  # noqa: E402
from segmminterest_b200.dropout import DropSite, quantise  # noqa: E402

Lq = int(sys.argv[1]) if len(sys.argv) > 1 else 500
drop = len(sys.argv) > 2 and sys.argv[2] == "1"
La, Lb = 40, 500
B, H, dh = 512, 16, 32
d = H * dh
dev = torch.device("cuda:0")
mq = torch.ones(B, Lq, dtype=torch.uint8, device=dev)
mka = (torch.arange(La, device=dev)[None] < 10).expand(B, La).contiguous().view(torch.uint8)
mkb = torch.ones(B, Lb, dtype=torch.uint8, device=dev)
rnd = lambda n: torch.randn(n, d, device=dev).mul_(0.5).bfloat16()  # noqa: E731
qa, qb, ka, va, kb, vb = rnd(B * Lq), rnd(B * Lq), rnd(B * La), rnd(B * La), rnd(B * Lb), rnd(B * Lb)
out = torch.empty(B * Lq, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, Lq, device=dev)
delta = torch.empty(B, H, Lq, device=dev)
dO = rnd(B * Lq)
g = [torch.empty_like(x) for x in (qa, ka, va, qb, kb, vb)]
thr8, scale = quantise(0.1)
site = DropSite(0x1234ABCD, thr8, scale) if drop else None
blocks = [dict(q=(qa.data_ptr(), d), k=(ka.data_ptr(), d), v=(va.data_ptr(), d), mask_k=mka, Lk=La),
          dict(q=(qb.data_ptr(), d), k=(kb.data_ptr(), d), v=(vb.data_ptr(), d), mask_k=mkb, Lk=Lb)]
side = ops.AttnSide(ops.BF16, ops.IMPL_TC, B, H, dh, Lq, mq, out, d, lse, blocks, drop=site)
side.fwd()
side.set_bwd(dO, d, delta, [dict(dq=(g[0].data_ptr(), d), dk=(g[1].data_ptr(), d), dv=(g[2].data_ptr(), d)),
                            dict(dq=(g[3].data_ptr(), d), dk=(g[4].data_ptr(), d), dv=(g[5].data_ptr(), d))])
assert side.bwd_all()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); side.bwd_all(); e1.record()
torch.cuda.synchronize()
print(f"all-keys bwd Lq={Lq} keys={La}+{Lb} dropout={drop}: {e0.elapsed_time(e1):.3f} ms per launch (B={B})")
buf = (C.c_longlong * 8192)()
lib = _lib.load()
lib.mmi_debug_trace.argtypes = [C.c_void_p, C.c_int]
print("rc", lib.mmi_debug_trace(buf, 8192))
t = list(buf)
NT = (La + 127) // 128 + (Lb + 127) // 128
T = (Lq + 63) // 64
base = t[4090]
print(f"cycles from CTA entry: tmem ready @{t[4091]-base}  CTA done @{t[4094]-base}  after final sync @{t[4092]-base}")
print("items of this CTA (cycles from entry): start@  key loop done@  accumulators complete@  epilogue stored@   (item length)")
for n in range(20):
    s = t[4000 + 4 * n: 4004 + 4 * n]
    if s[0] == 0 and n > 0:
        break
    nxt = t[4004 + 4 * n] if n < 19 and t[4004 + 4 * n] else s[3]
    print(f"  item {n:2d}: start@{s[0]-base:8d}  loop done +{s[1]-s[0]:6d}  acc done +{s[2]-s[1]:6d}  epilogue +{s[3]-s[2]:6d}   total {s[3]-s[0]:6d}")
print("softmax warp 2 per tile (cycles): start@  wait a_ready | tmem ld | compute | wait p_free | write+fence+arrive   || MMA warp: S issued@  back issued@")
for u in range(min(T * NT * 3, 45)):
    s = t[u * 8: u * 8 + 8]
    print(f"  t={u:2d} (i={(u // NT) % T}, j={u % NT}) start@{s[0]-base:7d}  a_ready {s[1]-s[0]:5d}  ld {s[2]-s[1]:5d}  compute {s[3]-s[2]:5d}  p_free {s[4]-s[3]:5d}  "
          f"write {s[5]-s[4]:5d}   || S@{s[6]-base:7d} back@{s[7]-base:7d}")
