"""GPU parity of each C-ABI kernel against the CPU oracle / plain torch fp32-fp64 maths.
Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "needs a CUDA device"
    from segmminterest_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


# ----------------------------------------------------------------------------- gather
@pytest.mark.parametrize("din", [640, 1024, 48, 768])
def test_gather_bit_exact_and_l1(dev, din):
    from oracle import gather_oracle
    from segmminterest_b200 import ops
    rng = np.random.default_rng(din)
    table = rng.standard_normal((997, din), dtype=np.float32)
    idx = rng.integers(-1, 997, size=(13, 57)).astype(np.int32)
    idx[0, :] = -1  # a fully padded history
    t = torch.from_numpy(table).to(dev)
    i = torch.from_numpy(idx).to(dev)
    out = torch.empty(13, 57, din, device=dev)
    mask = torch.empty(13, 57, dtype=torch.uint8, device=dev)
    ops.gather_l1norm(t, i, out, mask, normalise=False)
    ref, m = gather_oracle.gather_dense(table, idx)
    assert np.array_equal(out.cpu().numpy(), ref)          # bit-exact copy, zero pad rows
    assert np.array_equal(mask.cpu().numpy().astype(bool), m)
    ops.gather_l1norm(t, i, out, mask, normalise=True)
    refn = gather_oracle.l1_normalise(ref)
    got = out.cpu().numpy()
    assert np.allclose(got, refn, rtol=1e-6, atol=0)
    assert not got[~m].any()
    # bf16 output (tensor-core path input)
    outb = torch.empty(13, 57, din, device=dev, dtype=torch.bfloat16)
    ops.gather_l1norm(t, i, outb, mask, normalise=True)
    assert np.allclose(outb.float().cpu().numpy(), refn, rtol=2 ** -8, atol=0)


def test_gather_bf16_table(dev):
    from segmminterest_b200 import ops
    rng = np.random.default_rng(3)
    table = torch.from_numpy(rng.standard_normal((300, 768), dtype=np.float32)).to(dev).bfloat16()
    idx = torch.from_numpy(rng.integers(-1, 300, size=(5, 33)).astype(np.int32)).to(dev)
    out = torch.empty(5, 33, 768, device=dev, dtype=torch.bfloat16)
    mask = torch.empty(5, 33, dtype=torch.uint8, device=dev)
    ops.gather_l1norm(table, idx, out, mask, normalise=False)
    ref = table[idx.clamp_min(0).long()] * (idx >= 0)[..., None]
    assert torch.equal(out, ref.to(torch.bfloat16))
    assert torch.equal(mask.bool(), idx >= 0)


def test_gather_rejects_bad_din(dev):
    from segmminterest_b200 import _lib, ops
    t = torch.zeros(10, 6, device=dev)
    with pytest.raises(_lib.MMIError):
        ops.gather_l1norm(t, torch.zeros(4, dtype=torch.int32, device=dev), torch.empty(4, 6, device=dev), None)


# ----------------------------------------------------------------------------- GEMM (SIMT)
@pytest.mark.parametrize("layout", ["NT", "NN", "TN"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_simt_layouts(dev, layout, dtype):
    from segmminterest_b200 import ops
    torch.manual_seed(0)
    M, N, K = 333 * 4, 200, 136
    A = torch.randn(M, K, device=dev).to(dtype)
    Bm = torch.randn(N, K, device=dev).to(dtype)
    ref = A.double() @ Bm.double().T
    C = torch.empty(M, N, device=dev, dtype=dtype)
    if layout == "NT":
        ops.gemm(ops.GEMM_NT, ops.IMPL_SIMT, A, K, Bm, K, C, N, M, N, K)
    elif layout == "NN":
        Bt = Bm.T.contiguous()
        ops.gemm(ops.GEMM_NN, ops.IMPL_SIMT, A, K, Bt, N, C, N, M, N, K)
    else:
        At = A.T.contiguous()
        Bt = Bm.T.contiguous()
        C = torch.zeros(M, N, device=dev, dtype=torch.float32)
        ops.gemm(ops.GEMM_TN, ops.IMPL_SIMT, At, M, Bt, N, C, N, M, N, K, accumulate=True, split_k=3)
    tol = 2e-6 if dtype == torch.float32 or layout == "TN" else 4e-3
    assert _rel(C, ref) < tol


def test_gemm_simt_epilogue(dev):
    from segmminterest_b200 import ops
    torch.manual_seed(1)
    M, N, K, L = 96, 64, 48, 12
    A, W = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
    b, pe = torch.randn(N, device=dev), torch.randn(L, N, device=dev)
    C = torch.empty(M, N, device=dev)
    pre = torch.empty(M, N, device=dev)
    ops.gemm(ops.GEMM_NT, ops.IMPL_SIMT, A, K, W, K, C, N, M, N, K, bias=b, act=ops.ACT_GELU, preact=pre, add=pe, add_mod=L, ld_add=N)
    z = A.double() @ W.double().T + b.double()
    ref = torch.nn.functional.gelu(z) + pe.double().repeat(M // L, 1)
    assert _rel(pre, z) < 2e-6 and _rel(C, ref) < 2e-6
    # dgrad through GELU: y = (A W^T) * gelu'(Z) + R
    Z, R = torch.randn(M, N, device=dev), torch.randn(M, N, device=dev)
    ops.gemm(ops.GEMM_NT, ops.IMPL_SIMT, A, K, W, K, C, N, M, N, K, mul_gelu_grad=Z, add=R, add_mod=M, ld_add=N)
    zz = Z.double().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    ref = (A.double() @ W.double().T) * zz.grad + R.double()
    assert _rel(C, ref) < 2e-6


@pytest.mark.parametrize("layout", ["NT", "NN", "TN"])
@pytest.mark.parametrize("shape", [(333 * 4, 200, 136), (1000, 512, 640), (136, 72, 3072)])
def test_gemm_split_fp32_on_tensor_cores(dev, layout, shape):
    """fp32 mode on the tensor cores: fp32 operands as three bf16 terms (exact), six tcgen05 products -- hi*hi in one fp32
    accumulator, the five small ones in a second, K-slices of 512 joined by round-to-nearest atomics for TN -- including K
    that is not a multiple of the 64-wide k-block.  The tensor core truncates addends to the accumulator's ulp, so the bar
    grows with the length of the hi*hi chain (measured 4.7e-7 at K = 512, 3.2e-6 at K = 3072; FFMA kernel 4.1e-7 / 9.9e-7)."""
    from segmminterest_b200 import _lib, ops
    if not _lib.load().mmi_has_tc():
        pytest.skip("no tcgen05 device")
    torch.manual_seed(0)
    M, N, K = shape
    A = torch.randn(M, K, device=dev) * torch.logspace(-3, 3, K, device=dev)[None]       # wide dynamic range per column
    Bm = torch.randn(N, K, device=dev) / torch.logspace(-3, 3, K, device=dev)[None]
    ref = A.double() @ Bm.double().T
    ops.FP32_TC["on"] = True
    try:
        if layout == "NT":
            C = torch.empty(M, N, device=dev)
            ops.gemm(ops.GEMM_NT, ops.IMPL_SIMT, A, K, Bm, K, C, N, M, N, K)
        elif layout == "NN":
            C = torch.empty(M, N, device=dev)
            ops.gemm(ops.GEMM_NN, ops.IMPL_SIMT, A, K, Bm.T.contiguous(), N, C, N, M, N, K)
        else:
            C = torch.full((M, N), 0.5, device=dev)
            ops.gemm(ops.GEMM_TN, ops.IMPL_SIMT, A.T.contiguous(), M, Bm.T.contiguous(), N, C, N, M, N, K, accumulate=True, split_k=3)
            ref = ref + 0.5
    finally:
        ops.FP32_TC["on"] = False
    assert _rel(C, ref) < 2e-6 + 2e-9 * K


def test_gemm_split_fp32_epilogue(dev):
    """The fused epilogue of the split-fp32 path: bias, exact erf GELU + saved pre-activation, position-table add, and the
    dgrad form (gelu'(Z) multiply + residual), with dropout on one of them."""
    from segmminterest_b200 import _lib, ops
    if not _lib.load().mmi_has_tc():
        pytest.skip("no tcgen05 device")
    torch.manual_seed(1)
    M, N, K, L = 960, 64, 200, 12
    A, W = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
    b, pe = torch.randn(N, device=dev), torch.randn(L, N, device=dev)
    C, pre = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev)
    ops.FP32_TC["on"] = True
    try:
        ops.gemm(ops.GEMM_NT, ops.IMPL_SIMT, A, K, W, K, C, N, M, N, K, bias=b, act=ops.ACT_GELU, preact=pre, add=pe, add_mod=L, ld_add=N)
        z = A.double() @ W.double().T + b.double()
        ref = torch.nn.functional.gelu(z) + pe.double().repeat(M // L, 1)
        assert _rel(pre, z) < 5e-6 and _rel(C, ref) < 5e-6
        Z, R = torch.randn(M, N, device=dev), torch.randn(M, N, device=dev)
        ops.gemm(ops.GEMM_NT, ops.IMPL_SIMT, A, K, W, K, C, N, M, N, K, mul_gelu_grad=Z, add=R, add_mod=M, ld_add=N)
        zz = Z.double().requires_grad_(True)
        torch.nn.functional.gelu(zz).sum().backward()
        assert _rel(C, (A.double() @ W.double().T) * zz.grad + R.double()) < 5e-6
        # same call through the FFMA kernel: the two paths agree to the accumulation error above
        C2 = torch.empty_like(C)
        ops.FP32_TC["on"] = False
        ops.gemm(ops.GEMM_NT, ops.IMPL_SIMT, A, K, W, K, C2, N, M, N, K, mul_gelu_grad=Z, add=R, add_mod=M, ld_add=N)
        assert _rel(C, C2.double()) < 5e-6
    finally:
        ops.FP32_TC["on"] = False


# ----------------------------------------------------------------------------- LayerNorm, colsum, head
@pytest.mark.parametrize("d,rows", [(64, 777), (512, 777), (256, 5001), (768, 130), (1024, 33)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd(dev, d, rows, dtype):
    from segmminterest_b200 import _lib, ops
    torch.manual_seed(2)
    x = (torch.randn(rows, d, device=dev) * 2 + 0.5).to(dtype)
    g, b = torch.randn(d, device=dev), torch.randn(d, device=dev)
    dy, add = torch.randn(rows, d, device=dev).to(dtype), torch.randn(rows, d, device=dev).to(dtype)
    y = torch.empty_like(x)
    st = torch.empty(rows, 2, device=dev)
    ops.layernorm_fwd(x, rows, d, g, b, y, st)
    xr = x.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-12)
    tol = 2e-6 if dtype == torch.float32 else 6e-3
    assert _rel(y, yr) < tol
    yr.backward(dy.double())
    dx = torch.empty_like(x)
    dg, db = torch.ones(d, device=dev), torch.ones(d, device=dev)  # accumulate semantics
    ws = torch.empty(int(_lib.load().mmi_layernorm_bwd_workspace(d)), device=dev)
    ops.layernorm_bwd(dy, x, rows, d, g, st, add, dx, dg, db, ws)
    assert _rel(dx, xr.grad + add.double()) < tol
    assert _rel(dg - 1, gr.grad) < tol and _rel(db - 1, br.grad) < tol
    # fused column sums of dx (bias gradient of the Linear feeding the LayerNorm), accumulate semantics, no residual
    dx2, dxs = torch.empty_like(x), torch.full((d,), 2.0, device=dev)
    dg2, db2 = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    ops.layernorm_bwd(dy, x, rows, d, g, st, None, dx2, dg2, db2, ws, dxsum=dxs)
    assert _rel(dx2, xr.grad) < tol
    assert _rel(dg2, gr.grad) < tol and _rel(db2, br.grad) < tol
    assert _rel(dxs - 2, dx2.double().sum(0)) < 1e-5


def test_colsum_and_head(dev):
    from segmminterest_b200 import _lib, ops
    torch.manual_seed(3)
    M, N = 1234, 96
    x = torch.randn(M, 2 * N, device=dev)
    out = torch.ones(N, device=dev)
    ws = torch.empty(1 << 16, device=dev)
    ops.colsum_acc(x, M, N, 2 * N, out, ws)  # first N columns of a wider matrix
    assert _rel(out - 1, x[:, :N].double().sum(0)) < 2e-6
    d = 64
    X, w, b = torch.randn(M, d, device=dev), torch.randn(1, d, device=dev), torch.randn(1, device=dev)
    logits = torch.empty(M, device=dev)
    ops.head_fwd(X, M, d, w, b, logits)
    assert _rel(logits, X.double() @ w.double().T.squeeze(-1) + b.double()) < 2e-6
    dl, gs = torch.randn(M, device=dev), torch.tensor(0.5, device=dev)
    dx, dw, db = torch.empty_like(X), torch.zeros(1, d, device=dev), torch.zeros(1, device=dev)
    ops.head_bwd(X, M, d, w, dl, gs, dx, dw, db, torch.empty(int(_lib.load().mmi_head_bwd_workspace(d)), device=dev))
    assert _rel(dx, 0.5 * dl.double()[:, None] * w.double()) < 2e-6
    assert _rel(dw, 0.5 * (dl.double()[:, None] * X.double()).sum(0, keepdim=True)) < 2e-6
    assert _rel(db, 0.5 * dl.double().sum().reshape(1)) < 2e-6


# ----------------------------------------------------------------------------- attention
def _ref_attention(qa, ka, va, mka, qb, kb, vb, mkb, mq, H):
    """oracle.mmi_oracle.cross_attention's softmax part on explicit q/k/v (fp64)."""
    B, Lq, d = qa.shape
    dh = d // H

    def logits(q, k, mk):
        s = torch.einsum("bqhd,bkhd->bhqk", q.view(B, Lq, H, dh), k.view(B, -1, H, dh))
        m = (mq[:, :, None] & mk[:, None, :])[:, None].expand_as(s)
        return torch.where(m, s, torch.full_like(s, -10000.0))

    S = torch.cat([logits(qa, ka, mka), logits(qb, kb, mkb)], -1) / math.sqrt(dh)
    V = torch.cat([va, vb], 1).view(B, -1, H, dh)
    return torch.einsum("bhqk,bkhd->bqhd", S.softmax(-1), V).reshape(B, Lq, d)


@pytest.mark.parametrize("dh,dtype,impl,shape", [
    (32, torch.float32, "simt", (3, 2, 70, 40, 150)), (16, torch.float32, "simt", (3, 2, 70, 40, 150)),
    (32, torch.bfloat16, "simt", (3, 2, 70, 40, 150)),
    (32, torch.bfloat16, "tc", (3, 2, 70, 40, 150)), (32, torch.bfloat16, "tc", (2, 16, 500, 40, 500)),
    (32, torch.bfloat16, "tc", (2, 16, 40, 40, 500)), (32, torch.bfloat16, "tc", (1, 4, 129, 64, 65))])
def test_attention_fwd_bwd(dev, dh, dtype, impl, shape):
    _check_attention(dev, dh, dtype, impl, shape, ragged=True)


@pytest.mark.parametrize("shape", [(2, 16, 500, 40, 500), (2, 4, 40, 40, 100), (1, 2, 128, 64, 256), (1, 2, 300, 33, 1030), (1, 2, 1000, 40, 4000)])
def test_attention_tc_unmasked_fast_path(dev, shape):
    """Full histories (no masked key, no padded query): the mask-free fast path of the tcgen05 kernels."""
    _check_attention(dev, 32, torch.bfloat16, "tc", shape, ragged=False)


@pytest.mark.parametrize("ragged", [False, True])
def test_attention_tc_running_max_rescale(dev, ragged):
    """Keys whose logits grow tile after tile: the forward kernel's lazy running maximum has to move several
    times (O rescaled in TMEM), and must still equal the exact two-pass softmax."""
    _check_attention(dev, 32, torch.bfloat16, "tc", (2, 4, 200, 100, 700), ragged=ragged, key_ramp=0.012)


@pytest.mark.parametrize("shape", [(24, 16, 40, 40, 500), (12, 16, 200, 40, 300), (40, 8, 100, 100, 0), (21, 8, 40, 300, 0)])
def test_attention_allkeys_persistent_items(dev, shape):
    """More (b, h) items than SMs: every CTA of the persistent all-keys backward walks several items (barrier phases carried
    on global counters, K / V slots refilled while the previous item drains); one key block with a single key tile is the
    case where the refill has to wait for the previous item's last products."""
    _check_attention(dev, 32, torch.bfloat16, "tc", shape, ragged=True)


def _check_attention(dev, dh, dtype, impl, shape, ragged, key_ramp=0.0):
    from segmminterest_b200 import ops
    torch.manual_seed(4)
    B, H, Lq, La, Lb = shape
    d = H * dh
    impl = ops.IMPL_TC if impl == "tc" else ops.IMPL_SIMT

    def mk(L):
        n = torch.randint(1, L + 1, (B,)) if ragged and L else torch.full((B,), L)
        return (torch.arange(L)[None] < n[:, None])

    mq, mka, mkb = mk(Lq), mk(La), mk(Lb)
    mq[0, :] = True
    t = [torch.randn(B, L, d) * 0.7 for L in (Lq, La, La, Lq, Lb, Lb)]
    if key_ramp:   # |k_j| grows with j: later key tiles hold much larger logits than the first one
        t[0] *= 3.0; t[3] *= 3.0
        t[1] *= (1.0 + key_ramp * torch.arange(La))[None, :, None] ** 2
        t[4] *= (1.0 + key_ramp * (La + torch.arange(Lb)))[None, :, None] ** 2
    t = [x.to(dtype).to(dev) for x in t]
    qa, ka, va, qb, kb, vb = t
    ref_in = [x.double().cpu().requires_grad_(True) for x in t]
    ref = _ref_attention(ref_in[0], ref_in[1], ref_in[2], mka, ref_in[3], ref_in[4], ref_in[5], mkb, mq, H)
    out = torch.empty(B * Lq, d, device=dev, dtype=dtype)
    lse = torch.empty(B, H, Lq, device=dev)
    mqd, mkad, mkbd = [m.to(dev).view(torch.uint8) for m in (mq, mka, mkb)]
    blocks = [dict(q=(qa.data_ptr(), d), k=(ka.data_ptr(), d), v=(va.data_ptr(), d), mask_k=mkad, Lk=La),
              dict(q=(qb.data_ptr(), d), k=(kb.data_ptr(), d), v=(vb.data_ptr(), d), mask_k=mkbd, Lk=Lb)][:2 if Lb else 1]
    side = ops.AttnSide(ops.dt(out), impl, B, H, dh, Lq, mqd, out, d, lse, blocks)
    side.fwd()
    tol = 3e-6 if dtype == torch.float32 else 8e-3
    assert _rel(out.view(B, Lq, d), ref) < tol
    dO = (torch.randn(B, Lq, d) * 0.5).to(dtype).to(dev)
    ref.backward(dO.double().cpu())
    grads = [torch.zeros_like(x) for x in t]
    delta = torch.empty(B, H, Lq, device=dev)
    fused = impl == ops.IMPL_TC          # tensor-core kernels also accumulate the projections' bias gradients (column sums)
    db = [torch.full((d,), 0.25, device=dev) for _ in range(6)] if fused else [None] * 6
    ptr = [x.data_ptr() if x is not None else None for x in db]
    side.set_bwd(dO, d, delta, [dict(dq=(grads[0].data_ptr(), d), dk=(grads[1].data_ptr(), d), dv=(grads[2].data_ptr(), d),
                                     dbq=ptr[0], dbk=ptr[1], dbv=ptr[2]),
                                dict(dq=(grads[3].data_ptr(), d), dk=(grads[4].data_ptr(), d), dv=(grads[5].data_ptr(), d),
                                     dbq=ptr[3], dbk=ptr[4], dbv=ptr[5])][:len(blocks)])
    live = 3 * len(blocks)               # one key block only: block b's tensors are empty and untouched

    def check(tag):
        for g, r, name in list(zip(grads, ref_in, ["dqa", "dka", "dva", "dqb", "dkb", "dvb"]))[:live]:
            assert _rel(g, r.grad) < (5e-5 if dtype == torch.float32 else 1.5e-2), (tag, name)
        if fused:                        # += semantics (buffers started at 0.25), fp32 sums taken before the bf16 rounding
            for x, r, name in list(zip(db, ref_in, ["dbqa", "dbka", "dbva", "dbqb", "dbkb", "dbvb"]))[:live]:
                want = r.grad.sum((0, 1))
                # a column sum cancels heavily, so the bar is relative to the gradient it sums (random element errors of
                # relative size e give ||sum error|| ~ e * ||grad||_F); the gradients themselves are held to 1.5e-2, the sums to
                # 2.5e-2 (worst measured: 1.8e-2, dV bias sum of one 300-key block with ragged masks -- the bf16 rounding of P is
                # not sign-symmetric, so part of the element error adds up instead of cancelling)
                err = float((x.double().cpu() - 0.25 - want).norm())
                assert err < 2.5e-2 * float(r.grad.norm()) + 1e-20, (tag, name, err, float(r.grad.norm()), float(want.norm()))

    side.bwd_dq()
    for i in range(len(blocks)):
        side.bwd_dkv(i)
    check("dq + dkv kernels")
    if fused and dh == 32 and len(blocks) == 1:
        for g in grads:
            g.zero_()
        for x in db:
            x.fill_(0.25)
        assert side.bwd_all()
        check("all-keys kernel, one key block")
    elif fused and dh == 32:
        # the one-kernel backward (dK, dV and dQ per key block; dQ through the fp32 accumulator): same bars, and the
        # accumulator / counters are left zero
        for g in grads:
            g.zero_()
        for x in db:
            x.fill_(0.25)
        acc = [torch.zeros(B * Lq, d, device=dev) for _ in range(2)]
        cnt = [torch.zeros(B * H, device=dev, dtype=torch.int32) for _ in range(2)]
        side.set_fused(acc, cnt)
        delta.fill_(float("nan"))        # not read by the fused kernel
        assert side.bwd_fused(0) and side.bwd_fused(1)
        check("fused kernel")
        for a_, c_ in zip(acc, cnt):
            assert float(a_.abs().max()) == 0.0 and int(c_.abs().max()) == 0
        # the one-launch backward (one CTA per (b, h) owns every key of both blocks): same bars again
        for g in grads:
            g.zero_()
        for x in db:
            x.fill_(0.25)
        covered = side.bwd_all()
        assert covered == ((La + 127) // 128 + (Lb + 127) // 128 <= 5)
        if covered:
            check("all-keys kernel")


# ----------------------------------------------------------------------------- loss
def test_focal_loss_matches_reference_golden(dev):
    from segmminterest_b200 import ops
    z = np.load(os.path.join(GOLDEN, "loss_cases.npz"))
    logits = torch.from_numpy(z["logits"]).to(dev)
    gt = torch.from_numpy(z["gt_in"]).to(dev)
    B = logits.shape[0]
    ep = torch.from_numpy(z["exposure_prob"]).float().to(dev)
    scal, dl = torch.zeros(16, device=dev), torch.empty_like(logits)
    ops.focal_loss(logits, gt, ep, 1.0 / B, 1.0, True, scal, dl)
    s = scal.cpu().numpy()
    assert abs(s[0] - float(z["focal"])) < 2e-6 * abs(float(z["focal"]))
    assert abs(s[1] - float(z["mse"])) < 1e-5 * abs(float(z["mse"]))
    assert abs(s[2] - float(z["mse2"])) < 1e-5 * abs(float(z["mse2"]))
    assert np.array_equal(gt.cpu().numpy(), z["gt_out"])   # in-place rewrite, bit-exact
    assert _rel(dl, torch.from_numpy(z["grad_focal"])) < 5e-6


def test_interest_bpr_and_position_bias_match_reference_golden(dev):
    """mmi_loss_fwd_bwd with the reference's default loss `interestBPR` (+ focal) against the fixture produced by the
    unmodified reference (value and gradient), and the learnable position bias against autograd of the oracle."""
    from oracle import mmi_oracle
    from segmminterest_b200 import ops
    z = np.load(os.path.join(GOLDEN, "loss_cases.npz"))
    logits = torch.from_numpy(z["logits"]).to(dev)
    B, L = logits.shape
    ep = torch.from_numpy(z["exposure_prob"].astype(np.float32)).to(dev)
    for use_focal, use_bpr in ((False, True), (True, True)):
        gt = torch.from_numpy(z["gt_in"].copy()).to(dev)
        scal, dl = torch.zeros(16, device=dev), torch.empty_like(logits)
        ops.loss_fwd_bwd(logits, gt, ep, inv_bsz=1.0 / B, scalars=scal, dlogits=dl, use_focal=use_focal, w_focal=1.0, use_bpr=use_bpr,
                         w_bpr=1.0, rewrite_gt=True)
        s = scal.cpu().numpy()
        assert abs(s[4] - float(z["interestBPR"])) < 5e-6 * abs(float(z["interestBPR"]))
        want = z["grad_bpr"] + (z["grad_focal"] if use_focal else 0.0)
        assert _rel(dl, torch.from_numpy(want)) < 1e-5
        if use_focal:
            assert abs(s[3] - float(z["loss"])) < 5e-6 * abs(float(z["loss"]))
            assert np.array_equal(gt.cpu().numpy(), z["gt_out"])
        else:
            assert np.array_equal(gt.cpu().numpy(), z["gt_in"])   # no focal in the list: gt is left alone
    # learnable position bias: logits + (pos+1) * w + b, gradients of w and b (accumulate semantics)
    torch.manual_seed(11)
    bw, bb = (1 + 0.1 * torch.randn(L)).to(dev), (1 + 0.1 * torch.randn(L)).to(dev)
    gt = torch.from_numpy(z["gt_in"].copy()).to(dev)
    scal, dl, lo = torch.zeros(16, device=dev), torch.empty_like(logits), torch.empty_like(logits)
    dbw, dbb = torch.full((L,), 3.0, device=dev), torch.full((L,), -1.0, device=dev)
    ops.loss_fwd_bwd(logits, gt, ep, inv_bsz=1.0 / B, scalars=scal, dlogits=dl, use_focal=True, w_focal=0.7, use_bpr=True, w_bpr=1.3,
                     rewrite_gt=True, bias_weight=bw, bias_bias=bb, logits_out=lo, dbias_weight=dbw, dbias_bias=dbb)
    lr = torch.from_numpy(z["logits"]).double().requires_grad_(True)
    bwr, bbr = bw.double().cpu().requires_grad_(True), bb.double().cpu().requires_grad_(True)
    full = lr + (torch.arange(L, dtype=torch.float64) + 1) * bwr + bbr
    out = mmi_oracle.compute_loss(full, torch.from_numpy(z["gt_in"]), list(z["exposure_prob"]), ("focal", "interestBPR"),
                                  {"focal": 0.7, "interestBPR": 1.3})
    out["loss"].backward()
    assert _rel(lo, full.detach()) < 1e-6
    assert abs(scal[3].item() - out["loss"].item()) < 1e-5 * abs(out["loss"].item())
    assert _rel(dl, lr.grad) < 1e-5
    assert _rel(dbw - 3.0, bwr.grad) < 1e-4 and _rel(dbb + 1.0, bbr.grad) < 1e-4


@pytest.mark.parametrize("i", range(10))
def test_every_selectable_loss_matches_reference_golden(dev, i):
    """huber / hazard / surviveCE / interestCE / interestKL alone and mixed with focal (whose in-place gt rewrite the
    later losses see) and interestBPR: values, weighted total, d loss / d logits and the rewritten gt against the
    fixture produced by the unmodified reference (models/decoder_leave_focal.py:528-566)."""
    import json
    from segmminterest_b200 import ops
    from segmminterest_b200.model import LOSS_SLOT
    z = np.load(os.path.join(GOLDEN, "loss_cases_all.npz"))
    base = np.load(os.path.join(GOLDEN, "loss_cases.npz"))
    tag, lst, mask_loss = json.loads(str(z["variants"]))[i]
    wts = json.loads(str(z["weights"]))
    logits = torch.from_numpy(z[f"{tag}/logits"]).to(dev)
    gt = torch.from_numpy(base["gt_in"].copy()).to(dev)
    B = logits.shape[0]
    ep = torch.from_numpy(base["exposure_prob"].astype(np.float32)).to(dev)
    scal, dl = torch.zeros(16, device=dev), torch.empty_like(logits)
    others = {n: wts["mse" if n == "huber" else n] for n in lst if n not in ("focal", "interestBPR")}

    def after_focal(n):
        return n in lst and "focal" in lst and lst.index("focal") < lst.index(n)

    ops.loss_fwd_bwd(logits, gt, ep, inv_bsz=1.0 / B, scalars=scal, dlogits=dl, use_focal="focal" in lst, w_focal=wts["focal"],
                     use_bpr="interestBPR" in lst, w_bpr=wts["interestBPR"], rewrite_gt=True, others=others, mask_loss=mask_loss,
                     ce_after_focal=after_focal("interestCE"), kl_after_focal=after_focal("interestKL"))
    s = scal.cpu().numpy()
    for n in lst:
        want = float(z[f"{tag}/{n}"])
        assert abs(s[LOSS_SLOT[n]] - want) < 2e-5 * abs(want), (tag, n, s[LOSS_SLOT[n]], want)
    for n, slot in (("mse", 1), ("mse2", 2), ("loss", 3)):
        want = float(z[f"{tag}/{n}"])
        assert abs(s[slot] - want) < 2e-5 * abs(want), (tag, n, s[slot], want)
    assert np.array_equal(gt.cpu().numpy(), z[f"{tag}/gt_out"])
    assert _rel(dl, torch.from_numpy(z[f"{tag}/grad"])) < 2e-5, tag


# ----------------------------------------------------------------------------- validation metrics (a-15 / 8f-4)
def test_eval_metrics_match_reference_golden_and_oracle(dev):
    """mmi_eval_metrics: ProbAUC (exact Mann-Whitney == sklearn.roc_auc_score) and the per-row metrics against the lists
    the unmodified reference's main_eval_batch collected, then the drop-in main_eval_batch protocol, then a larger
    random batch (with exact ties) against the oracle."""
    from oracle import mmi_oracle
    from segmminterest_b200.evaluation import DeviceMetrics, main_eval_batch
    from types import SimpleNamespace
    z = np.load(os.path.join(GOLDEN, "eval_cases.npz"))
    base = np.load(os.path.join(GOLDEN, "loss_cases.npz"))
    logits = torch.from_numpy(base["logits"]).to(dev)
    gt = torch.from_numpy(base["gt_in"]).to(dev)
    rows, out = DeviceMetrics()(logits, gt, list(base["exposure_prob"]))
    r = rows.cpu().numpy()
    for name, col in (("LeaveMSE", 0), ("view_lengths", 1), ("duration_lengths", 2), ("LeaveCTR", 3), ("LeaveCTR_view", 4), ("JaccardSim", 5)):
        assert np.allclose(r[:, col], z[name], rtol=2e-5, atol=2e-6, equal_nan=True), name
    assert abs(out[0].item() - float(z["ProbAUC"][0])) < 1e-6
    # the driver protocol: results_list keys select the metrics; same lists as the reference
    res = {k: [] for k in ("ProbAUC", "JaccardSim", "LeaveMSE", "view_lengths", "duration_lengths", "LeaveCTR", "LeaveCTR_view")}
    res = main_eval_batch(SimpleNamespace(TOP_K_mask=0, draw_case=0), torch.from_numpy(z["interests"]).to(dev), gt, None, res)
    for k in res:
        assert np.allclose(np.asarray(res[k]), z[k], rtol=2e-5, atol=2e-6, equal_nan=True), k
    # test_type='old' (the interests are the survival probabilities, my_evaluation.py:270-271) + the `logits=` MAES
    # bookkeeping (:309-320), against the unmodified reference's lists
    res = {k: [] for k in ("ProbAUC", "JaccardSim", "LeaveMSE", "view_lengths", "LeaveCTR", "LeaveCTR_view")}
    res["MAES"] = 0.0
    res["pred_leave"] = []
    res = main_eval_batch(SimpleNamespace(TOP_K_mask=0, draw_case=0), torch.from_numpy(z["interests"]).to(dev), gt, None, res,
                          test_type="old", logits=logits * 0.125)
    for k in ("ProbAUC", "JaccardSim", "LeaveMSE", "view_lengths", "LeaveCTR", "LeaveCTR_view"):
        assert np.allclose(np.asarray(res[k]), z["old/" + k], rtol=2e-5, atol=2e-6, equal_nan=True), k
    assert abs(res["MAES"] - float(z["old/MAES"])) < 1e-9 and np.array_equal(torch.stack(res["pred_leave"]).numpy(), z["old/pred_leave"])
    # larger batch, quantised logits => many exact ties
    rng = np.random.default_rng(17)
    from segmminterest_b200 import synth
    B = 700
    gt2 = synth.make_labels(rng, rng.integers(1, 41, size=B))
    lg2 = (np.round(rng.standard_normal((B, 40)) * 4) / 4).astype(np.float32)
    rows2, out2 = DeviceMetrics()(torch.from_numpy(lg2).to(dev), torch.from_numpy(gt2).to(dev), [1.0] * 40)
    want = mmi_oracle.prob_auc_batch(lg2, gt2)
    assert abs(out2[0].item() - want) < 2e-6, (out2[0].item(), want)
    wrows = mmi_oracle.eval_rows(torch.sigmoid(torch.from_numpy(lg2)), torch.from_numpy(gt2)).numpy()
    assert np.allclose(rows2.cpu().numpy(), wrows, rtol=3e-5, atol=3e-6, equal_nan=True)


# ----------------------------------------------------------------------------- clip + AdamW
def test_clip_adamw_matches_torch(dev):
    from segmminterest_b200 import _lib, ops
    torch.manual_seed(5)
    n = 100_003
    p = torch.randn(n, device=dev)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-4)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    norm = torch.zeros(2, device=dev)
    ws = torch.empty(int(_lib.load().mmi_clip_adamw_workspace(n)), device=dev)
    lp = torch.empty(n, device=dev, dtype=torch.bfloat16)
    for step in (1, 2, 3):
        g = torch.randn(n, device=dev) * (0.5 if step == 2 else 0.01)
        ref.grad = g.clone()
        nref = torch.nn.utils.clip_grad_norm_([ref], 10.0)
        opt.step()
        ops.clip_adamw(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-4, 10.0, step, norm, lp, ws)
        assert abs(norm[0].item() - nref.item()) < 1e-5 * nref.item()
        assert torch.allclose(p, ref.detach(), rtol=1e-6, atol=1e-7)
        assert torch.equal(lp, p.bfloat16())


def test_cast_bf16_transpose(dev):
    from segmminterest_b200 import ops
    x = torch.randn(70, 130, device=dev)
    y = torch.empty(130, 70, device=dev, dtype=torch.bfloat16)
    ops.cast_bf16(x, y, 70, 130, transpose=True)
    assert torch.equal(y, x.T.contiguous().bfloat16())


# ----------------------------------------------------------------------------- loader drop-in (a-1 / a-2)
def test_device_frame_loader_reproduces_reference_batch(dev):
    """DeviceFrameLoader against the batch the unmodified FrameDatasetSeq_SegMM + DataCollator produced
    (tests/golden/gather_small.npz): all twelve keys, bit-exact, in the reference's key order; then the fused
    L1 normalisation against main...SegMM.py:272-273."""
    import json
    import pandas as pd
    from segmminterest_b200 import DeviceFrameLoader
    z = np.load(os.path.join(GOLDEN, "gather_small.npz"))
    rows = json.loads(str(z["rows_json"]))
    df = pd.DataFrame(rows, columns=["user_id", "video_id", "time_ms", "duration_ms", "playing_time_x", "label_1D", "history_items",
                                     "history_playing", "history_lengths"])

    class Corpus:
        data_df = {"test": df}
        user_input_dict = json.loads(str(z["user_input_json"]))

    table = torch.from_numpy(z["table"]).to(dev)
    kw = dict(phase="test", batch_size=3, user2id=json.loads(str(z["user2id_json"])), item2id=json.loads(str(z["item2id_json"])))
    ld = DeviceFrameLoader(Corpus(), json.loads(str(z["lineid_json"])), table, **kw)
    batches = list(ld)
    assert len(batches) == len(ld) == 2
    ref_keys = [k[4:] for k in z.files if k.startswith("out/")]
    assert [k for k in batches[0] if k not in ("usr_idx", "vid_idx")] == ref_keys       # DataCollator's key order
    for k in ref_keys:
        got = torch.cat([b[k] for b in batches]).cpu().numpy()
        assert got.dtype == z["out/" + k].dtype and np.array_equal(got, z["out/" + k]), k
    ldn = DeviceFrameLoader(Corpus(), json.loads(str(z["lineid_json"])), table, normalise=True, **kw)
    user = torch.cat([b["user"] for b in ldn]).cpu()
    ref = torch.from_numpy(z["out/user"])
    ref = ref / (ref.norm(p=1, dim=-1, keepdim=True) + 1e-6)
    assert _rel(user, ref) < 1e-6
