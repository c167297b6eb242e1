"""Thin tensor -> pointer wrappers over the C ABI (include/mmi_b200.h).  torch is used for
device memory and the current stream only; every computation below is one of our kernels."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .profiler import TIMER
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, BF16, F32, GEMM_NN, GEMM_NT, GEMM_TN, IMPL_SIMT, IMPL_TC  # noqa: F401

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def dt(t: torch.Tensor) -> int:
    return _DT[t.dtype]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MMIError("segmminterest_b200 kernels need CUDA tensors; there is no CPU fallback")


def _drop(site):
    """DropSite (dropout.py) or None -> mmi_dropout (thr8 = 0: off)"""
    d = _lib.Dropout()
    if site is not None and site.thr8:
        d.key, d.thr8, d.scale = site.key, site.thr8, site.scale
    else:
        d.key, d.thr8, d.scale = 0, 0, 1.0
    return d


def _drop_ptr(site):
    return C.byref(_drop(site)) if (site is not None and site.thr8) else None


def dropout_mask(site, row0, rows, cols, mask, group0=0):
    """test hook: mask[r, c] = 1 iff element (row0 + r, c) of the site survives"""
    _need_cuda(mask)
    assert mask.dtype == torch.uint8 and mask.is_contiguous() and mask.numel() == rows * cols
    d = _drop(site)
    rc = _lib.load().mmi_dropout_mask(C.byref(d), row0, rows, cols, group0, mask.data_ptr(), _stream())
    _lib.check(rc, "mmi_dropout_mask")
    LaunchCounter.n += 1


class LaunchCounter:
    """Counts kernel launches issued through this module (bench.py `gpu_launches`)."""
    n = 0


def gather_l1norm(table: torch.Tensor, idx: torch.Tensor, out: torch.Tensor, mask: torch.Tensor | None,
                  normalise: bool = True, n_valid: int | None = None):
    """n_valid (profiling only): number of ids >= 0, so that the timer credits the ALGORITHMIC bytes -- table rows are read
    for valid ids only; every output row, id and mask byte is written / read."""
    _need_cuda(table, idx, out)
    lib = _lib.load()
    assert idx.dtype == torch.int32 and idx.is_contiguous() and table.is_contiguous() and out.is_contiguous()
    n_tokens = idx.numel()
    nv_ = n_tokens if n_valid is None else int(n_valid)
    work = float(table.shape[1]) * (nv_ * table.element_size() + n_tokens * out.element_size()) + 5.0 * n_tokens
    with TIMER.region("gather", work):
        rc = lib.mmi_gather_l1norm_fwd(table.data_ptr(), dt(table), table.shape[0], table.shape[1], idx.data_ptr(), n_tokens,
                                       out.data_ptr(), dt(out), _ptr(mask), 1 if normalise else 0, _stream())
    _lib.check(rc, "mmi_gather_l1norm_fwd")
    LaunchCounter.n += 1


# fp32 GEMMs routed to the split-bf16 tensor-core path (set by the engine for precision="fp32" on a tcgen05 device;
# MMI_FP32_TC=0 keeps the FFMA kernels).  One shared workspace, grown to the largest product seen.
FP32_TC = {"on": False, "ws": None}


def _split_ok(layout, M, N, K, lda, ldb):
    """shapes the tcgen05 kernel takes: N a multiple of 8 (fp32 epilogue pairs + TMA strides of a TN operand), M >= 1"""
    if layout == GEMM_TN:
        return N % 8 == 0 and M % 8 == 0
    return N % 8 == 0


def gemm(layout, impl, A, lda, B, ldb, Cm, ldc, M, N, K, *, bias=None, act=ACT_NONE, preact=None, mul_gelu_grad=None,
         add=None, add_mod=0, ld_add=0, accumulate=False, split_k=1, in_dtype=None, out_dtype=None, save_act_grad=False,
         mul_is_grad=False, drop=None, mul_scale=1.0):
    lib = _lib.load()
    a = _lib.GemmArgs()
    in_dt = dt(A) if in_dtype is None else in_dtype
    if impl == IMPL_SIMT and FP32_TC["on"] and in_dt == F32 and (dt(Cm) if out_dtype is None else out_dtype) == F32 and _split_ok(layout, M, N, K, lda, ldb):
        # strict-parity mode on the tensor cores: fp32 operands as three bf16 terms each, six products in fp32 accumulators
        impl = IMPL_TC
        need = int(lib.mmi_gemm_split_workspace(layout, M, N, K))
        ws = FP32_TC.get("ws")
        dev = getattr(A, "device", None) or (ws.device if ws is not None else torch.device("cuda", torch.cuda.current_device()))
        if ws is None or ws.numel() < need or ws.device != dev:
            FP32_TC["ws"] = ws = None          # release before growing
            FP32_TC["ws"] = ws = torch.empty(need, dtype=torch.uint8, device=dev)
        a.split_ws, a.split_ws_bytes = ws.data_ptr(), ws.numel()
    a.layout, a.impl = layout, impl
    a.in_dtype = in_dt
    a.out_dtype = dt(Cm) if out_dtype is None else out_dtype
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.B, a.ldb, a.C, a.ldc = A.data_ptr(), lda, B.data_ptr(), ldb, Cm.data_ptr(), ldc
    a.bias = _ptr(bias)
    a.act = act
    a.preact = _ptr(preact)
    a.ld_preact = N if preact is not None else 0
    a.mul_gelu_grad = _ptr(mul_gelu_grad)
    a.ld_mul = N if mul_gelu_grad is not None else 0
    a.add = _ptr(add)
    a.ld_add = ld_add
    a.add_mod = add_mod
    a.add_dtype = dt(add) if add is not None else F32
    a.accumulate = 1 if accumulate else 0
    a.split_k = split_k
    a.save_act_grad = 1 if save_act_grad else 0
    a.mul_is_grad = int(mul_is_grad)        # 0: gelu'(operand), 1: operand is gelu'(z), 2: operand is a ReLU (+ dropout) output
    a.mul_scale = float(mul_scale)
    a.drop = _drop(drop)
    cat = "gemm_tc" if impl == IMPL_TC else "gemm_simt"
    if TIMER.detail:
        cat += f" {('NT', 'NN', 'TN')[layout]} M={M} N={N} K={K}" + (" gelu" if act else "") + (" mulgelu" if mul_gelu_grad is not None else "") \
            + (" add" if add is not None else "") + (" acc" if accumulate else "") + (" drop" if a.drop.thr8 else "")
    with TIMER.region(cat, 2.0 * M * N * K):
        rc = lib.mmi_gemm(C.byref(a), _stream())
    _lib.check(rc, "mmi_gemm")
    LaunchCounter.n += 1


def adaptive_pool_fwd(x, B, L, d, out_len, y):
    """AdaptiveAvgPool1d(out_len) along the token axis: x [B, L, d] -> y [B, out_len, d]"""
    with TIMER.region("pool"):
        rc = _lib.load().mmi_adaptive_pool_fwd(x.data_ptr(), dt(x), B, L, d, out_len, y.data_ptr(), _stream())
    _lib.check(rc, "mmi_adaptive_pool_fwd")
    LaunchCounter.n += 1


def adaptive_pool_bwd(dy, B, L, d, out_len, dx):
    with TIMER.region("pool"):
        rc = _lib.load().mmi_adaptive_pool_bwd(dy.data_ptr(), dt(dy), B, L, d, out_len, dx.data_ptr(), _stream())
    _lib.check(rc, "mmi_adaptive_pool_bwd")
    LaunchCounter.n += 1


def colsum_acc(x, M, N, ldx, out, ws):
    with TIMER.region("colsum"):
        rc = _lib.load().mmi_colsum_acc(x.data_ptr(), dt(x), M, N, ldx, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "mmi_colsum_acc")
    LaunchCounter.n += 2


def layernorm_fwd(x, rows, d, gamma, beta, y, stats, eps=1e-12, drop=None):
    """drop (DropSite): y = dropout(LN(x)) -- the embedding dropout (encoder.py:386,472)"""
    with TIMER.region("ln_fwd"):
        rc = _lib.load().mmi_layernorm_fwd_drop(x.data_ptr(), dt(x), rows, d, gamma.data_ptr(), beta.data_ptr(), eps, y.data_ptr(),
                                                _ptr(stats), _drop_ptr(drop), _stream())
    _lib.check(rc, "mmi_layernorm_fwd")
    LaunchCounter.n += 1


def layernorm_bwd(dy, x, rows, d, gamma, stats, add, dx, dgamma, dbeta, ws, dxsum=None, dy_drop=None, dx_drop=None,
                  dx_dropped=None):
    """dxsum (optional, fp32 [d]) += column sums of dx: the bias gradient of the Linear feeding this LayerNorm.
    dy_drop: dy is read through the dropout mask of that site (backward of y = dropout(LN(x))).
    dx_drop + dx_dropped: also writes mask * scale * dx, the gradient into the Linear whose dropped-out output was added to
    the residual; dxsum then sums that tensor."""
    lib = _lib.load()
    assert ws.numel() >= lib.mmi_layernorm_bwd_workspace(d)
    dxp = _drop_ptr(dx_drop)
    with TIMER.region("ln_bwd"):
        rc = lib.mmi_layernorm_bwd_drop(dy.data_ptr(), x.data_ptr(), dt(x), rows, d, gamma.data_ptr(), stats.data_ptr(), _ptr(add),
                                        dx.data_ptr(), _ptr(dgamma), _ptr(dbeta), _ptr(dxsum), ws.data_ptr(), _drop_ptr(dy_drop),
                                        dxp, _ptr(dx_dropped) if dxp is not None else None, _stream())
    _lib.check(rc, "mmi_layernorm_bwd")
    LaunchCounter.n += 2


class AttnSide:
    """One query side of the 4-way attention: two key blocks sharing a softmax."""

    def __init__(self, dtype, impl, B, H, dh, Lq, mask_q, out, ldo, lse, blocks, drop=None):
        a = _lib.AttnArgs()
        a.drop = _drop(drop)           # logits dropout (encoder.py:145-150); the backward calls reuse it
        a.dtype, a.impl, a.B, a.H, a.dh, a.Lq = dtype, impl, B, H, dh, Lq
        a.mask_q = mask_q.data_ptr()
        a.nblk = len(blocks)
        for i, b in enumerate(blocks):
            k = a.blk[i]
            k.q, k.ldq = b["q"]
            k.k, k.ldk = b["k"]
            k.v, k.ldv = b["v"]
            k.mask_k = b["mask_k"].data_ptr()
            k.Lk = b["Lk"]
        a.out, a.ldo = out.data_ptr(), ldo
        a.lse = lse.data_ptr()
        self.a = a

    def flops(self, which=None):
        """QK^T + PV MACs*2 over all (padded) positions of the given key block(s)."""
        a = self.a
        lk = sum(a.blk[i].Lk for i in range(a.nblk)) if which is None else a.blk[which].Lk
        return 4.0 * a.B * a.H * a.Lq * lk * a.dh

    def _cat(self, name, which=None):
        if not TIMER.detail:
            return name
        a = self.a
        lk = "+".join(str(a.blk[i].Lk) for i in range(a.nblk)) if which is None else str(a.blk[which].Lk)
        return f"{name} B={a.B} H={a.H} Lq={a.Lq} Lk={lk}"

    def fwd(self):
        with TIMER.region(self._cat("attn_fwd"), self.flops()):
            rc = _lib.load().mmi_attn_fwd(C.byref(self.a), _stream())
        _lib.check(rc, "mmi_attn_fwd")
        LaunchCounter.n += 1

    def set_bwd(self, dout, lddo, delta, grads):
        a = self.a
        a.dout, a.lddo, a.delta = dout.data_ptr(), lddo, delta.data_ptr()
        for i, g in enumerate(grads):
            k = a.blk[i]
            k.dq, k.lddq = g["dq"]
            k.dk, k.lddk = g["dk"]
            k.dv, k.lddv = g["dv"]
            k.dbq, k.dbk, k.dbv = g.get("dbq"), g.get("dbk"), g.get("dbv")   # fused bias-gradient sums (TC path), None = off

    def set_fused(self, dq_acc, dq_count):
        """Workspace of mmi_attn_bwd_fused: per key block an fp32 [B*Lq, H*dh] accumulator and B*H int32 counters, all zero
        on entry (the kernel leaves them zero)."""
        for i in range(self.a.nblk):
            self.a.dq_acc[i] = dq_acc[i].data_ptr()
            self.a.dq_count[i] = dq_count[i].data_ptr()

    def bwd_fused(self, which) -> bool:
        """dq, dk, dv (+ bias-gradient sums) of key block `which` in one kernel; False when the library asks for the
        two-kernel path (rc 1)."""
        with TIMER.region(self._cat("attn_bwd_fused", which), 2.5 * self.flops(which)):
            rc = _lib.load().mmi_attn_bwd_fused(C.byref(self.a), which, _stream())
        if rc == 1:
            return False
        _lib.check(rc, "mmi_attn_bwd_fused")
        LaunchCounter.n += 1
        return True

    def bwd_all(self) -> bool:
        """dq, dk, dv of BOTH key blocks (+ bias-gradient sums) in one launch, one CTA per (b, h); False when the shapes are
        not covered (more than 5 key tiles) and the caller has to use the other kernels."""
        with TIMER.region(self._cat("attn_bwd_all"), 2.5 * self.flops()):
            rc = _lib.load().mmi_attn_bwd_all(C.byref(self.a), _stream())
        if rc == 1:
            return False
        _lib.check(rc, "mmi_attn_bwd_all")
        LaunchCounter.n += 1
        return True

    def bwd_dq(self):
        with TIMER.region(self._cat("attn_bwd_dq"), 1.5 * self.flops()):
            rc = _lib.load().mmi_attn_bwd_dq(C.byref(self.a), _stream())
        _lib.check(rc, "mmi_attn_bwd_dq")
        LaunchCounter.n += 1

    def bwd_dkv(self, which):
        with TIMER.region(self._cat("attn_bwd_dkv", which), 2.0 * self.flops(which)):
            rc = _lib.load().mmi_attn_bwd_dkv(C.byref(self.a), which, _stream())
        _lib.check(rc, "mmi_attn_bwd_dkv")
        LaunchCounter.n += 1


def head_fwd(x, rows, d, w, b, logits, add=None):
    """logits[r] = w . x[r] (+ b) (+ add[r])"""
    with TIMER.region("head"):
        rc = _lib.load().mmi_head_fwd(x.data_ptr(), dt(x), rows, d, w.data_ptr(), _ptr(b), _ptr(add), logits.data_ptr(), _stream())
    _lib.check(rc, "mmi_head_fwd")
    LaunchCounter.n += 1


def head_bwd(x, rows, d, w, dlogits, gscale, dx, dw, db, ws):
    lib = _lib.load()
    assert ws.numel() >= lib.mmi_head_bwd_workspace(d)
    with TIMER.region("head"):
        rc = lib.mmi_head_bwd(x.data_ptr(), dt(x), rows, d, w.data_ptr(), dlogits.data_ptr(), _ptr(gscale), dx.data_ptr(),
                              dw.data_ptr(), _ptr(db), ws.data_ptr(), _stream())
    _lib.check(rc, "mmi_head_bwd")
    LaunchCounter.n += 2


def focal_loss(logits, gt, exposure_prob, inv_bsz, weight, rewrite_gt, scalars, dlogits):
    B, L = logits.shape
    assert gt.dtype == torch.int64 and gt.is_contiguous() and logits.is_contiguous() and logits.dtype == torch.float32
    with TIMER.region("loss"):
        rc = _lib.load().mmi_focal_loss_fwd_bwd(logits.data_ptr(), gt.data_ptr(), B, L, exposure_prob.data_ptr(), inv_bsz, weight,
                                                1 if rewrite_gt else 0, scalars.data_ptr(), dlogits.data_ptr(), _stream())
    _lib.check(rc, "mmi_focal_loss_fwd_bwd")
    LaunchCounter.n += 1


def loss_fwd_bwd(logits, gt, exposure_prob, *, inv_bsz, scalars, dlogits, use_focal=True, w_focal=1.0, use_bpr=False, w_bpr=1.0,
                 bpr_scale=1.0, rewrite_gt=True, bias_weight=None, bias_bias=None, logits_out=None, dbias_weight=None,
                 dbias_bias=None, others=None, mask_loss=0, ce_after_focal=False, kl_after_focal=False):
    """Fused loss (any mix of focal, interestBPR and -- `others` = {name: weight} -- huber, hazard, surviveCE, interestCE,
    interestKL) + learnable position bias + diagnostics + dlogits.  scalars: fp32[16] (layout in include/mmi_b200.h)."""
    B, L = logits.shape
    if scalars.numel() < 16:
        raise ValueError("loss scalars buffer must hold 16 floats")
    assert gt.dtype == torch.int64 and gt.is_contiguous() and logits.is_contiguous() and logits.dtype == torch.float32
    a = _lib.LossArgs()
    a.logits, a.gt, a.B, a.L = logits.data_ptr(), gt.data_ptr(), B, L
    a.exposure_prob = _ptr(exposure_prob)
    a.bias_weight, a.bias_bias = _ptr(bias_weight), _ptr(bias_bias)
    a.inv_bsz, a.w_focal, a.w_bpr, a.bpr_scale = inv_bsz, w_focal, w_bpr, bpr_scale
    a.use_focal, a.use_bpr, a.rewrite_gt = int(use_focal), int(use_bpr), int(rewrite_gt)
    a.logits_out, a.scalars, a.dlogits = _ptr(logits_out), scalars.data_ptr(), dlogits.data_ptr()
    a.dbias_weight, a.dbias_bias = _ptr(dbias_weight), _ptr(dbias_bias)
    for name, wt in (others or {}).items():
        if name not in ("huber", "hazard", "surviveCE", "interestCE", "interestKL"):
            raise ValueError(f"unknown loss {name!r}")
        setattr(a, "use_" + name, 1)
        setattr(a, "w_" + name, float(wt))
    a.mask_loss, a.ce_after_focal, a.kl_after_focal = int(bool(mask_loss)), int(bool(ce_after_focal)), int(bool(kl_after_focal))
    with TIMER.region("loss"):
        rc = _lib.load().mmi_loss_fwd_bwd(C.byref(a), _stream())
    _lib.check(rc, "mmi_loss_fwd_bwd")
    LaunchCounter.n += 1


def id_embed_fwd(table, ids, B, L, d, out, frame_w=None, frame_b=None, pe=None, frame_pos=None):
    """frame_pos: fp32 [B, L] positions fed to the frame projection instead of 0..L-1 (the 'noPos' ablation)"""
    assert ids.dtype == torch.int64 and ids.is_contiguous() and table.dtype == torch.float32
    assert frame_pos is None or (frame_pos.dtype == torch.float32 and frame_pos.is_contiguous() and frame_pos.numel() == B * L)
    with TIMER.region("id_embed"):
        rc = _lib.load().mmi_id_embed_fwd(table.data_ptr(), table.shape[0], table.shape[1], ids.data_ptr(), B, L, d, _ptr(frame_w),
                                          _ptr(frame_b), _ptr(pe), _ptr(frame_pos), out.data_ptr(), dt(out), _stream())
    _lib.check(rc, "mmi_id_embed_fwd")
    LaunchCounter.n += 1


def id_embed_bwd(de, ids, n_rows, tw, B, L, d, dtable, dframe_w=None, dframe_b=None, frame_pos=None):
    with TIMER.region("id_embed"):
        rc = _lib.load().mmi_id_embed_bwd(de.data_ptr(), dt(de), ids.data_ptr(), n_rows, tw, B, L, d, dtable.data_ptr(), _ptr(dframe_w),
                                          _ptr(dframe_b), _ptr(frame_pos), _stream())
    _lib.check(rc, "mmi_id_embed_bwd")
    LaunchCounter.n += 1


def id_rows_bwd(de, tw, B, L, d, rows, dframe_w=None, dframe_b=None, frame_pos=None):
    """rows [B, tw] fp32 = per-interaction gradient rows of an ID table (no table atomics): the row-sparse exchange of dp.py"""
    with TIMER.region("id_embed"):
        rc = _lib.load().mmi_id_rows_bwd(de.data_ptr(), dt(de), tw, B, L, d, rows.data_ptr(), _ptr(dframe_w), _ptr(dframe_b),
                                         _ptr(frame_pos), _stream())
    _lib.check(rc, "mmi_id_rows_bwd")
    LaunchCounter.n += 1


def scatter_rows_add(ids, rows, tw, n_rows, dtable):
    """dtable[ids[i], :] += rows[i, :]"""
    assert ids.dtype == torch.int64 and ids.is_contiguous() and rows.dtype == torch.float32 and rows.is_contiguous()
    with TIMER.region("id_embed"):
        rc = _lib.load().mmi_scatter_rows_add(ids.data_ptr(), rows.data_ptr(), ids.numel(), tw, n_rows, dtable.data_ptr(), _stream())
    _lib.check(rc, "mmi_scatter_rows_add")
    LaunchCounter.n += 1


def rowdot_fwd(t, ldt, y, ldy, R, C_, out, add1=None, add2=None):
    with TIMER.region("head"):
        rc = _lib.load().mmi_rowdot_fwd(t.data_ptr(), ldt, y.data_ptr(), ldy, dt(t), R, C_, _ptr(add1), _ptr(add2), out.data_ptr(), _stream())
    _lib.check(rc, "mmi_rowdot_fwd")
    LaunchCounter.n += 1


def rowdot_bwd(g, gscale, t, ldt, y, ldy, R, C_, dt_out, dy_out, dy_add=None):
    with TIMER.region("head"):
        rc = _lib.load().mmi_rowdot_bwd(g.data_ptr(), _ptr(gscale), t.data_ptr(), ldt, y.data_ptr(), ldy, dt(t), R, C_, _ptr(dy_add),
                                        dt_out.data_ptr(), dy_out.data_ptr(), _stream())
    _lib.check(rc, "mmi_rowdot_bwd")
    LaunchCounter.n += 1


def clip_adamw(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, wd, max_norm, step, norm_out, bf16_out, ws):
    lib = _lib.load()
    n = params.numel()
    assert ws.numel() >= lib.mmi_clip_adamw_workspace(n)
    with TIMER.region("clip_adamw"):
        rc = lib.mmi_clip_adamw(params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), n, lr, beta1, beta2,
                                eps, wd, max_norm, step, norm_out.data_ptr(), _ptr(bf16_out), ws.data_ptr(), _stream())
    _lib.check(rc, "mmi_clip_adamw")
    LaunchCounter.n += 3


def cast_bf16(src, dst, rows, cols, transpose=False):
    with TIMER.region("cast"):
        rc = _lib.load().mmi_cast_bf16(src.data_ptr(), dst.data_ptr(), rows, cols, 1 if transpose else 0, _stream())
    _lib.check(rc, "mmi_cast_bf16")
    LaunchCounter.n += 1


def eval_metrics(logits, gt, exposure_prob, rows, out, workspace, interests=False, old=False):
    """Device validation metrics (mmi_eval_metrics): rows [B,6] = pred_view_length, view_length, duration, LeaveCTR,
    LeaveCTR_view, JaccardSim; out[0] = ProbAUC of the batch."""
    B, L = logits.shape
    _need_cuda(logits, gt, exposure_prob, rows, out, workspace)
    assert logits.dtype == torch.float32 and logits.is_contiguous() and gt.dtype == torch.int64 and gt.is_contiguous()
    assert rows.numel() >= B * 6 and out.numel() >= 4
    assert workspace.numel() * workspace.element_size() >= _lib.load().mmi_eval_metrics_workspace(B, L)
    with TIMER.region("eval_metrics"):
        rc = _lib.load().mmi_eval_metrics(logits.data_ptr(), gt.data_ptr(), B, L, _ptr(exposure_prob), (2 if old else 1) if interests else 0,
                                          workspace.data_ptr(), rows.data_ptr(), out.data_ptr(), _stream())
    _lib.check(rc, "mmi_eval_metrics")
    LaunchCounter.n += 3
