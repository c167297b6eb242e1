"""Orchestrates the MMinterest step on the C-ABI kernels: flat parameter / gradient
buffers, activation workspace, forward, loss, hand-written backward.

Data layout in HBM:
  * tokens are row-major [B*L, d]; candidate ("vid", L=40) and history ("usr", L=Lt) sides
    are separate tensors because they use different weights;
  * all live parameters sit in ONE flat fp32 buffer in forward order; the six projections
    that read the same input are adjacent, so `W6 = flat[off : off+6*d*d].view(6d, d)` IS the
    fused projection weight (no concat, no copy) and the weight gradient GEMM writes the
    matching slice of the flat gradient buffer; nn.Parameters are views of the flat buffer.
  * parameters the reference never trains (layer N-1, history side of layer N-2, pe_lns,
    txt_lvl_projs, patch_merge -- SURVEY section 0 fact 5) stay ordinary Parameters outside
    the flat buffer: no gradient, no AdamW update, exactly like the reference.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib, ops
from .ops import ACT_GELU, ACT_NONE, GEMM_NN, GEMM_NT, GEMM_TN, IMPL_SIMT, IMPL_TC

ALIGN = 64  # floats (256 B): every parameter group starts on a TMA-friendly boundary


@dataclass
class EngineConfig:
    d_model: int = 512
    nhead: int = 16
    num_layers: int = 6
    din_vid: int = 1024
    din_usr: int = 1024
    max_usr_len: int = 100
    max_vid_len: int = 40
    use_pe: bool = True
    precision: str = "fp32"   # 'fp32' (FFMA, strict parity) | 'bf16' (tcgen05 tensor cores)


class _Slot:
    __slots__ = ("name", "param", "off", "numel", "shape")

    def __init__(self, name, param, off):
        self.name, self.param, self.off = name, param, off
        self.numel, self.shape = param.numel(), tuple(param.shape)


class Engine:
    def __init__(self, cfg: EngineConfig, model, device):
        _lib.load()
        self.cfg, self.model, self.device = cfg, model, device
        if cfg.precision not in ("fp32", "bf16"):
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {cfg.precision!r}")
        if cfg.d_model % cfg.nhead or (cfg.d_model // cfg.nhead) not in (16, 32):
            raise NotImplementedError("head dim must be 16 or 32")
        if cfg.d_model % 4 or cfg.d_model > 1024:
            raise NotImplementedError("d_model must be a multiple of 4 and <= 1024")
        self.act_dtype = torch.float32 if cfg.precision == "fp32" else torch.bfloat16
        self.slots: list[_Slot] = []
        self.groups = {}
        self._layout()
        self.flat = None
        self.flat_grad = None
        self.flat_lp = None    # bf16 shadow of `flat` (bf16 mode)
        self.flat_lpT = None   # per-matrix transposed bf16 shadow (dgrad operand of the TC path)
        self.anchor = torch.zeros((), device=device, requires_grad=True)
        self._ws = {}
        self._saved = None
        self.use_tc = cfg.precision == "bf16" and bool(_lib.load().mmi_has_tc())
        n_red = max(_lib.load().mmi_layernorm_bwd_workspace(cfg.d_model), _lib.load().mmi_head_bwd_workspace(cfg.d_model),
                    4 * cfg.d_model * max(cfg.max_usr_len, cfg.max_vid_len), 1 << 20)
        self.red_ws = torch.empty(int(n_red), device=device, dtype=torch.float32)
        self.scalars = torch.zeros(8, device=device, dtype=torch.float32)

    # ------------------------------------------------------------------ parameter layout
    def _layout(self):
        cfg, bb = self.cfg, self.model.backbone1
        N = cfg.num_layers
        off = 0

        def group(key, params):
            nonlocal off
            off = (off + ALIGN - 1) // ALIGN * ALIGN
            start = off
            for name, p in params:
                self.slots.append(_Slot(name, p, off))
                off += p.numel()
            self.groups[key] = (start, off - start)

        P = "backbone1."
        group("vid_proj.w", [(P + "vid_proj.weight", bb.vid_proj.weight)])
        group("vid_proj.b", [(P + "vid_proj.bias", bb.vid_proj.bias)])
        group("usr_proj.w", [(P + "usr_proj.weight", bb.usr_proj.weight)])
        group("usr_proj.b", [(P + "usr_proj.bias", bb.usr_proj.bias)])
        if cfg.use_pe:
            group("vid_pe", [(P + "vid_pe.weight", bb.vid_pe.weight)])
            group("usr_pe", [(P + "usr_pe.weight", bb.usr_pe.weight)])
        for s in ("vid", "usr"):
            ln = getattr(bb, s + "_ln")
            group(f"{s}_ln.g", [(f"{P}{s}_ln.weight", ln.weight)])
            group(f"{s}_ln.b", [(f"{P}{s}_ln.bias", ln.bias)])
        for i in range(N - 1):
            L = bb.encoder.layers[i]
            ca = L.cross_attn
            full = i < N - 2
            q = f"{P}encoder.layers.{i}."
            # six projections of the candidate tokens / of the history tokens, adjacent
            vid6 = [("v2v", 0), ("v2v", 1), ("v2v", 2), ("t2v", 0), ("v2t", 1), ("v2t", 2)]
            usr6 = [("t2v", 1), ("t2v", 2), ("v2t", 0), ("t2t", 0), ("t2t", 1), ("t2t", 2)]
            if not full:
                vid6, usr6 = vid6[:4], usr6[:2]
            for side, six in (("vid", vid6), ("usr", usr6)):
                group(f"L{i}.{side}.w6", [(f"{q}cross_attn.{b}_proj.{j}.weight", getattr(ca, b + "_proj")[j].weight) for b, j in six])
                group(f"L{i}.{side}.b6", [(f"{q}cross_attn.{b}_proj.{j}.bias", getattr(ca, b + "_proj")[j].bias) for b, j in six])
            for side in (("vid", "usr") if full else ("vid",)):
                ff, ln = getattr(ca, "ff_" + side), getattr(ca, "ln_" + side)
                group(f"L{i}.{side}.wo", [(f"{q}cross_attn.ff_{side}.weight", ff.weight)])
                group(f"L{i}.{side}.bo", [(f"{q}cross_attn.ff_{side}.bias", ff.bias)])
                group(f"L{i}.{side}.ln1.g", [(f"{q}cross_attn.ln_{side}.weight", ln.weight)])
                group(f"L{i}.{side}.ln1.b", [(f"{q}cross_attn.ln_{side}.bias", ln.bias)])
                mlp, ln2 = getattr(L, "ff_" + side), getattr(L, "ln_" + side)
                group(f"L{i}.{side}.w1", [(f"{q}ff_{side}.layers.0.weight", mlp.layers[0].weight)])
                group(f"L{i}.{side}.b1", [(f"{q}ff_{side}.layers.0.bias", mlp.layers[0].bias)])
                group(f"L{i}.{side}.w2", [(f"{q}ff_{side}.layers.1.weight", mlp.layers[1].weight)])
                group(f"L{i}.{side}.b2", [(f"{q}ff_{side}.layers.1.bias", mlp.layers[1].bias)])
                group(f"L{i}.{side}.ln2.g", [(f"{q}ln_{side}.weight", ln2.weight)])
                group(f"L{i}.{side}.ln2.b", [(f"{q}ln_{side}.bias", ln2.bias)])
        group("head.w", [("stage_mlp1.weight", self.model.stage_mlp1.weight)])
        group("head.b", [("stage_mlp1.bias", self.model.stage_mlp1.bias)])
        if getattr(self.model, "bias_weight", None) is not None:   # learnable position bias (decoder_leave_focal.py:442-444)
            group("bias_weight", [("bias_weight", self.model.bias_weight)])
            group("bias_bias", [("bias_bias", self.model.bias_bias)])
        self.n_flat = (off + ALIGN - 1) // ALIGN * ALIGN
        self.live_names = [s.name for s in self.slots]

    def ensure_bound(self):
        """(Re)creates the flat buffers when the Parameters are not views of them (first call,
        after .cuda()/.to(), after someone replaced .data)."""
        ok = self.flat is not None
        if ok:
            base = self.flat.data_ptr()
            for s in self.slots:
                if s.param.data_ptr() != base + 4 * s.off or s.param.dtype != torch.float32:
                    ok = False
                    break
        if ok:
            return
        flat = torch.zeros(self.n_flat, device=self.device, dtype=torch.float32)
        for s in self.slots:
            flat[s.off:s.off + s.numel].copy_(s.param.data.reshape(-1).to(self.device, torch.float32))
            s.param.data = flat[s.off:s.off + s.numel].view(s.shape)
        self.flat = flat
        self.flat_grad = torch.zeros_like(flat)
        if self.cfg.precision == "bf16":
            self.flat_lp = torch.zeros(self.n_flat, device=self.device, dtype=torch.bfloat16)
            self.flat_lpT = torch.zeros(self.n_flat, device=self.device, dtype=torch.bfloat16)
        for s in self.slots:
            s.param.grad = None

    def bind_grads(self):
        """Makes every live Parameter's .grad a view of the flat gradient buffer; zeroes the
        buffer when the caller dropped the grads (optimizer.zero_grad(set_to_none=True))."""
        views_ok = True
        base = self.flat_grad.data_ptr()
        for s in self.slots:
            g = s.param.grad
            if g is None or g.data_ptr() != base + 4 * s.off:
                views_ok = False
                break
        if views_ok:
            return
        self.flat_grad.zero_()
        for s in self.slots:
            s.param.grad = self.flat_grad[s.off:s.off + s.numel].view(s.shape)

    # views into the flat buffers ------------------------------------------------------
    def w(self, key, lp=False):
        off, n = self.groups[key]
        buf = self.flat_lp if lp else self.flat
        return buf[off:off + n]

    def wT(self, key):
        off, n = self.groups[key]
        return self.flat_lpT[off:off + n]

    def g(self, key):
        off, n = self.groups[key]
        return self.flat_grad[off:off + n]

    def refresh_low_precision(self):
        """bf16 shadow of the parameters (+ per-matrix transposes for the dgrad operand)."""
        if self.cfg.precision != "bf16":
            return
        ops.cast_bf16(self.flat, self.flat_lp, 1, self.n_flat)
        if self.use_tc:
            d = self.cfg.d_model
            for key, (off, n) in self.groups.items():
                if key.endswith((".w6", ".wo", ".w1", ".w2")):
                    rows = n // d
                    ops.cast_bf16(self.flat[off:off + n], self.flat_lpT[off:off + n], rows, d, transpose=True)

    # ------------------------------------------------------------------ workspace
    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(shape, device=self.device, dtype=dtype)
            self._ws[key] = t
        return t

    def _impl(self, M, N, K, layout):
        """tcgen05 path whenever TMA's 16-byte stride rule holds (all model dims that are multiples
        of 8); otherwise the SIMT kernel (still CUDA, still ours)."""
        if not self.use_tc:
            return IMPL_SIMT
        if layout == GEMM_NT:
            ok = K % 8 == 0
        else:  # TN: A [K, M], B [K, N]
            ok = M % 8 == 0 and N % 8 == 0
        return IMPL_TC if ok else IMPL_SIMT

    # ------------------------------------------------------------------ building blocks
    def _linear(self, x, M, K, wkey, bkey, N, out, *, act=ACT_NONE, preact=None, add=None, add_mod=0, ld_add=0,
                n_prefix=None):
        """out[M,N] = act(x[M,K] W[N,K]^T + b) (+ add).  On the tensor-core path `preact` receives
        gelu'(z) instead of z (the backward epilogue then needs no erf; see mmi_gemm save_act_grad)."""
        lp = self.cfg.precision == "bf16"
        W = self.w(wkey, lp)
        ops.gemm(GEMM_NT, self._impl(M, N, K, GEMM_NT), x, K, W, K, out, N, M, N, K, bias=self.w(bkey), act=act,
                 preact=preact, add=add, add_mod=add_mod, ld_add=ld_add, save_act_grad=self.use_tc and preact is not None)

    def _linear_bwd(self, dy, x, M, N, K, wkey, bkey, dx, *, mul_gelu_grad=None, add=None, need_dx=True, bias_done=False):
        """dW += dy^T x ; db += colsum(dy) ; dx = dy W (* gelu'(mul)) (+ add).  bias_done: the kernel that produced dy
        (LayerNorm backward) already accumulated the bias gradient."""
        lp = self.cfg.precision == "bf16"
        if not bias_done:
            ops.colsum_acc(dy, M, N, N, self.g(bkey), self.red_ws)
        impl_w = self._impl(N, K, M, GEMM_TN)
        if impl_w == IMPL_TC:
            split = 0  # auto: fill the SMs
        else:
            tiles = ((N + 127) // 128) * ((K + 127) // 128)
            split = max(1, min(32, (2 * 148) // tiles, (M + 4095) // 4096))
        ops.gemm(GEMM_TN, impl_w, dy, N, x, K, self.g(wkey), K, N, K, M, accumulate=True, split_k=split, out_dtype=_lib.F32)
        if need_dx:
            if self._impl(M, K, N, GEMM_NT) == IMPL_TC:
                ops.gemm(GEMM_NT, IMPL_TC, dy, N, self.wT(wkey), N, dx, K, M, K, N, mul_gelu_grad=mul_gelu_grad, add=add,
                         add_mod=M if add is not None else 0, ld_add=K, mul_is_grad=self.use_tc)
            else:
                ops.gemm(GEMM_NN, IMPL_SIMT, dy, N, self.w(wkey, lp), K, dx, K, M, K, N, mul_gelu_grad=mul_gelu_grad,
                         add=add, add_mod=M if add is not None else 0, ld_add=K, mul_is_grad=self.use_tc)

    # ------------------------------------------------------------------ forward
    def forward(self, usr_image, usr_mask, vid_image, vid_mask):
        """usr_image [B,Lt,Din] / vid_image [B,Lv,Din] already L1-normalised (the driver does it,
        main...SegMM.py:272-273; our own data path fuses it into the gather).  Returns fp32
        logits [B, Lv] (a workspace tensor: clone before the next call)."""
        cfg = self.cfg
        ops._need_cuda(usr_image, vid_image, usr_mask, vid_mask)
        self.ensure_bound()
        B, Lt, _ = usr_image.shape
        Lv = vid_image.shape[1]
        if Lt > cfg.max_usr_len or Lv > cfg.max_vid_len:
            raise ValueError(f"sequence longer than the position table: Lt={Lt} (max {cfg.max_usr_len}), Lv={Lv} (max {cfg.max_vid_len})")
        T = self.act_dtype
        d, H, N = cfg.d_model, cfg.nhead, cfg.num_layers
        self.refresh_low_precision()
        x_in = {"usr": usr_image.to(T).contiguous().view(B * Lt, -1), "vid": vid_image.to(T).contiguous().view(B * Lv, -1)}
        mask = {"usr": usr_mask.to(torch.bool).contiguous().view(torch.uint8), "vid": vid_mask.to(torch.bool).contiguous().view(torch.uint8)}
        Ls = {"usr": Lt, "vid": Lv}
        Ts = {"usr": B * Lt, "vid": B * Lv}
        din = {"usr": cfg.din_usr, "vid": cfg.din_vid}
        sv = {"B": B, "Lt": Lt, "Lv": Lv, "x_in": x_in, "mask": mask, "layers": []}
        X = {}
        for s in ("vid", "usr"):
            e = self._buf(f"emb_pre.{s}", (Ts[s], d), T)
            st = self._buf(f"emb_st.{s}", (Ts[s], 2), torch.float32)
            x0 = self._buf(f"x0.{s}", (Ts[s], d), T)
            pe = self.w(f"{s}_pe") if cfg.use_pe else None
            self._linear(x_in[s], Ts[s], din[s], f"{s}_proj.w", f"{s}_proj.b", d, e, add=pe, add_mod=Ls[s] if pe is not None else 0, ld_add=d)
            ops.layernorm_fwd(e, Ts[s], d, self.w(f"{s}_ln.g"), self.w(f"{s}_ln.b"), x0, st)
            sv[f"emb_pre.{s}"], sv[f"emb_st.{s}"] = e, st
            X[s] = x0
        esz = x0.element_size()
        for i in range(N - 1):
            full = i < N - 2
            nq = {"vid": 6 if full else 4, "usr": 6 if full else 2}
            lay = {"full": full, "nq": nq, "x": dict(X)}
            qkv = {}
            for s in ("vid", "usr"):
                qkv[s] = self._buf(f"qkv.{i}.{s}", (Ts[s], nq[s] * d), T)
                self._linear(X[s], Ts[s], d, f"L{i}.{s}.w6", f"L{i}.{s}.b6", nq[s] * d, qkv[s])
            lay["qkv"] = qkv

            def col(s, j):
                return (qkv[s].data_ptr() + j * d * esz, nq[s] * d)

            sides = ("vid", "usr") if full else ("vid",)
            attn = {}
            for s in sides:
                a_out = self._buf(f"attn.{i}.{s}", (Ts[s], d), T)
                lse = self._buf(f"lse.{i}.{s}", (B, H, Ls[s]), torch.float32)
                if s == "vid":
                    blocks = [dict(q=col("vid", 0), k=col("vid", 1), v=col("vid", 2), mask_k=mask["vid"], Lk=Lv),
                              dict(q=col("vid", 3), k=col("usr", 0), v=col("usr", 1), mask_k=mask["usr"], Lk=Lt)]
                else:
                    blocks = [dict(q=col("usr", 2), k=col("vid", 4), v=col("vid", 5), mask_k=mask["vid"], Lk=Lv),
                              dict(q=col("usr", 3), k=col("usr", 4), v=col("usr", 5), mask_k=mask["usr"], Lk=Lt)]
                attn_impl = IMPL_TC if (self.use_tc and d // H == 32 and d % 8 == 0) else IMPL_SIMT
                side = ops.AttnSide(ops.dt(a_out), attn_impl, B, H, d // H, Ls[s], mask[s], a_out, d, lse, blocks)
                side.fwd()
                attn[s] = (side, a_out, lse)
            lay["attn"] = attn
            for s in sides:
                p1 = self._buf(f"p1.{i}.{s}", (Ts[s], d), T)
                st1 = self._buf(f"st1.{i}.{s}", (Ts[s], 2), torch.float32)
                x1 = self._buf(f"x1.{i}.{s}", (Ts[s], d), T)
                z1 = self._buf(f"z1.{i}.{s}", (Ts[s], d), T)
                g1 = self._buf(f"g1.{i}.{s}", (Ts[s], d), T)
                p2 = self._buf(f"p2.{i}.{s}", (Ts[s], d), T)
                st2 = self._buf(f"st2.{i}.{s}", (Ts[s], 2), torch.float32)
                x2 = self._buf(f"x2.{i}.{s}", (Ts[s], d), T)
                self._linear(attn[s][1], Ts[s], d, f"L{i}.{s}.wo", f"L{i}.{s}.bo", d, p1, add=X[s], add_mod=Ts[s], ld_add=d)
                ops.layernorm_fwd(p1, Ts[s], d, self.w(f"L{i}.{s}.ln1.g"), self.w(f"L{i}.{s}.ln1.b"), x1, st1)
                self._linear(x1, Ts[s], d, f"L{i}.{s}.w1", f"L{i}.{s}.b1", d, g1, act=ACT_GELU, preact=z1)
                self._linear(g1, Ts[s], d, f"L{i}.{s}.w2", f"L{i}.{s}.b2", d, p2, add=x1, add_mod=Ts[s], ld_add=d)
                ops.layernorm_fwd(p2, Ts[s], d, self.w(f"L{i}.{s}.ln2.g"), self.w(f"L{i}.{s}.ln2.b"), x2, st2)
                lay[s] = dict(p1=p1, st1=st1, x1=x1, z1=z1, g1=g1, p2=p2, st2=st2)
                X[s] = x2
            sv["layers"].append(lay)
        sv["x_out"] = X["vid"]
        logits = self._buf("logits", (B, Lv), torch.float32)
        ops.head_fwd(X["vid"], Ts["vid"], d, self.w("head.w"), self.w("head.b"), logits)
        self._saved = sv
        return logits

    # ------------------------------------------------------------------ loss
    def loss(self, logits, gt, exposure_prob, inv_bsz, loss_cfg=None, bpr_scale=1.0, need_grad=True):
        """Fused loss (focal and / or interestBPR) + learnable position bias + diagnostics + dlogits
        (models/decoder_leave_focal.py:490-572).  `gt` is rewritten in place like the reference (:534-535) when
        focal is on.  Returns (scalars [focal, mse, mse2, loss, interestBPR, ...], logits incl. bias): device
        tensors, no sync.  loss_cfg: dict(use_focal, w_focal, use_bpr, w_bpr); None = focal with weight 1."""
        B, L = logits.shape
        if gt.dtype != torch.int64 or not gt.is_contiguous() or not gt.is_cuda:
            raise ValueError("gt must be a contiguous CUDA int64 tensor (it is rewritten in place, like the reference)")
        cfg = loss_cfg or dict(use_focal=True, w_focal=1.0, use_bpr=False, w_bpr=0.0)
        ep = self._ws.get("ep")
        if ep is None or self._ws.get("ep_src") != tuple(exposure_prob[:L]):
            ep = torch.tensor(list(exposure_prob[:L]), device=self.device, dtype=torch.float32)
            self._ws["ep"], self._ws["ep_src"] = ep, tuple(exposure_prob[:L])
        dlogits = self._buf("dlogits", (B, L), torch.float32)
        has_bias = "bias_weight" in self.groups
        logits_out = self._buf("logits_b", (B, L), torch.float32) if has_bias else None
        if has_bias and need_grad:
            self.bind_grads()
        ops.loss_fwd_bwd(logits, gt, ep, inv_bsz=inv_bsz, scalars=self.scalars, dlogits=dlogits, use_focal=cfg["use_focal"],
                         w_focal=cfg["w_focal"], use_bpr=cfg["use_bpr"], w_bpr=cfg["w_bpr"], bpr_scale=bpr_scale, rewrite_gt=True,
                         bias_weight=self.w("bias_weight") if has_bias else None, bias_bias=self.w("bias_bias") if has_bias else None,
                         logits_out=logits_out, dbias_weight=self._pending_bias_grad(0) if has_bias and need_grad else None,
                         dbias_bias=self._pending_bias_grad(1) if has_bias and need_grad else None)
        if self._saved is not None:
            self._saved["dlogits"] = dlogits
        return self.scalars, (logits_out if has_bias else logits)

    def _pending_bias_grad(self, which):
        """The loss kernel produces d loss / d bias directly; it lands in a scratch buffer and is added to the flat
        gradient in backward() scaled by the upstream gradient (so autograd's grad_out is honoured)."""
        t = self._buf("dbias", (2, self.groups["bias_weight"][1]), torch.float32)
        if which == 0:
            t.zero_()
        return t[which]

    # ------------------------------------------------------------------ backward
    def backward(self, gscale=None, on_ready=None):
        """Accumulates d(loss)/d(param) into the flat gradient buffer.  `gscale` is the upstream
        gradient of the loss (device scalar) -- never read on the host."""
        sv = self._saved
        if sv is None or "dlogits" not in sv:
            raise RuntimeError("backward() called before forward()+loss()")
        self.bind_grads()
        cfg = self.cfg
        T = self.act_dtype
        d, H, N = cfg.d_model, cfg.nhead, cfg.num_layers
        B, Lt, Lv = sv["B"], sv["Lt"], sv["Lv"]
        Ls = {"usr": Lt, "vid": Lv}
        Ts = {"usr": B * Lt, "vid": B * Lv}
        din = {"usr": cfg.din_usr, "vid": cfg.din_vid}
        mask = sv["mask"]

        def scratch(name, s, width=d):
            return self._buf(f"bw.{name}.{s}", (Ts[s], width), T)

        if "bias_weight" in self.groups:   # d loss / d (bias_weight, bias_bias) from the loss kernel, times the upstream gradient
            db = self._buf("dbias", (2, self.groups["bias_weight"][1]), torch.float32)
            gs = gscale if gscale is not None else 1.0
            self.g("bias_weight").add_(db[0] * gs)
            self.g("bias_bias").add_(db[1] * gs)
        dX = {"vid": scratch("dx", "vid"), "usr": None}
        ops.head_bwd(sv["x_out"], Ts["vid"], d, self.w("head.w"), sv["dlogits"], gscale, dX["vid"], self.g("head.w"),
                     self.g("head.b"), self.red_ws)
        if on_ready is not None:
            on_ready(self.groups["head.w"][0])
        for i in reversed(range(N - 1)):
            lay = sv["layers"][i]
            full, nq, Xin = lay["full"], lay["nq"], lay["x"]
            sides = ("vid", "usr") if full else ("vid",)
            dP1, dA = {}, {}
            for s in sides:
                a = lay[s]
                pre = f"L{i}.{s}."
                dp2 = scratch("dp2", s)
                ops.layernorm_bwd(dX[s], a["p2"], Ts[s], d, self.w(pre + "ln2.g"), a["st2"], None, dp2, self.g(pre + "ln2.g"),
                                  self.g(pre + "ln2.b"), self.red_ws, dxsum=self.g(pre + "b2"))
                dz1 = scratch("dz1", s)
                self._linear_bwd(dp2, a["g1"], Ts[s], d, d, pre + "w2", pre + "b2", dz1, mul_gelu_grad=a["z1"], bias_done=True)
                dx1 = scratch("dx1", s)
                self._linear_bwd(dz1, a["x1"], Ts[s], d, d, pre + "w1", pre + "b1", dx1, add=dp2)
                dp1 = scratch("dp1", s)
                ops.layernorm_bwd(dx1, a["p1"], Ts[s], d, self.w(pre + "ln1.g"), a["st1"], None, dp1, self.g(pre + "ln1.g"),
                                  self.g(pre + "ln1.b"), self.red_ws, dxsum=self.g(pre + "bo"))
                da = scratch("da", s)
                self._linear_bwd(dp1, lay["attn"][s][1], Ts[s], d, d, pre + "wo", pre + "bo", da, bias_done=True)
                dP1[s], dA[s] = dp1, da
            dqkv = {s: scratch("dqkv", s, nq[s] * d) for s in ("vid", "usr")}
            esz = dqkv["vid"].element_size()

            def gcol(s, j):
                return (dqkv[s].data_ptr() + j * d * esz, nq[s] * d)

            for s in sides:
                side = lay["attn"][s][0]
                delta = self._buf(f"delta.{s}", (B, H, Ls[s]), torch.float32)
                if s == "vid":
                    grads = [dict(dq=gcol("vid", 0), dk=gcol("vid", 1), dv=gcol("vid", 2)),
                             dict(dq=gcol("vid", 3), dk=gcol("usr", 0), dv=gcol("usr", 1))]
                else:
                    grads = [dict(dq=gcol("usr", 2), dk=gcol("vid", 4), dv=gcol("vid", 5)),
                             dict(dq=gcol("usr", 3), dk=gcol("usr", 4), dv=gcol("usr", 5))]
                side.set_bwd(dA[s], d, delta, grads)
                side.bwd_dq()
                side.bwd_dkv(0)
                side.bwd_dkv(1)
            new_dX = {}
            for s in ("vid", "usr"):
                pre = f"L{i}.{s}."
                out = scratch("dx", s) if dX[s] is None or True else None
                # dX[s] (the grad w.r.t. this layer's OUTPUT) is dead by now: dp2 consumed it
                self._linear_bwd(dqkv[s], Xin[s], Ts[s], nq[s] * d, d, pre + "w6", pre + "b6", out, add=dP1.get(s))
                new_dX[s] = out
            dX = new_dX
            if on_ready is not None:
                on_ready(self.groups[f"L{i}.vid.w6"][0])  # first group of layer i in the flat layout
        for s in ("vid", "usr"):
            de = self._buf(f"bw.de.{s}", (Ts[s], d), T)
            if dX[s] is None:
                continue
            ops.layernorm_bwd(dX[s], sv[f"emb_pre.{s}"], Ts[s], d, self.w(f"{s}_ln.g"), sv[f"emb_st.{s}"], None, de,
                              self.g(f"{s}_ln.g"), self.g(f"{s}_ln.b"), self.red_ws, dxsum=self.g(f"{s}_proj.b"))
            if cfg.use_pe:
                # d pe[l,:] = sum_b dE[b,l,:]  -> column sums of dE viewed as [B, L*d]
                ops.colsum_acc(de, B, Ls[s] * d, Ls[s] * d, self.g(f"{s}_pe")[: Ls[s] * d], self.red_ws)
            self._linear_bwd(de, sv["x_in"][s], Ts[s], d, din[s], f"{s}_proj.w", f"{s}_proj.b", None, need_dx=False, bias_done=True)
        if on_ready is not None:
            on_ready(0)
        self._saved = None
