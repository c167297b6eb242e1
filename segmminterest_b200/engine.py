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

import os
from dataclasses import dataclass

import torch

from . import _lib, ops
from .dropout import DropSite, quantise, site_key
from .ops import ACT_GELU, ACT_NONE, ACT_RELU, GEMM_NN, GEMM_NT, GEMM_TN, IMPL_SIMT, IMPL_TC

ALIGN = 64  # floats (256 B): every parameter group starts on a TMA-friendly boundary

# The four attention blocks of SegFormerXAttention (models/encoder.py:17-24): projection j of block n (0 = query, 1 = key,
# 2 = value).  Q_SIDE / KEY_SIDE: which tokens the query / the key+value projections read.  PROJ_ORDER: the order in which
# the projections of one token side sit in the flat parameter buffer, so that the ones in use form ONE fused GEMM.
Q_SIDE = {"v2v": "vid", "t2v": "vid", "v2t": "usr", "t2t": "usr"}
KEY_SIDE = {"v2v": "vid", "t2v": "usr", "v2t": "vid", "t2t": "usr"}
PROJ_ORDER = {"vid": [("v2v", 0), ("v2v", 1), ("v2v", 2), ("t2v", 0), ("v2t", 1), ("v2t", 2)],
              "usr": [("t2v", 1), ("t2v", 2), ("v2t", 0), ("t2t", 0), ("t2t", 1), ("t2t", 2)]}
# blocks that feed the candidate / history queries per ablation (encoder.py:108-161; 'SelfAtt' returns no history update)
ATTN_BLOCKS = {"ours": {"vid": ("v2v", "t2v"), "usr": ("v2t", "t2t")},
               "CrossAtt": {"vid": ("t2v",), "usr": ("v2t",)},
               "SelfAtt": {"vid": ("v2v",), "usr": ()}}
MLP_ABLATIONS = ("SelfMLP", "CrossMLP", "w/oAtt")


def attn_ablation(ablation_type):
    """the reference tests substrings for the attention ablations ('noUser_SelfAtt' is SelfAtt for the model; 'noUser' lives
    in the driver; 'noPos' keeps the full encoder) and equality for the MLP ablations (encoder.py:392-400,503-511)"""
    a = ablation_type or "ours"
    if a in MLP_ABLATIONS:
        return a
    return "CrossAtt" if "CrossAtt" in a else ("SelfAtt" if "SelfAtt" in a else "ours")


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
    ablation: str = "ours"    # 'ours' | 'CrossAtt' | 'SelfAtt' | 'SelfMLP' | 'CrossMLP' | 'w/oAtt'
    no_pos: bool = False      # 'noPos': a fresh random permutation of the frame positions per call (ID-input towers only)


class _View:
    """A column window of a workspace tensor (pointer + dtype only; the leading dimension is passed separately)."""
    __slots__ = ("_ptr", "dtype")

    def __init__(self, ptr, like):
        self._ptr, self.dtype = ptr, like.dtype

    def data_ptr(self):
        return self._ptr


class _Slot:
    __slots__ = ("name", "param", "off", "numel", "shape")

    def __init__(self, name, param, off):
        self.name, self.param, self.off = name, param, off
        self.numel, self.shape = param.numel(), tuple(param.shape)


class _Tower:
    """One backbone (SegFormerX parameter container) and how its two inputs arrive: 'image' = [B,L,Din] features through a
    Linear projection, 'id' = int64 ids through an embedding table (models/encoder.py:352-362,426-435)."""

    def __init__(self, tag, prefix, module):
        import torch.nn as nn
        self.tag, self.prefix, self.m = tag, prefix, module
        self.vid_kind = "id" if isinstance(module.vid_proj, nn.Embedding) else "image"
        self.usr_kind = "id" if isinstance(module.usr_proj, nn.Embedding) else "image"

    def k(self, name):
        """group key: bare for backbone1 (kept from the single-backbone engine), 'b2/...' for backbone2"""
        return name if self.tag == "b1" else f"{self.tag}/{name}"


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
        self.towers = [_Tower("b1", "backbone1.", model.backbone1)]
        if getattr(model, "backbone2", None) is not None:
            self.towers.append(_Tower("b2", "backbone2.", model.backbone2))
            if getattr(model, "fusion_heads", None) == -3:
                # vid_feat_lvls1 + vid_feat_lvls2 is a LIST concatenation and [-1] picks backbone2's output
                # (decoder_leave_focal.py:621-623): backbone1 never reaches the loss, so it is neither computed nor trained
                # here (its parameters stay at grad = None, exactly what autograd leaves behind in the reference)
                self.towers = self.towers[1:]
        self.fusion = getattr(model, "fusion_module", None)
        self.slots: list[_Slot] = []
        self.groups = {}
        self._layout()
        self.flat = None
        self.flat_grad = None
        self.flat_lp = None    # bf16 shadow of `flat` (bf16 mode)
        self.flat_lpT = None   # per-matrix transposed bf16 shadow (dgrad operand of the TC path)
        self.anchor = torch.zeros((), device=device, requires_grad=True)
        # dropout (dropout.py): probability of the next forward (0 = off), run seed, forward-call counter
        self.drop_p = 0.0
        self.drop_p_mlp = None     # None: the FFN's inner dropout follows drop_p (the reference's 0.1 / 0.1)
        self.drop_seed = int(torch.initial_seed())
        self._drop_call = 0
        self._drop_cur = None      # (thr8, scale, call) of the forward in flight, None = off
        self._ws = {}
        self._saved = None
        # data parallelism: ID-table gradients leave backward as (ids, rows) pairs instead of being added into the dense
        # table gradient; TrainStep exchanges them between the ranks and scatters the union (dp.exchange_rows)
        self.sparse_tables = False
        self.table_rows = []       # [(group key, tw, ids [B], rows [B, tw])] of the backward in flight
        self.use_tc = cfg.precision == "bf16" and bool(_lib.load().mmi_has_tc())
        # fp32 mode: GEMMs on the tensor cores as six bf16 x bf16 products of an exact three-term split of each fp32 operand
        # (hi*hi in one TMEM accumulator, the five small products in a second one, weight gradients in slices of 512 tokens
        # that meet through round-to-nearest atomics: 4.7e-7 / 3.2e-6 relative at K = 512 / 3072 against 4.1e-7 / 9.9e-7 for
        # the FFMA kernel, profiles/r02_s12_split_gemm_error.txt).  MMI_FP32_TC=0 keeps the FFMA kernels.  Attention,
        # LayerNorm and the loss stay fp32 SIMT.
        self.fp32_tc = cfg.precision != "bf16" and bool(_lib.load().mmi_has_tc()) and os.environ.get("MMI_FP32_TC", "1") != "0"
        # attention backward on the tensor-core path: "all" = one CTA per (b, h) owning every key (default, <= 640 keys),
        # "fused" = one kernel per key block with dQ reduced through an fp32 accumulator (measured slower than the pair it
        # replaces, kept for A/B runs), anything else = the dq + dk/dv kernel pair (also the fallback for long histories)
        mode = os.environ.get("MMI_ATTN_BWD", "all")
        self.allkeys_attn_bwd = mode == "all"
        self.fused_attn_bwd = mode == "fused"
        n_red = max(_lib.load().mmi_layernorm_bwd_workspace(cfg.d_model), _lib.load().mmi_head_bwd_workspace(cfg.d_model),
                    4 * cfg.d_model * max(cfg.max_usr_len, cfg.max_vid_len), 1 << 20)
        self.red_ws = torch.empty(int(n_red), device=device, dtype=torch.float32)
        self.scalars = torch.zeros(16, device=device, dtype=torch.float32)

    @property
    def mlp_ablation(self):
        return self.cfg.ablation in MLP_ABLATIONS

    def table_groups(self):
        """[(offset, numel)] of the embedding tables inside the flat buffers (towers with ID inputs)"""
        out = []
        for tw in self.towers:
            for s_, kind in (("vid", tw.vid_kind), ("usr", tw.usr_kind)):
                key = tw.k(f"{s_}_proj.w")
                if kind == "id" and key in self.groups:
                    out.append(self.groups[key])
        return out

    def layer_plan(self, i):
        """(query sides that are computed in layer i, {token side: [(block, j), ...] projections in buffer order})."""
        blocks = ATTN_BLOCKS[self.cfg.ablation]
        full = i < self.cfg.num_layers - 2          # the history side of layer N-2 never reaches the output
        qsides = tuple(s for s in ("vid", "usr") if blocks[s] and (s == "vid" or full))
        active = {n for s in qsides for n in blocks[s]}
        return qsides, {side: [(n, j) for n, j in PROJ_ORDER[side] if n in active] for side in ("vid", "usr")}

    @property
    def uses_history(self):
        """SelfAtt, SelfMLP and w/oAtt never read the history tokens (their usr_* parameters stay at grad = None)"""
        return self.cfg.ablation not in ("SelfAtt", "SelfMLP", "w/oAtt")

    # ------------------------------------------------------------------ parameter layout
    def _layout(self):
        cfg = self.cfg
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

        for tw in self.towers:
            bb, P, k = tw.m, tw.prefix, tw.k
            group(k("vid_proj.w"), [(P + "vid_proj.weight", bb.vid_proj.weight)])
            if tw.vid_kind == "id":
                group(k("frameid.w"), [(P + "frameid_proj.weight", bb.frameid_proj.weight)])
                group(k("frameid.b"), [(P + "frameid_proj.bias", bb.frameid_proj.bias)])
            else:
                group(k("vid_proj.b"), [(P + "vid_proj.bias", bb.vid_proj.bias)])
            token_sides = ("vid", "usr") if self.uses_history else ("vid",)   # SelfAtt never reads the history tokens
            if "usr" in token_sides:
                group(k("usr_proj.w"), [(P + "usr_proj.weight", bb.usr_proj.weight)])
                if tw.usr_kind == "image":
                    group(k("usr_proj.b"), [(P + "usr_proj.bias", bb.usr_proj.bias)])
            if cfg.use_pe:
                for s in token_sides:
                    group(k(f"{s}_pe"), [(f"{P}{s}_pe.weight", getattr(bb, s + "_pe").weight)])
            for s in token_sides:
                ln = getattr(bb, s + "_ln")
                group(k(f"{s}_ln.g"), [(f"{P}{s}_ln.weight", ln.weight)])
                group(k(f"{s}_ln.b"), [(f"{P}{s}_ln.bias", ln.bias)])
            if self.mlp_ablation and cfg.ablation != "w/oAtt":     # MLP_Block encoder (w/oAtt builds it and never calls it)
                for n, lin in enumerate(bb.encoder_mlp.linears()):
                    idx = bb.encoder_mlp.linear_indices()[n]
                    group(k(f"M{n}.w1"), [(f"{P}encoder_mlp.mlp.{idx}.weight", lin.weight)])
                    group(k(f"M{n}.b1"), [(f"{P}encoder_mlp.mlp.{idx}.bias", lin.bias)])
            for i in range(0 if self.mlp_ablation else N - 1):
                L = bb.encoder.layers[i]
                ca = L.cross_attn
                qsides, projs = self.layer_plan(i)
                q = f"{P}encoder.layers.{i}."
                # the projections in use of the candidate tokens / of the history tokens, adjacent (one fused GEMM each)
                for side in ("vid", "usr"):
                    if projs[side]:
                        group(k(f"L{i}.{side}.w6"), [(f"{q}cross_attn.{b}_proj.{j}.weight", getattr(ca, b + "_proj")[j].weight) for b, j in projs[side]])
                        group(k(f"L{i}.{side}.b6"), [(f"{q}cross_attn.{b}_proj.{j}.bias", getattr(ca, b + "_proj")[j].bias) for b, j in projs[side]])
                for side in qsides:
                    ff, ln = getattr(ca, "ff_" + side), getattr(ca, "ln_" + side)
                    group(k(f"L{i}.{side}.wo"), [(f"{q}cross_attn.ff_{side}.weight", ff.weight)])
                    group(k(f"L{i}.{side}.bo"), [(f"{q}cross_attn.ff_{side}.bias", ff.bias)])
                    group(k(f"L{i}.{side}.ln1.g"), [(f"{q}cross_attn.ln_{side}.weight", ln.weight)])
                    group(k(f"L{i}.{side}.ln1.b"), [(f"{q}cross_attn.ln_{side}.bias", ln.bias)])
                    mlp, ln2 = getattr(L, "ff_" + side), getattr(L, "ln_" + side)
                    group(k(f"L{i}.{side}.w1"), [(f"{q}ff_{side}.layers.0.weight", mlp.layers[0].weight)])
                    group(k(f"L{i}.{side}.b1"), [(f"{q}ff_{side}.layers.0.bias", mlp.layers[0].bias)])
                    group(k(f"L{i}.{side}.w2"), [(f"{q}ff_{side}.layers.1.weight", mlp.layers[1].weight)])
                    group(k(f"L{i}.{side}.b2"), [(f"{q}ff_{side}.layers.1.bias", mlp.layers[1].bias)])
                    group(k(f"L{i}.{side}.ln2.g"), [(f"{q}ln_{side}.weight", ln2.weight)])
                    group(k(f"L{i}.{side}.ln2.b"), [(f"{q}ln_{side}.bias", ln2.bias)])
        if self.fusion is not None:    # InteractionAggregation (models/decoder_leave_focal.py:392-423)
            fm = self.fusion
            group("fusion.wxy", [("fusion_module.w_xy", fm.w_xy)])
            group("fusion.wx", [("fusion_module.w_x.weight", fm.w_x.weight)])
            group("fusion.bx", [("fusion_module.w_x.bias", fm.w_x.bias)])
            group("fusion.wy", [("fusion_module.w_y.weight", fm.w_y.weight)])
            group("fusion.by", [("fusion_module.w_y.bias", fm.w_y.bias)])
        else:
            group("head.w", [("stage_mlp1.weight", self.model.stage_mlp1.weight)])
            group("head.b", [("stage_mlp1.bias", self.model.stage_mlp1.bias)])
            if getattr(self.model, "stage_mlp2", None) is not None:      # fusion_heads == 0: a second Linear head
                group("head2.w", [("stage_mlp2.weight", self.model.stage_mlp2.weight)])
                group("head2.b", [("stage_mlp2.bias", self.model.stage_mlp2.bias)])
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
            if self.fusion is not None:     # W_h [dx, dy] -> W_h^T [dy, dx] per head (B operand of X_h W_h)
                off, n = self.groups["fusion.wxy"]
                H = self.fusion.num_heads
                per = n // H
                dxh = d // H
                for h in range(H):
                    ops.cast_bf16(self.flat[off + h * per:off + (h + 1) * per], self.flat_lpT[off + h * per:off + (h + 1) * per], dxh, dxh,
                                  transpose=True)

    # ------------------------------------------------------------------ workspace
    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(shape, device=self.device, dtype=dtype)
            self._ws[key] = t
        return t

    def _zbuf(self, name, shape, dtype):
        """workspace tensor that is zero when first handed out (its users leave it zero)"""
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.zeros(shape, device=self.device, dtype=dtype)
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

    # ------------------------------------------------------------------ dropout sites
    DROP_EMB, DROP_ATTN, DROP_ATTN_OUT, DROP_MLP1, DROP_MLP2, DROP_BLOCK = range(6)   # DROP_BLOCK: hidden units of MLP_Block

    def _site(self, tw, layer, side, kind):
        """DropSite of one nn.Dropout call of the reference in the forward in flight (None when dropout is off):
        kind DROP_EMB encoder.py:472/386, DROP_ATTN :145-150, DROP_ATTN_OUT :163-164, DROP_MLP1 mlp.py:21-22,
        DROP_MLP2 encoder.py:198-202.  The key depends on (seed, forward-call counter, tower, layer, side, kind)."""
        if self._drop_cur is None:
            return None
        thr8, scale, call = self._drop_cur
        if kind == self.DROP_MLP1 and self.drop_p_mlp is not None:
            thr8, scale = quantise(self.drop_p_mlp)
            if not thr8:
                return None
        sid = (((0 if tw.tag == "b1" else 1) * 64 + layer) * 2 + (0 if side == "vid" else 1)) * 8 + kind
        return DropSite(site_key(self.drop_seed, call, sid), thr8, scale)

    # ------------------------------------------------------------------ building blocks
    def _linear(self, x, M, K, wkey, bkey, N, out, *, act=ACT_NONE, preact=None, add=None, add_mod=0, ld_add=0,
                n_prefix=None, drop=None):
        """out[M,N] = act(x[M,K] W[N,K]^T + b) (+ add).  On the tensor-core path `preact` receives
        gelu'(z) instead of z (the backward epilogue then needs no erf; see mmi_gemm save_act_grad)."""
        lp = self.cfg.precision == "bf16"
        W = self.w(wkey, lp)
        ops.gemm(GEMM_NT, self._impl(M, N, K, GEMM_NT), x, K, W, K, out, N, M, N, K, bias=self.w(bkey), act=act,
                 preact=preact, add=add, add_mod=add_mod, ld_add=ld_add, save_act_grad=self.use_tc and preact is not None,
                 drop=drop)

    def _linear_bwd(self, dy, x, M, N, K, wkey, bkey, dx, *, mul_gelu_grad=None, add=None, need_dx=True, bias_done=False,
                    drop=None, relu_out=None, relu_scale=1.0):
        """dW += dy^T x ; db += colsum(dy) ; dx = dy W (* gelu'(mul)) (+ add).  bias_done: the kernel that produced dy
        (LayerNorm backward) already accumulated the bias gradient.  drop: dropout of the site that FOLLOWED the GELU whose
        derivative is multiplied in (strict-parity path only: the tensor-core path folds it into the saved gelu')."""
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
            # relu_out: x itself is dropout(relu(.)) of the previous MLP_Block unit -- dx is multiplied by its derivative
            # (x > 0 ? relu_scale : 0) in the epilogue, which makes dx the gradient w.r.t. that unit's pre-activation
            mul, mig = (relu_out, 2) if relu_out is not None else (mul_gelu_grad, 1 if self.use_tc else 0)
            if self._impl(M, K, N, GEMM_NT) == IMPL_TC:
                ops.gemm(GEMM_NT, IMPL_TC, dy, N, self.wT(wkey), N, dx, K, M, K, N, mul_gelu_grad=mul, add=add,
                         add_mod=M if add is not None else 0, ld_add=K, mul_is_grad=mig, drop=drop, mul_scale=relu_scale)
            else:
                ops.gemm(GEMM_NN, IMPL_SIMT, dy, N, self.w(wkey, lp), K, dx, K, M, K, N, mul_gelu_grad=mul,
                         add=add, add_mod=M if add is not None else 0, ld_add=K, mul_is_grad=mig, drop=drop, mul_scale=relu_scale)

    # ------------------------------------------------------------------ forward
    def forward(self, usr_image, usr_mask, vid_image, vid_mask, usr_id=None, vid_id=None, refresh=True, need_bwd=True):
        """usr_image [B,Lt,Din] / vid_image [B,Lv,Din] already L1-normalised (the driver does it,
        main...SegMM.py:272-273; our own data path fuses it into the gather); usr_id / vid_id int64 [B] for towers with
        ID inputs.  Returns fp32 logits [B, Lv] before the position bias (a workspace tensor: clone before the next call)."""
        cfg = self.cfg
        ops.FP32_TC["on"] = self.fp32_tc
        self.ensure_bound()
        if refresh:                      # bf16 weight shadows; a micro-batched step refreshes them once, not per slice
            self.refresh_low_precision()
        d = cfg.d_model
        self._need_bwd = need_bwd        # forward-only calls (mode="inference") skip the saved GELU' tensor
        if self.drop_p > 0.0:            # a fresh set of masks per forward call: the counter goes into every site key
            thr8, scale = quantise(self.drop_p)
            self._drop_call += 1
            self._drop_cur = (thr8, scale, self._drop_call) if thr8 else None
        else:
            self._drop_cur = None
        sv = {"towers": {}}
        outs = []
        B = Lv = None
        for tw in self.towers:
            x_out, B, Lv = self._tower_forward(tw, sv, usr_image, usr_mask, vid_image, vid_mask, usr_id, vid_id)
            outs.append(x_out)
        sv["B"], sv["Lv"] = B, Lv
        R = B * Lv
        logits = self._buf("logits", (B, Lv), torch.float32)
        if self.fusion is None and len(outs) == 1:
            ops.head_fwd(outs[0], R, d, self.w("head.w"), self.w("head.b"), logits)
        elif self.fusion is None:
            # two backbones without InteractionAggregation (decoder_leave_focal.py:624-631): two chained Linear(d, 1) passes
            (w1, b1), (w2, b2) = self._two_head_params()
            part = self._buf("head.part", (R,), torch.float32)
            ops.head_fwd(outs[0], R, d, w1, b1, part)
            ops.head_fwd(outs[1], R, d, w2, b2, logits, add=part)
        else:
            # InteractionAggregation: w_x.x + w_y.y + sum_h x_h^T W_h y_h  (decoder_leave_focal.py:411-423)
            T = self.act_dtype
            H = self.fusion.num_heads
            dxh = d // H
            x, y = outs
            t = self._buf("fusion.t", (R, d), T)
            esz = x.element_size()
            lp = cfg.precision == "bf16"
            off, n = self.groups["fusion.wxy"]
            per = n // H
            for h in range(H):
                xa = _View(x.data_ptr() + h * dxh * esz, x)
                tc = _View(t.data_ptr() + h * dxh * esz, t)
                if self._impl(R, dxh, dxh, GEMM_NT) == IMPL_TC:
                    ops.gemm(GEMM_NT, IMPL_TC, xa, d, self.flat_lpT[off + h * per:off + (h + 1) * per], dxh, tc, d, R, dxh, dxh)
                else:
                    wsrc = (self.flat_lp if lp else self.flat)[off + h * per:off + (h + 1) * per]
                    ops.gemm(GEMM_NN, IMPL_SIMT, xa, d, wsrc, dxh, tc, d, R, dxh, dxh)
            lin_x = self._buf("fusion.linx", (R,), torch.float32)
            lin_y = self._buf("fusion.liny", (R,), torch.float32)
            ops.head_fwd(x, R, d, self.w("fusion.wx"), self.w("fusion.bx"), lin_x)
            ops.head_fwd(y, R, d, self.w("fusion.wy"), self.w("fusion.by"), lin_y)
            ops.rowdot_fwd(t, d, y, d, R, d, logits, add1=lin_x, add2=lin_y)
            sv["fusion.t"] = t
        sv["x_out"] = outs
        self._saved = sv
        return logits

    def _two_head_params(self, grad=False):
        """((w1, b1), (w2, b2)) of the two chained Linear(d, 1) passes for fusion_heads 0 / -1 / -2; b2 is None when the
        bias belongs to the first pass (grad=True: the matching views of the flat gradient buffer)."""
        d = self.cfg.d_model
        buf = self.g if grad else self.w
        fh = self.model.fusion_heads
        if fh == 0:        # stage_mlp1(x1) + stage_mlp2(x2)
            return (buf("head.w"), buf("head.b")), (buf("head2.w"), buf("head2.b"))
        if fh == -1:       # Linear(2d, 1)(cat([x1, x2]))
            return (buf("head.w")[:d], buf("head.b")), (buf("head.w")[d:], None)
        return (buf("head.w"), buf("head.b")), (buf("head.w"), None)     # -2: Linear(d, 1)(x1 + x2)

    def _tower_forward(self, tw, sv, usr_image, usr_mask, vid_image, vid_mask, usr_id, vid_id):
        cfg, k, bb = self.cfg, tw.k, tw.m
        T = self.act_dtype
        d, H, N = cfg.d_model, cfg.nhead, cfg.num_layers
        B = (vid_id if tw.vid_kind == "id" else vid_image).shape[0]
        Lv = bb.max_vid_len if tw.vid_kind == "id" else vid_image.shape[1]
        Lt = 1 if tw.usr_kind == "id" else usr_image.shape[1]
        if Lt > bb.max_usr_len or Lv > bb.max_vid_len:
            raise ValueError(f"sequence longer than the position table: Lt={Lt} (max {bb.max_usr_len}), Lv={Lv} (max {bb.max_vid_len})")
        ops._need_cuda(vid_mask, usr_mask if tw.usr_kind == "image" else None)
        Ls = {"usr": Lt, "vid": Lv}
        Ts = {"usr": B * Lt, "vid": B * Lv}
        mask = {"vid": vid_mask.to(torch.bool).contiguous().view(torch.uint8)}
        token_sides = ("vid", "usr") if self.uses_history else ("vid",)
        if "usr" not in token_sides:
            pass
        elif tw.usr_kind == "id":       # one token per user, mask of ones (encoder.py:478-481)
            mask["usr"] = self._buf(k("ones_mask"), (B, 1), torch.uint8)
            mask["usr"].fill_(1)
        else:
            mask["usr"] = usr_mask.to(torch.bool).contiguous().view(torch.uint8)
        ts = {"B": B, "Lt": Lt, "Lv": Lv, "mask": mask, "layers": [], "x_in": {}, "ids": {}}
        X = {}
        for s in token_sides:
            kind = tw.vid_kind if s == "vid" else tw.usr_kind
            e = self._buf(k(f"emb_pre.{s}"), (Ts[s], d), T)
            st = self._buf(k(f"emb_st.{s}"), (Ts[s], 2), torch.float32)
            x0 = self._buf(k(f"x0.{s}"), (Ts[s], d), T)
            pe = self.w(k(f"{s}_pe")) if cfg.use_pe else None
            if kind == "image":
                img = vid_image if s == "vid" else usr_image
                ops._need_cuda(img)
                xin = img.to(T).contiguous().view(Ts[s], -1)
                ts["x_in"][s] = xin
                self._linear(xin, Ts[s], xin.shape[1], k(f"{s}_proj.w"), k(f"{s}_proj.b"), d, e, add=pe,
                             add_mod=Ls[s] if pe is not None else 0, ld_add=d)
            else:
                ids = (vid_id if s == "vid" else usr_id)
                if ids is None:
                    raise ValueError(f"{tw.prefix[:-1]} takes {s} ids but none were passed")
                ids = ids.to(torch.int64).contiguous()
                ops._need_cuda(ids)
                ts["ids"][s] = ids
                table = self.w(k(f"{s}_proj.w"))
                tw_cols = d // 2 if s == "vid" else d
                table = table.view(-1, tw_cols)
                fpos = None
                if s == "vid" and cfg.no_pos:
                    # 'noPos' (encoder.py:428-429): one torch.randperm(Lv) per interaction, drawn on the host from torch's
                    # default generator exactly like the reference, fed to the frame projection instead of 0..Lv-1
                    fpos = torch.stack([torch.randperm(Ls[s]) for _ in range(B)]).float().to(self.device).contiguous()
                if s == "vid":
                    ts["frame_pos"] = fpos
                ops.id_embed_fwd(table, ids, B, Ls[s], d, e, frame_w=self.w(k("frameid.w")) if s == "vid" else None,
                                 frame_b=self.w(k("frameid.b")) if s == "vid" else None, pe=pe, frame_pos=fpos)
            ts[f"drop.emb.{s}"] = self._site(tw, 63, s, self.DROP_EMB)
            ops.layernorm_fwd(e, Ts[s], d, self.w(k(f"{s}_ln.g")), self.w(k(f"{s}_ln.b")), x0, st, drop=ts[f"drop.emb.{s}"])
            ts[f"emb_pre.{s}"], ts[f"emb_st.{s}"] = e, st
            X[s] = x0
        esz = x0.element_size()
        if self.mlp_ablation:
            sv["towers"][tw.tag] = ts
            return self._mlp_encoder_forward(tw, ts, X, B, Lt, Lv), B, Lv
        for i in range(N - 1):
            sides, projs = self.layer_plan(i)
            nq = {s: len(projs[s]) for s in ("vid", "usr")}
            cols = {s: {pj: c for c, pj in enumerate(projs[s])} for s in ("vid", "usr")}
            lay = {"sides": sides, "nq": nq, "cols": cols, "x": dict(X)}
            qkv = {}
            for s in token_sides:
                if nq[s]:
                    qkv[s] = self._buf(k(f"qkv.{i}.{s}"), (Ts[s], nq[s] * d), T)
                    self._linear(X[s], Ts[s], d, k(f"L{i}.{s}.w6"), k(f"L{i}.{s}.b6"), nq[s] * d, qkv[s])
            lay["qkv"] = qkv

            def col(n, j):
                """(pointer, leading dimension) of projection j of block n inside the fused projection buffer of its side"""
                s_ = Q_SIDE[n] if j == 0 else KEY_SIDE[n]
                return (qkv[s_].data_ptr() + cols[s_][(n, j)] * d * esz, nq[s_] * d)

            attn = {}
            for s in sides:
                a_out = self._buf(k(f"attn.{i}.{s}"), (Ts[s], d), T)
                lse = self._buf(k(f"lse.{i}.{s}"), (B, H, Ls[s]), torch.float32)
                blocks = [dict(q=col(n, 0), k=col(n, 1), v=col(n, 2), mask_k=mask[KEY_SIDE[n]], Lk=Ls[KEY_SIDE[n]])
                          for n in ATTN_BLOCKS[cfg.ablation][s]]
                attn_impl = IMPL_TC if (self.use_tc and d // H == 32 and d % 8 == 0) else IMPL_SIMT
                side = ops.AttnSide(ops.dt(a_out), attn_impl, B, H, d // H, Ls[s], mask[s], a_out, d, lse, blocks,
                                    drop=self._site(tw, i, s, self.DROP_ATTN))
                side.fwd()
                attn[s] = (side, a_out, lse)
            lay["attn"] = attn
            for s in sides:
                p1 = self._buf(k(f"p1.{i}.{s}"), (Ts[s], d), T)
                st1 = self._buf(k(f"st1.{i}.{s}"), (Ts[s], 2), torch.float32)
                x1 = self._buf(k(f"x1.{i}.{s}"), (Ts[s], d), T)
                z1 = self._buf(k(f"z1.{i}.{s}"), (Ts[s], d), T)
                g1 = self._buf(k(f"g1.{i}.{s}"), (Ts[s], d), T)
                p2 = self._buf(k(f"p2.{i}.{s}"), (Ts[s], d), T)
                st2 = self._buf(k(f"st2.{i}.{s}"), (Ts[s], 2), torch.float32)
                x2 = self._buf(k(f"x2.{i}.{s}"), (Ts[s], d), T)
                dr = {kind: self._site(tw, i, s, kind) for kind in (self.DROP_ATTN_OUT, self.DROP_MLP1, self.DROP_MLP2)}
                self._linear(attn[s][1], Ts[s], d, k(f"L{i}.{s}.wo"), k(f"L{i}.{s}.bo"), d, p1, add=X[s], add_mod=Ts[s], ld_add=d,
                             drop=dr[self.DROP_ATTN_OUT])
                ops.layernorm_fwd(p1, Ts[s], d, self.w(k(f"L{i}.{s}.ln1.g")), self.w(k(f"L{i}.{s}.ln1.b")), x1, st1)
                self._linear(x1, Ts[s], d, k(f"L{i}.{s}.w1"), k(f"L{i}.{s}.b1"), d, g1, act=ACT_GELU, preact=z1 if self._need_bwd else None,
                             drop=dr[self.DROP_MLP1])
                self._linear(g1, Ts[s], d, k(f"L{i}.{s}.w2"), k(f"L{i}.{s}.b2"), d, p2, add=x1, add_mod=Ts[s], ld_add=d,
                             drop=dr[self.DROP_MLP2])
                ops.layernorm_fwd(p2, Ts[s], d, self.w(k(f"L{i}.{s}.ln2.g")), self.w(k(f"L{i}.{s}.ln2.b")), x2, st2)
                lay[s] = dict(p1=p1, st1=st1, x1=x1, z1=z1, g1=g1, p2=p2, st2=st2, drop=dr)
                X[s] = x2
            ts["layers"].append(lay)
        sv["towers"][tw.tag] = ts
        return X["vid"], B, Lv

    # ------------------------------------------------------------------ MLP ablations (SelfMLP / CrossMLP / w/oAtt)
    def _mlp_encoder_forward(self, tw, ts, X, B, Lt, Lv):
        """models/encoder.py:503-511: the attention encoder is replaced by MLP_Block ([Linear -> ReLU -> Dropout] x hidden,
        Linear) over the candidate tokens (SelfMLP), over [history ; candidate] tokens followed by AdaptiveAvgPool1d(40)
        along the token axis (CrossMLP), or by nothing at all (w/oAtt: the embedded candidate tokens go to the head)."""
        cfg, k = self.cfg, tw.k
        T, d = self.act_dtype, cfg.d_model
        if cfg.ablation == "w/oAtt":
            ts["mlp"] = None
            return X["vid"]
        if cfg.ablation == "CrossMLP":
            L = Lt + Lv
            xcat = self._buf(k("mlp.xcat"), (B, L, d), T)           # torch.cat((usr_feat, vid_feat), dim=-2): a strided copy
            xcat[:, :Lt].copy_(X["usr"].view(B, Lt, d))
            xcat[:, Lt:].copy_(X["vid"].view(B, Lv, d))
            h, rows = xcat.view(B * L, d), B * L
        else:
            L = Lv
            h, rows = X["vid"], B * Lv
        nlin = len(tw.m.encoder_mlp.linears())
        units = []
        for n in range(nlin):
            last = n == nlin - 1
            site = None if last else self._site(tw, n, "vid", self.DROP_BLOCK)
            y = self._buf(k(f"mlp.y{n}"), (rows, d), T)
            self._linear(h, rows, d, k(f"M{n}.w1"), k(f"M{n}.b1"), d, y, act=ACT_NONE if last else ACT_RELU, drop=site)
            units.append(dict(x=h, y=y, scale=site.scale if site is not None else 1.0))
            h = y
        ts["mlp"] = dict(units=units, rows=rows, L=L)
        if cfg.ablation == "CrossMLP":
            out = self._buf(k("mlp.pool"), (B * Lv, d), T)
            ops.adaptive_pool_fwd(h, B, L, d, Lv, out)
            return out
        return h

    def _mlp_encoder_backward(self, tw, ts, dx_out):
        """gradient w.r.t. the embedded tokens {side: tensor} from the gradient of the encoder output"""
        cfg, k = self.cfg, tw.k
        T, d = self.act_dtype, cfg.d_model
        B, Lt, Lv = ts["B"], ts["Lt"], ts["Lv"]
        m = ts["mlp"]
        if m is None:
            return {"vid": dx_out, "usr": None}
        rows, L, units = m["rows"], m["L"], m["units"]
        g = dx_out
        if cfg.ablation == "CrossMLP":
            g = self._buf(k("bw.mlp.dpool"), (rows, d), T)
            ops.adaptive_pool_bwd(dx_out, B, L, d, Lv, g)
        for n in reversed(range(len(units))):
            u = units[n]
            dx = self._buf(k(f"bw.mlp.dx{n & 1}"), (rows, d), T)
            prev = units[n - 1] if n > 0 else None      # the unit whose dropout(relu(.)) output is this Linear's input
            self._linear_bwd(g, u["x"], rows, d, d, k(f"M{n}.w1"), k(f"M{n}.b1"), dx,
                             relu_out=prev["y"] if prev is not None else None, relu_scale=prev["scale"] if prev is not None else 1.0)
            g = dx
        if cfg.ablation == "CrossMLP":
            gv = g.view(B, L, d)
            du = self._buf(k("bw.mlp.du"), (B * Lt, d), T)
            dv = self._buf(k("bw.mlp.dv"), (B * Lv, d), T)
            du.view(B, Lt, d).copy_(gv[:, :Lt])
            dv.view(B, Lv, d).copy_(gv[:, Lt:])
            return {"vid": dv, "usr": du}
        return {"vid": g, "usr": None}

    # ------------------------------------------------------------------ loss
    def loss(self, logits, gt, exposure_prob, inv_bsz, loss_cfg=None, bpr_scale=1.0, need_grad=True):
        """Fused loss (focal and / or interestBPR) + learnable position bias + diagnostics + dlogits
        (models/decoder_leave_focal.py:490-572).  `gt` is rewritten in place like the reference (:534-535) when
        focal is on.  Returns (scalars [focal, mse, mse2, loss, interestBPR, huber, hazard, surviveCE, interestCE,
        interestKL, ...], logits incl. bias): device tensors, no sync.  loss_cfg: dict(use_focal, w_focal, use_bpr, w_bpr
        [, others={name: weight}, mask_loss, ce_after_focal, kl_after_focal]); None = focal with weight 1."""
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
                         others=cfg.get("others"), mask_loss=cfg.get("mask_loss", 0), ce_after_focal=cfg.get("ce_after_focal", False),
                         kl_after_focal=cfg.get("kl_after_focal", False),
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
        gradient of the loss (device scalar) -- never read on the host.  on_ready(lo): every gradient at flat
        offsets >= lo is final (the layout is forward order, backward walks it from the top)."""
        sv = self._saved
        if sv is None or "dlogits" not in sv:
            raise RuntimeError("backward() called before forward()+loss()")
        ops.FP32_TC["on"] = self.fp32_tc
        self.bind_grads()
        cfg = self.cfg
        T = self.act_dtype
        d = cfg.d_model
        B, Lv = sv["B"], sv["Lv"]
        R = B * Lv
        if "bias_weight" in self.groups:   # d loss / d (bias_weight, bias_bias) from the loss kernel, times the upstream gradient
            db = self._buf("dbias", (2, self.groups["bias_weight"][1]), torch.float32)
            gs = gscale if gscale is not None else 1.0
            self.g("bias_weight").add_(db[0] * gs)
            self.g("bias_bias").add_(db[1] * gs)
        dX_out = []
        if self.fusion is None and len(sv["x_out"]) == 1:
            dx = self._buf("bw.dx.b1.vid", (R, d), T)
            ops.head_bwd(sv["x_out"][0], R, d, self.w("head.w"), sv["dlogits"], gscale, dx, self.g("head.w"), self.g("head.b"), self.red_ws)
            dX_out.append(dx)
            if on_ready is not None:
                on_ready(self.groups["head.w"][0])
        elif self.fusion is None:
            (w1, _), (w2, _) = self._two_head_params()
            (g1, gb1), (g2, gb2) = self._two_head_params(grad=True)
            dx1, dx2 = self._buf("bw.dx.b1.vid", (R, d), T), self._buf("bw.dx.b2.vid", (R, d), T)
            ops.head_bwd(sv["x_out"][0], R, d, w1, sv["dlogits"], gscale, dx1, g1, gb1, self.red_ws)
            ops.head_bwd(sv["x_out"][1], R, d, w2, sv["dlogits"], gscale, dx2, g2, gb2, self.red_ws)
            dX_out = [dx1, dx2]
            if on_ready is not None:
                on_ready(self.groups["head.w"][0])
        else:
            H = self.fusion.num_heads
            dxh = d // H
            x, y = sv["x_out"]
            t = sv["fusion.t"]
            esz = x.element_size()
            lp = cfg.precision == "bf16"
            dx_lin = self._buf("bw.fusion.dxlin", (R, d), T)
            dy_lin = self._buf("bw.fusion.dylin", (R, d), T)
            ops.head_bwd(x, R, d, self.w("fusion.wx"), sv["dlogits"], gscale, dx_lin, self.g("fusion.wx"), self.g("fusion.bx"), self.red_ws)
            ops.head_bwd(y, R, d, self.w("fusion.wy"), sv["dlogits"], gscale, dy_lin, self.g("fusion.wy"), self.g("fusion.by"), self.red_ws)
            dt_ = self._buf("bw.fusion.dt", (R, d), T)
            dy = self._buf("bw.dx.b2.vid", (R, d), T)
            ops.rowdot_bwd(sv["dlogits"], gscale, t, d, y, d, R, d, dt_, dy, dy_add=dy_lin)
            dx = self._buf("bw.dx.b1.vid", (R, d), T)
            off, n = self.groups["fusion.wxy"]
            per = n // H
            for h in range(H):
                xa = _View(x.data_ptr() + h * dxh * esz, x)
                dta = _View(dt_.data_ptr() + h * dxh * esz, dt_)
                dxa = _View(dx.data_ptr() + h * dxh * esz, dx)
                adda = _View(dx_lin.data_ptr() + h * dxh * esz, dx_lin)
                gw = self.flat_grad[off + h * per:off + (h + 1) * per]
                # dW_h += X_h^T dT_h ; dX_h = dT_h W_h^T + g w_x
                impl_w = self._impl(dxh, dxh, R, GEMM_TN)
                ops.gemm(GEMM_TN, impl_w, xa, d, dta, d, gw, dxh, dxh, dxh, R, accumulate=True, split_k=0 if impl_w == IMPL_TC else 1,
                         out_dtype=_lib.F32)
                wsrc = (self.flat_lp if lp else self.flat)[off + h * per:off + (h + 1) * per]
                ops.gemm(GEMM_NT, self._impl(R, dxh, dxh, GEMM_NT), dta, d, wsrc, dxh, dxa, d, R, dxh, dxh, add=adda, add_mod=R, ld_add=d)
            dX_out = [dx, dy]
            if on_ready is not None:
                on_ready(self.groups["fusion.wxy"][0])
        for tw, dxo in reversed(list(zip(self.towers, dX_out))):
            self._tower_backward(tw, sv["towers"][tw.tag], dxo, on_ready)
        self._saved = None

    def _tower_backward(self, tw, ts, dx_out, on_ready):
        cfg, k = self.cfg, tw.k
        T = self.act_dtype
        d, H, N = cfg.d_model, cfg.nhead, cfg.num_layers
        B, Lt, Lv = ts["B"], ts["Lt"], ts["Lv"]
        Ls = {"usr": Lt, "vid": Lv}
        Ts = {"usr": B * Lt, "vid": B * Lv}

        def scratch(name, s, width=d):
            return self._buf(f"bw.{name}.{tw.tag}.{s}.{width}", (Ts[s], width), T)

        dX = {"vid": dx_out, "usr": None}
        token_sides = ("vid", "usr") if self.uses_history else ("vid",)
        if self.mlp_ablation:
            dX = self._mlp_encoder_backward(tw, ts, dx_out)
            if on_ready is not None and cfg.ablation != "w/oAtt":
                on_ready(self.groups[k("M0.w1")][0])
        for i in reversed(range(0 if self.mlp_ablation else N - 1)):
            lay = ts["layers"][i]
            sides, nq, cols, Xin = lay["sides"], lay["nq"], lay["cols"], lay["x"]
            dP1, dA = {}, {}
            for s in sides:
                a = lay[s]
                pre = k(f"L{i}.{s}.")
                dr = a["drop"]
                # with dropout the LayerNorm backward writes two tensors: dp (residual branch) and dpm = dropout'(dp), the
                # gradient into the Linear whose dropped-out output was added to the residual (its bias gradient sums dpm)
                dp2 = scratch("dp2", s)
                dp2m = scratch("dp2m", s) if dr[self.DROP_MLP2] is not None else dp2
                ops.layernorm_bwd(dX[s], a["p2"], Ts[s], d, self.w(pre + "ln2.g"), a["st2"], None, dp2, self.g(pre + "ln2.g"),
                                  self.g(pre + "ln2.b"), self.red_ws, dxsum=self.g(pre + "b2"), dx_drop=dr[self.DROP_MLP2],
                                  dx_dropped=dp2m)
                dz1 = scratch("dz1", s)
                self._linear_bwd(dp2m, a["g1"], Ts[s], d, d, pre + "w2", pre + "b2", dz1, mul_gelu_grad=a["z1"], bias_done=True,
                                 drop=None if self.use_tc else dr[self.DROP_MLP1])
                dx1 = scratch("dx1", s)
                self._linear_bwd(dz1, a["x1"], Ts[s], d, d, pre + "w1", pre + "b1", dx1, add=dp2)
                dp1 = scratch("dp1", s)
                dp1m = scratch("dp1m", s) if dr[self.DROP_ATTN_OUT] is not None else dp1
                ops.layernorm_bwd(dx1, a["p1"], Ts[s], d, self.w(pre + "ln1.g"), a["st1"], None, dp1, self.g(pre + "ln1.g"),
                                  self.g(pre + "ln1.b"), self.red_ws, dxsum=self.g(pre + "bo"), dx_drop=dr[self.DROP_ATTN_OUT],
                                  dx_dropped=dp1m)
                da = scratch("da", s)
                self._linear_bwd(dp1m, lay["attn"][s][1], Ts[s], d, d, pre + "wo", pre + "bo", da, bias_done=True)
                dP1[s], dA[s] = dp1, da
            dqkv = {s: scratch("dqkv", s, nq[s] * d) for s in token_sides if nq[s]}
            esz = dqkv["vid"].element_size()
            # tensor-core attention: the backward kernels add the fp32 column sums of dq / dk / dv (= the bias gradients of
            # the fused projections) straight into the flat gradient buffer, so dqkv is not re-read by a colsum pass
            fused_bias = all(lay["attn"][s][0].a.impl == IMPL_TC for s in sides)
            gb6 = {s: self.g(k(f"L{i}.{s}.b6")) for s in dqkv}

            def gset(n):
                out = {}
                for j, (kp, kb) in enumerate((("dq", "dbq"), ("dk", "dbk"), ("dv", "dbv"))):
                    s_ = Q_SIDE[n] if j == 0 else KEY_SIDE[n]
                    c = cols[s_][(n, j)]
                    out[kp] = (dqkv[s_].data_ptr() + c * d * esz, nq[s_] * d)
                    out[kb] = gb6[s_].data_ptr() + 4 * c * d if fused_bias else None
                return out

            for s in sides:
                side = lay["attn"][s][0]
                delta = self._buf(f"delta.{tw.tag}.{s}", (B, H, Ls[s]), torch.float32)
                names = ATTN_BLOCKS[cfg.ablation][s]
                side.set_bwd(dA[s], d, delta, [gset(n) for n in names])
                # one launch: dq, dk, dv of both key blocks by persistent CTAs that own every key of a (b, h) item (<= 640 keys).
                # Measured at the c2 shapes with dropout and bias sums live (tools/attn_bench.py, profiles/r02_s8_*): history-side
                # queries 5.86 ms against 6.99 ms for dq + 2 x dk/dv, candidate-side queries 1.28 against 1.65 ms
                if side.a.impl == IMPL_TC and self.allkeys_attn_bwd and side.bwd_all():
                    continue
                if side.a.impl == IMPL_TC and self.fused_attn_bwd:
                    # one kernel per key block: dK, dV and that block's dQ (partial tiles reduced through an fp32 accumulator
                    # that the kernel leaves zeroed, so one pair of buffers serves every layer of the tower)
                    acc = [self._zbuf(f"dqacc.{tw.tag}.{s}.{w_}", (Ts[s], d), torch.float32) for w_ in range(len(names))]
                    cnt = [self._zbuf(f"dqcnt.{tw.tag}.{s}.{w_}", (B * H,), torch.int32) for w_ in range(len(names))]
                    side.set_fused(acc, cnt)
                    done = [side.bwd_fused(w_) for w_ in range(len(names))]
                    assert all(done) or not any(done)
                    if all(done):
                        continue
                side.bwd_dq()
                for w_ in range(len(names)):
                    side.bwd_dkv(w_)
            new_dX = {}
            for s in token_sides:
                pre = k(f"L{i}.{s}.")
                if not nq[s]:          # no projection of these tokens is live in this layer: the gradient passes through
                    new_dX[s] = dP1.get(s, dX.get(s))
                    continue
                out = scratch("dxl", s)
                # dX[s] (the grad w.r.t. this layer's OUTPUT) is dead by now: dp2 consumed it
                resid = dP1.get(s)
                if resid is None and dX.get(s) is not None and s not in sides:
                    resid = dX[s]      # tokens that are only keys here but were updated by a LATER layer: add that path
                self._linear_bwd(dqkv[s], Xin[s], Ts[s], nq[s] * d, d, pre + "w6", pre + "b6", out, add=resid, bias_done=fused_bias)
                new_dX[s] = out
            dX = new_dX
            if on_ready is not None:
                on_ready(self.groups[k(f"L{i}.vid.w6")][0])  # first group of layer i in the flat layout
        for s in token_sides:
            if dX.get(s) is None:
                continue
            kind = tw.vid_kind if s == "vid" else tw.usr_kind
            de = self._buf(f"bw.de.{tw.tag}.{s}", (Ts[s], d), T)
            if kind == "image":
                ops.layernorm_bwd(dX[s], ts[f"emb_pre.{s}"], Ts[s], d, self.w(k(f"{s}_ln.g")), ts[f"emb_st.{s}"], None, de,
                                  self.g(k(f"{s}_ln.g")), self.g(k(f"{s}_ln.b")), self.red_ws, dxsum=self.g(k(f"{s}_proj.b")),
                                  dy_drop=ts[f"drop.emb.{s}"])
            else:
                # one token per row (the ID tower's user token): d pe[0, :] is the plain column sum of dE, taken by the
                # LayerNorm backward from its fp32 values (a separate pass over the bf16 dE loses digits to cancellation)
                pe_fused = cfg.use_pe and Ls[s] == 1
                ops.layernorm_bwd(dX[s], ts[f"emb_pre.{s}"], Ts[s], d, self.w(k(f"{s}_ln.g")), ts[f"emb_st.{s}"], None, de,
                                  self.g(k(f"{s}_ln.g")), self.g(k(f"{s}_ln.b")), self.red_ws, dy_drop=ts[f"drop.emb.{s}"],
                                  dxsum=self.g(k(f"{s}_pe"))[:d] if pe_fused else None)
            if cfg.use_pe and not (kind != "image" and Ls[s] == 1):
                # d pe[l,:] = sum_b dE[b,l,:]  -> column sums of dE viewed as [B, L*d]
                ops.colsum_acc(de, B, Ls[s] * d, Ls[s] * d, self.g(k(f"{s}_pe"))[: Ls[s] * d], self.red_ws)
            if kind == "image":
                xin = ts["x_in"][s]
                self._linear_bwd(de, xin, Ts[s], d, xin.shape[1], k(f"{s}_proj.w"), k(f"{s}_proj.b"), None, need_dx=False, bias_done=True)
            else:
                tw_cols = d // 2 if s == "vid" else d
                gtab = self.g(k(f"{s}_proj.w"))
                if self.sparse_tables:
                    rows = torch.empty(B, tw_cols, device=self.device, dtype=torch.float32)
                    ops.id_rows_bwd(de, tw_cols, B, Ls[s], d, rows, dframe_w=self.g(k("frameid.w")) if s == "vid" else None,
                                    dframe_b=self.g(k("frameid.b")) if s == "vid" else None,
                                    frame_pos=ts.get("frame_pos") if s == "vid" else None)
                    self.table_rows.append((k(f"{s}_proj.w"), tw_cols, ts["ids"][s], rows))
                else:
                    ops.id_embed_bwd(de, ts["ids"][s], gtab.numel() // tw_cols, tw_cols, B, Ls[s], d, gtab,
                                     dframe_w=self.g(k("frameid.w")) if s == "vid" else None,
                                     dframe_b=self.g(k("frameid.b")) if s == "vid" else None,
                                     frame_pos=ts.get("frame_pos") if s == "vid" else None)
        if on_ready is not None:
            on_ready(self.groups[k("vid_proj.w")][0])
