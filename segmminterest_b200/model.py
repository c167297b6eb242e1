"""Drop-in replacements for the reference's model classes.

`SegFormerX` and `MultiScaleTemporalDetrLeaveFocal` keep the reference's constructor
keywords, forward keywords, output dict and `state_dict()` key schema
(models/encoder.py:327-520, models/decoder_leave_focal.py:425-658) so that
main_for_seq_leave_earlystop_SegMM.py can build and train them unchanged.  The sub-modules
below are *parameter containers only* (their own forward is never called): all compute runs
in `engine.Engine` on hand-written sm_100a kernels through the C ABI.  There is no eager /
CPU fallback -- calling the model without CUDA raises.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib
from .engine import Engine, EngineConfig, attn_ablation

PHOTO_MAX = 40
# loss name -> slot of the fused loss kernel's scalar vector (include/mmi_b200.h, mmi_loss_fwd_bwd)
LOSS_SLOT = {"focal": 0, "interestBPR": 4, "huber": 5, "hazard": 6, "surviveCE": 7, "interestCE": 8, "interestKL": 9}


def _init_bert(module):
    """models/encoder.py:412-423 (applied to the whole backbone after construction)."""
    if isinstance(module, (nn.Linear, nn.Embedding, nn.Conv1d)):
        module.weight.data.normal_(mean=0.0, std=0.02)
    if isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)
    if isinstance(module, nn.Linear) and module.bias is not None:
        module.bias.data.zero_()


def _linear3(d):
    return nn.ModuleList([nn.Linear(d, d) for _ in range(3)])


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: compute runs in segmminterest_b200.engine.Engine")


class SegFormerXAttention(_Container):
    """Parameters of models/encoder.py:12-42 (same attribute names and order)."""

    def __init__(self, d_model, num_head, sr_ratio=1, dropout=0.1, ablation_type="ours"):
        super().__init__()
        if sr_ratio != 1:
            raise NotImplementedError("sr_ratio > 1 is never used by the reference drivers (main...SegMM.py:94)")
        self.t2v_proj = _linear3(d_model)
        self.v2v_proj = _linear3(d_model)
        self.t2t_proj = _linear3(d_model)
        self.v2t_proj = _linear3(d_model)
        self.ff_usr = nn.Linear(d_model, d_model)
        self.ff_vid = nn.Linear(d_model, d_model)
        self.ln_usr = nn.LayerNorm(d_model, 1e-12)
        self.ln_vid = nn.LayerNorm(d_model, 1e-12)
        self.num_head = num_head
        self.d_head = d_model // num_head


class MLP(_Container):
    """Parameters of kn_util/nn_utils/layers/mlp.py:6-15."""

    def __init__(self, dims):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])


class SegFormerXEncoderLayer(_Container):
    """Parameters of models/encoder.py:178-187."""

    def __init__(self, d_model, num_head, ff_dim, sr_ratio, dropout, ablation_type="ours"):
        super().__init__()
        self.cross_attn = SegFormerXAttention(d_model, num_head, sr_ratio, dropout, ablation_type)
        self.ff_usr = MLP([d_model, ff_dim, d_model])
        self.ff_vid = MLP([d_model, ff_dim, d_model])
        self.ln_usr = nn.LayerNorm(d_model, eps=1e-12)
        self.ln_vid = nn.LayerNorm(d_model, eps=1e-12)


class SegFormerXEncoder(_Container):
    """Parameters of models/encoder.py:254-285, including the never-called pe_lns /
    txt_lvl_projs / patch_merge so checkpoints round-trip."""

    def __init__(self, d_model_in, d_model_lvls, num_head_lvls, sr_ratio_lvls, ff_dim_lvls, use_patch_merge, dropout,
                 ablation_type="ours"):
        super().__init__()
        assert len(d_model_lvls) == len(num_head_lvls) == len(sr_ratio_lvls) == len(ff_dim_lvls)
        self.layers = nn.ModuleList([
            SegFormerXEncoderLayer(d, h, ff, sr, dropout, ablation_type)
            for d, h, sr, ff in zip(d_model_lvls, num_head_lvls, sr_ratio_lvls, ff_dim_lvls)])
        lv = [d_model_in] + list(d_model_lvls)
        self.pe_lns = nn.ModuleList([nn.LayerNorm(d_model_lvls[i], 1e-12) for i in range(len(d_model_lvls))])
        self.txt_lvl_projs = nn.ModuleList([
            nn.Sequential(nn.Linear(lv[i - 1], lv[i]), nn.LayerNorm(lv[i], eps=1e-12)) for i in range(1, len(lv))])
        self.use_patch_merge = use_patch_merge
        self.patch_merge = nn.ModuleList([
            nn.Conv1d(lv[i - 1], lv[i], kernel_size=3, stride=2, padding=1) for i in range(1, len(lv))])


class MLP_Block(_Container):
    """Parameters of models/encoder.py:210-252 as SegFormerX builds it for the MLP ablations (:392-400): per hidden unit
    Linear -> ReLU -> Dropout inside one nn.Sequential named `mlp`, then the output Linear -- so the Linear layers sit at
    indices 0, 3, 6, ... of `mlp`, like in the reference's state_dict."""

    def __init__(self, input_dim, hidden_units, output_dim, dropout_rates=0.0):
        super().__init__()
        layers = []
        dims = [input_dim] + list(hidden_units)
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1]))
            layers.append(nn.ReLU())
            if dropout_rates > 0:
                layers.append(nn.Dropout(p=dropout_rates))
        layers.append(nn.Linear(dims[-1], output_dim))
        self.mlp = nn.Sequential(*layers)
        self.dropout_p = float(dropout_rates)

    def linears(self):
        return [m for m in self.mlp if isinstance(m, nn.Linear)]

    def linear_indices(self):
        return [i for i, m in enumerate(self.mlp) if isinstance(m, nn.Linear)]


MLP_ABLATIONS = ("SelfMLP", "CrossMLP", "w/oAtt")      # compared with == in the reference (encoder.py:392-400,503-511)


class SegFormerX(_Container):
    """Constructor signature of models/encoder.py:330-350."""

    def __init__(self, d_model_in=128, d_model_lvls=[128, 256, 512, 1024], num_head_lvls=[2, 4, 8, 16],
                 ff_dim_lvls=[256, 512, 1024, 2048], sr_ratio_lvls=[8, 4, 2, 1], input_vid_dim=768, input_usr_dim=768,
                 max_vid_len=256, max_usr_len=20, dropout=0.1, pe_kernel_size=3,
                 use_patch_merge=[True, False, True, False], output_layers=None, model_cfg=None, user_id_max=-1,
                 video_id_max=-1, use_pe=1):
        super().__init__()
        abl = getattr(model_cfg, "ablation_type", "ours") if model_cfg is not None else "ours"
        if abl not in ("ours", "CrossAtt", "SelfAtt", "noUser", "noUser_SelfAtt", "noPos") + MLP_ABLATIONS:
            # every choice of the driver's --ablation_type (main...SegMM.py:532): 'noUser*' only changes what the DRIVER feeds
            # (random user features, :275-277); 'SelfMLP' / 'CrossMLP' / 'w/oAtt' replace the encoder by an MLP_Block
            # (encoder.py:392-400,503-511); 'noPos' draws a random permutation of the frame positions per call (:428-429)
            raise NotImplementedError(f"ablation_type={abl!r} is not one of the reference's choices")
        if any(use_patch_merge) or any(s != 1 for s in sr_ratio_lvls):
            raise NotImplementedError("patch_merge / sr_ratio>1 are never enabled by the reference drivers")
        if any(d != d_model_in for d in d_model_lvls) or any(f != d_model_in for f in ff_dim_lvls) \
                or len(set(num_head_lvls)) != 1:
            raise NotImplementedError("all levels must share d_model / ff_dim / nhead (main...SegMM.py:89-91)")
        if output_layers not in ([-1], (-1,)):
            raise NotImplementedError("output_layers must be [-1] (main...SegMM.py:95)")
        # models/encoder.py:352-362: an id_max != -1 turns the projection into an embedding table
        if video_id_max != -1:
            self.vid_proj = nn.Embedding(video_id_max + 1, d_model_in // 2)
            self.frameid_proj = nn.Linear(1, d_model_in // 2)
        else:
            self.vid_proj = nn.Linear(input_vid_dim, d_model_in)
        if user_id_max != -1:
            self.usr_proj = nn.Embedding(user_id_max + 1, d_model_in)
        else:
            self.usr_proj = nn.Linear(input_usr_dim, d_model_in)
        self.debug = getattr(model_cfg, "debug", 0)
        self.num_layers_enc = len(d_model_lvls)
        self.use_pe = use_pe
        self.vid_pe = nn.Embedding(max_vid_len, d_model_in)
        self.usr_pe = nn.Embedding(max_usr_len, d_model_in)
        self.vid_ln = nn.LayerNorm(d_model_in, eps=1e-12)
        self.usr_ln = nn.LayerNorm(d_model_in, eps=1e-12)
        self.ablation_type = abl
        if abl == "CrossMLP":
            self.encoder_mlp = MLP_Block(d_model_lvls[0], list(d_model_lvls[2:-2]), d_model_lvls[0], dropout_rates=dropout)
            self.encoder_pooling = nn.AdaptiveAvgPool1d(40)
        elif abl in ("SelfMLP", "w/oAtt"):
            self.encoder_mlp = MLP_Block(d_model_lvls[0], list(d_model_lvls[1:-1]), d_model_lvls[0], dropout_rates=dropout)
        else:
            self.encoder = SegFormerXEncoder(d_model_in, list(d_model_lvls), list(num_head_lvls), list(sr_ratio_lvls),
                                             list(ff_dim_lvls), list(use_patch_merge), dropout, abl)
        self.output_layers = list(output_layers)
        self.dropout_p = dropout
        self.d_model = d_model_in
        self.nhead = num_head_lvls[0]
        self.input_vid_dim, self.input_usr_dim = input_vid_dim, input_usr_dim
        self.max_vid_len, self.max_usr_len = max_vid_len, max_usr_len
        self.apply(_init_bert)


class _EngineStep(torch.autograd.Function):
    """Autograd anchor: forward is already done by the engine; backward runs the engine's
    hand-written backward, which writes parameter gradients straight into the flat gradient
    buffer the Parameters' .grad alias (see Engine.bind_grads)."""

    @staticmethod
    def forward(ctx, anchor, loss, engine):
        ctx.engine = engine
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.engine.backward(grad_out.contiguous().float())
        return None, None, None


class InteractionAggregation(_Container):
    """Parameters of models/decoder_leave_focal.py:392-410 (output_dim 1): w_x, w_y Linear(d, 1) and the bilinear
    weight stored flat [H * dx * dy, 1]."""

    def __init__(self, x_dim, y_dim, output_dim=1, num_heads=1):
        super().__init__()
        if output_dim != 1 or x_dim != y_dim or num_heads <= 0 or x_dim % num_heads:
            raise NotImplementedError("InteractionAggregation: output_dim 1, x_dim == y_dim divisible by num_heads > 0")
        self.num_heads, self.output_dim = num_heads, output_dim
        self.w_x = nn.Linear(x_dim, output_dim)
        self.w_y = nn.Linear(y_dim, output_dim)
        for lin in (self.w_x, self.w_y):                      # kn_util/nn_utils/init.py:52-62
            nn.init.xavier_uniform_(lin.weight.data)
            lin.bias.data.zero_()
        self.head_x_dim = x_dim // num_heads
        self.head_y_dim = y_dim // num_heads
        self.w_xy = nn.Parameter(torch.empty(num_heads * self.head_x_dim * self.head_y_dim, output_dim))
        nn.init.xavier_normal_(self.w_xy)


class MultiScaleTemporalDetrLeaveFocal(nn.Module):
    """models/decoder_leave_focal.py:425-658: ctor (backbone1, backbone2, head, frame_pooler, model_cfg); forward
    keywords and return dict unchanged.  One backbone + Linear head, or two backbones (the reference's default
    'both' configuration: image features + ID embeddings) fused by InteractionAggregation; losses `focal` and
    `interestBPR`; optional learnable position bias.  `precision` ('fp32' strict parity | 'bf16' tensor-core) is read
    from model_cfg.mmi_precision when present (default 'fp32')."""

    def __init__(self, backbone1, backbone2, head, frame_pooler, model_cfg) -> None:
        super().__init__()
        if head is not None:
            raise NotImplementedError("head must be None (main...SegMM.py:106,129)")
        it = getattr(model_cfg, "input_type", {"user": "image", "photo": "image"})
        for lt in model_cfg.loss_type_list:     # names the reference's compute_loss does not know are silently skipped there
            if lt not in LOSS_SLOT:             # (and then KeyError in the weighted sum, :561-566); fail early instead
                raise ValueError(f"unknown loss_type {lt!r}; the reference knows {sorted(LOSS_SLOT)}")
        self.backbone1 = backbone1
        self.backbone2 = backbone2
        self.model_cfg = model_cfg
        self.head = head
        self.frame_pooler = frame_pooler
        self.debug = model_cfg.debug
        self.input_type = it
        self.bias_weight = None
        self.bias_bias = None
        if getattr(model_cfg, "learnable_bias", 0):   # models/decoder_leave_focal.py:442-444
            self.bias_weight = nn.Parameter(torch.ones(1, PHOTO_MAX), requires_grad=True)
            self.bias_bias = nn.Parameter(torch.ones(1, PHOTO_MAX), requires_grad=True)
        self.exposure_prob = model_cfg.exposure_prob
        d_model = model_cfg.d_model
        if backbone2 is None:
            self.stage_mlp1 = nn.Linear(d_model, 1)
            nn.init.xavier_uniform_(self.stage_mlp1.weight.data)  # kn_util/nn_utils/init.py:52-62
            self.stage_mlp1.bias.data.zero_()
        else:                                          # models/decoder_leave_focal.py:457-470
            fh = getattr(model_cfg, "fusion_heads", 2)
            if fh > 0:
                self.fusion_module = InteractionAggregation(d_model, d_model, output_dim=1, num_heads=fh)
            elif fh in (0, -1, -2, -3):
                # -3 concatenates the two backbones' output LISTS and takes [-1] (decoder_leave_focal.py:621-623), i.e. it
                # scores backbone2 alone with a Linear(d, 1) head; backbone1 runs in the reference but never reaches the loss
                self.stage_mlp1 = nn.Linear(2 * d_model if fh == -1 else d_model, 1)
                heads = [self.stage_mlp1]
                if fh == 0:
                    self.stage_mlp2 = nn.Linear(d_model, 1)
                    heads.append(self.stage_mlp2)
                for lin in heads:
                    nn.init.xavier_uniform_(lin.weight.data)
                    lin.bias.data.zero_()
            else:
                raise NotImplementedError(f"fusion_heads={fh}: the reference defines > 0 (InteractionAggregation), 0, -1, -2, -3")
            self.fusion_heads = fh
        self._engine = None
        self.precision = getattr(model_cfg, "mmi_precision", "fp32")

    # -- engine plumbing ---------------------------------------------------------------
    def engine(self) -> Engine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.MMIError("MultiScaleTemporalDetrLeaveFocal (b200) needs the model on a CUDA device: "
                                "there is no CPU fallback")
        bb = self.backbone1
        if self._engine is None or self._engine.device != dev or self._engine.cfg.precision != self.precision:
            cfg = EngineConfig(d_model=bb.d_model, nhead=bb.nhead, num_layers=bb.num_layers_enc,
                               din_vid=bb.input_vid_dim, din_usr=bb.input_usr_dim, max_usr_len=bb.max_usr_len,
                               max_vid_len=bb.max_vid_len, use_pe=bool(bb.use_pe), precision=self.precision,
                               ablation=attn_ablation(bb.ablation_type), no_pos="noPos" in (bb.ablation_type or ""))
            self._engine = Engine(cfg, self, dev)
        self._engine.ensure_bound()
        return self._engine

    def loss_cfg(self):
        lw = self.model_cfg.loss_weight
        names = list(self.model_cfg.loss_type_list)
        # like the reference (:561-566) a missing weight is a KeyError; huber is weighted by loss_weight['mse']
        others = {n: float(lw["mse" if n == "huber" else n]) for n in names if n not in ("focal", "interestBPR")}

        def after_focal(n):
            return n in names and "focal" in names and names.index("focal") < names.index(n)

        return dict(use_focal="focal" in names, w_focal=float(lw["focal"]) if "focal" in names else 0.0,
                    use_bpr="interestBPR" in names, w_bpr=float(lw["interestBPR"]) if "interestBPR" in names else 0.0,
                    others=others, mask_loss=int(getattr(self.model_cfg, "mask_loss", 0)),
                    ce_after_focal=after_focal("interestCE"), kl_after_focal=after_focal("interestKL"))

    def dropout_p(self):
        """Dropout probability of the next forward: nn.Module semantics -- the constructor value of the backbone
        (the reference driver never passes one, so 0.1: models/encoder.py:342, main...SegMM.py:88-104) under train(), 0
        under eval().  The FFN's inner dropout is MLP's own default 0.1 whatever the backbone was given
        (kn_util/nn_utils/layers/mlp.py:8; models/encoder.py:183-184) -- it only differs when a caller overrides p."""
        return float(self.backbone1.dropout_p) if self.training else 0.0

    def forward(self, usr_image, usr_id, usr_mask, vid_image, vid_id, vid_mask, gt=None, mode="train", **kwargs):
        eng = self.engine()
        eng.drop_p = self.dropout_p()
        B = usr_id.shape[0]
        logits = eng.forward(usr_image, usr_mask, vid_image, vid_mask, usr_id=usr_id, vid_id=vid_id,
                             need_bwd=mode != "inference" and torch.is_grad_enabled())
        if mode == "inference":
            if self.bias_weight is None:
                return dict(logits=logits.clone(), gt=gt)
            # logits + (pos+1) * bias_weight + bias_bias (models/decoder_leave_focal.py:650-658): the loss kernel adds it
            g = gt if gt is not None else torch.full((B, PHOTO_MAX), -2, dtype=torch.int64, device=logits.device)
            _, lb = eng.loss(logits, g.clone().contiguous(), self.exposure_prob, 1.0 / B,
                             dict(use_focal=False, w_focal=0.0, use_bpr=False, w_bpr=0.0), need_grad=False)
            return dict(logits=lb.clone(), gt=gt)
        if mode not in ("train", "test"):
            return None
        cfg = self.loss_cfg()
        scal, lb = eng.loss(logits, gt, self.exposure_prob, 1.0 / B, cfg, need_grad=torch.is_grad_enabled())
        loss = scal[3]
        if torch.is_grad_enabled():
            loss = _EngineStep.apply(eng.anchor, loss, eng)
        out = {name: scal[LOSS_SLOT[name]].clone() for name in self.model_cfg.loss_type_list}
        out.update({"mse": scal[1].clone(), "mse2": scal[2].clone(), "loss": loss, "logits": lb.clone(), "gt": gt})
        return out


class QueryBasedDecoder(nn.Module):
    """Imported by the driver (main...SegMM.py:5) but defined nowhere in the reference and
    never used; placeholder so the import surface is complete."""


def reference_state_shapes(d_model=512, num_layers=6, din=1024, max_usr_len=100, max_vid_len=40):
    """name -> shape of the reference state_dict (image modality, single backbone), in the
    reference's own order (dumped from the unmodified reference)."""
    d = d_model
    s = {}
    p = "backbone1."
    s[p + "vid_proj.weight"] = (d, din); s[p + "vid_proj.bias"] = (d,)
    s[p + "usr_proj.weight"] = (d, din); s[p + "usr_proj.bias"] = (d,)
    s[p + "vid_pe.weight"] = (max_vid_len, d); s[p + "usr_pe.weight"] = (max_usr_len, d)
    for n in ("vid_ln", "usr_ln"):
        s[p + n + ".weight"] = (d,); s[p + n + ".bias"] = (d,)
    for i in range(num_layers):
        q = f"{p}encoder.layers.{i}."
        for blk in ("t2v", "v2v", "t2t", "v2t"):
            for j in range(3):
                s[f"{q}cross_attn.{blk}_proj.{j}.weight"] = (d, d)
                s[f"{q}cross_attn.{blk}_proj.{j}.bias"] = (d,)
        for n in ("ff_usr", "ff_vid"):
            s[f"{q}cross_attn.{n}.weight"] = (d, d); s[f"{q}cross_attn.{n}.bias"] = (d,)
        for n in ("ln_usr", "ln_vid"):
            s[f"{q}cross_attn.{n}.weight"] = (d,); s[f"{q}cross_attn.{n}.bias"] = (d,)
        for n in ("ff_usr", "ff_vid"):
            for j in range(2):
                s[f"{q}{n}.layers.{j}.weight"] = (d, d); s[f"{q}{n}.layers.{j}.bias"] = (d,)
        for n in ("ln_usr", "ln_vid"):
            s[f"{q}{n}.weight"] = (d,); s[f"{q}{n}.bias"] = (d,)
    for i in range(num_layers):
        s[f"{p}encoder.pe_lns.{i}.weight"] = (d,); s[f"{p}encoder.pe_lns.{i}.bias"] = (d,)
    for i in range(num_layers):
        s[f"{p}encoder.txt_lvl_projs.{i}.0.weight"] = (d, d); s[f"{p}encoder.txt_lvl_projs.{i}.0.bias"] = (d,)
        s[f"{p}encoder.txt_lvl_projs.{i}.1.weight"] = (d,); s[f"{p}encoder.txt_lvl_projs.{i}.1.bias"] = (d,)
    for i in range(num_layers):
        s[f"{p}encoder.patch_merge.{i}.weight"] = (d, d, 3); s[f"{p}encoder.patch_merge.{i}.bias"] = (d,)
    s["stage_mlp1.weight"] = (1, d); s["stage_mlp1.bias"] = (1,)
    return s


def build_model(args, din=1024, max_usr_len=100, max_vid_len=40, dropout=0.1, n_users=0, n_items=0):
    """init_model() of main_for_seq_leave_earlystop_SegMM.py:60-130 for every input_type (image / id / both);
    n_users / n_items stand for reader.n_users / reader.n_items."""
    n = args.num_layers_enc
    it = getattr(args, "input_type", {"user": "image", "photo": "image"})

    def bb(user_id_max, video_id_max, usr_len):
        return SegFormerX(d_model_in=args.d_model, d_model_lvls=[args.d_model] * n, num_head_lvls=[args.nhead] * n,
                          ff_dim_lvls=[args.d_model] * n, input_vid_dim=din, input_usr_dim=din, max_vid_len=max_vid_len,
                          max_usr_len=usr_len, sr_ratio_lvls=[1] * n, use_patch_merge=[False] * n, output_layers=[-1],
                          model_cfg=args, user_id_max=user_id_max, video_id_max=video_id_max, use_pe=args.use_pe, dropout=dropout)

    if it["user"] == "both" or it["photo"] == "both":
        u1, l1, u2, l2 = {"both": (-1, max_usr_len, n_users, 1), "id": (n_users, 1, n_users, 1),
                          "image": (-1, max_usr_len, -1, max_usr_len)}[it["user"]]
        v1, v2 = {"both": (-1, n_items), "id": (n_items, n_items), "image": (-1, -1)}[it["photo"]]
        return MultiScaleTemporalDetrLeaveFocal(bb(u1, v1, l1), bb(u2, v2, l2), None, nn.Identity(), args)
    u1, l1 = (n_users, 1) if it["user"] == "id" else (-1, max_usr_len)
    v1 = n_items if it["photo"] == "id" else -1
    return MultiScaleTemporalDetrLeaveFocal(bb(u1, v1, l1), None, None, nn.Identity(), args)
