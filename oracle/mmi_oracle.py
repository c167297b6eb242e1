"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's MMinterest
training-step arithmetic, written as plain functional
PyTorch on a ``state_dict``.  It is the checker the CUDA path is compared with;
nothing in ``segmminterest_b200/`` may import it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this file against
fixtures under ``tests/golden/`` that were produced by running the reference's
own unmodified ``models/encoder.py`` + ``models/decoder_leave_focal.py``
(``oracle/make_golden.py`` via ``oracle/ref_shim.py``), torch 2.11.0 CPU.
The reference ships no tests / golden vectors of its own.

Each function cites the reference file:line it restates (paths relative to
/root/reference/MMinterest).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# a-3  L1 normalisation of gathered features
# --------------------------------------------------------------------------
def l1_normalise(x: torch.Tensor) -> torch.Tensor:
    """main_for_seq_leave_earlystop_SegMM.py:272-273: x / (||x||_1 + 1e-6)."""
    return x / (x.norm(p=1, dim=-1, keepdim=True) + 1e-6)


# --------------------------------------------------------------------------
# a-4  embedding: Linear(Din->d) + PE + LN(eps 1e-12)
# --------------------------------------------------------------------------
# Dropout hook.  The reference's five kinds of nn.Dropout sites are restated with an optional callable
#     drop(kind, tower, layer, side, x, blocks=None) -> dropout(x)
# (kind: DROP_* below; tower 0/1 = backbone1/2; layer 63 = the embedding; side 0 = candidate, 1 = history; blocks = key
# block widths of a concatenated logits tensor).  None = eval mode.  torch's own generator cannot be replayed by any other
# implementation, so the parity tests pass a callable that applies the masks of oracle/dropout_ref.py -- the counter-based
# generator of the CUDA path -- which makes "training with dropout" a deterministic comparison.
DROP_EMB, DROP_ATTN, DROP_ATTN_OUT, DROP_MLP1, DROP_MLP2 = range(5)
DROP_BLOCK = 5      # hidden layers of the MLP_Block encoder of the SelfMLP / CrossMLP ablations (layer = hidden-layer index)
MLP_ABLATIONS = ("SelfMLP", "CrossMLP", "w/oAtt")     # compared with == in the reference (encoder.py:392-400,503-511)


def _tower(prefix):
    return 1 if prefix.startswith("backbone2.") else 0


def torch_dropout(p=0.1):
    """the reference's own behaviour (F.dropout from torch's global generator) as a drop callable: CPU baseline timing"""
    def drop(kind, tower, layer, side, x, blocks=None):
        return F.dropout(x, p, True)
    return drop


def embed(sd, prefix, usr, vid, use_pe=True, drop=None, no_pos=False):
    """models/encoder.py:425-473.  Image inputs are [B,L,Din] floats (Linear projections); ID inputs
    are int64 [B] (encoder.py:352-362,426-435,478-488): the video id is repeated over the 40 segments
    and embedded into d/2 columns, the other d/2 come from Linear(1 -> d/2) of the segment position; the user id
    becomes ONE token (the caller then uses an all-ones user mask)."""
    d = sd[prefix + "vid_ln.weight"].shape[0]
    if vid.ndim == 1:
        B, Lv = vid.shape[0], 40
        emb = sd[prefix + "vid_proj.weight"][vid][:, None, :].expand(B, Lv, -1)
        if no_pos:      # 'noPos' (encoder.py:428-429): one torch.randperm(Lv) per interaction from the default generator
            pos = torch.stack([torch.randperm(Lv) for _ in range(B)]).to(emb.dtype)[:, :, None]
        else:
            pos = torch.arange(Lv, dtype=emb.dtype)[None, :, None].expand(B, Lv, 1)
        fr = F.linear(pos, sd[prefix + "frameid_proj.weight"], sd[prefix + "frameid_proj.bias"])
        v = torch.cat([emb, fr], -1)
    else:
        v = F.linear(vid, sd[prefix + "vid_proj.weight"], sd[prefix + "vid_proj.bias"])
    if usr.ndim == 1:
        u = sd[prefix + "usr_proj.weight"][usr][:, None, :]
    else:
        u = F.linear(usr, sd[prefix + "usr_proj.weight"], sd[prefix + "usr_proj.bias"])
    if use_pe:
        v = v + sd[prefix + "vid_pe.weight"][None, : v.shape[1]]
        u = u + sd[prefix + "usr_pe.weight"][None, : u.shape[1]]
    v = F.layer_norm(v, (d,), sd[prefix + "vid_ln.weight"], sd[prefix + "vid_ln.bias"], 1e-12)
    u = F.layer_norm(u, (d,), sd[prefix + "usr_ln.weight"], sd[prefix + "usr_ln.bias"], 1e-12)
    if drop is not None:            # encoder.py:458,470: self.do on both embeddings
        v = drop(DROP_EMB, _tower(prefix), 63, 0, v)
        u = drop(DROP_EMB, _tower(prefix), 63, 1, u)
    return v, u


# --------------------------------------------------------------------------
# a-5  one block of attention logits, with the "set to -10000" mask
# --------------------------------------------------------------------------
def attn_logits(sd, p, feat_k, mask_k, feat_q, mask_q, nhead):
    """models/encoder.py:44-73.  q = proj[0](feat_q), k = proj[1](feat_k);
    logits[b,h,i,j] = <q_i, k_j>; positions with mask_q[i] & mask_k[j] == 0 are
    SET to -10000 (not added)."""
    B, Lq, d = feat_q.shape
    Lk = feat_k.shape[1]
    dh = d // nhead
    q = F.linear(feat_q, sd[p + "0.weight"], sd[p + "0.bias"]).view(B, Lq, nhead, dh)
    k = F.linear(feat_k, sd[p + "1.weight"], sd[p + "1.bias"]).view(B, Lk, nhead, dh)
    logits = torch.einsum("bqhd,bkhd->bhqk", q, k)
    m = (mask_q[:, :, None] & mask_k[:, None, :])[:, None].expand(B, nhead, Lq, Lk)
    return torch.where(m, logits, torch.full_like(logits, -10000.0))


# --------------------------------------------------------------------------
# a-6  four-way attention + output projection + residual LN
# --------------------------------------------------------------------------
# which attention blocks feed the candidate ("vid") / history ("usr") queries per ablation (encoder.py:108-161,172-175)
ATTN_BLOCKS = {"ours": (("v2v", "t2v"), ("v2t", "t2t")), "CrossAtt": (("t2v",), ("v2t",)), "SelfAtt": (("v2v",), None)}


def attn_ablation(ablation_type):
    """the reference tests `'CrossAtt' in ablation_type` / `'SelfAtt' in ablation_type` (substring, so 'noUser_SelfAtt' is
    SelfAtt for the model; the 'noUser' half lives in the driver, main...SegMM.py:275-277)"""
    a = ablation_type or "ours"
    return "CrossAtt" if "CrossAtt" in a else ("SelfAtt" if "SelfAtt" in a else "ours")


def cross_attention(sd, p, vid, vid_mask, usr, usr_mask, nhead, need_usr=True, ablation="ours", drop=None, where=(0, 0)):
    """models/encoder.py:75-175 with sr_ratio=1; drop: the logits dropout (:145-150, on the concatenated logits AFTER the
    -10000 fill and BEFORE the scale -- a dropped masked logit becomes 0 and is visible to the softmax) and the dropout of
    the output projections (:163-164); where = (tower, layer).  'ours': joint softmax over [v2v | t2v] for candidate
    queries and [v2t | t2t] for history queries; 'CrossAtt' keeps t2v / v2t only, 'SelfAtt' keeps v2v only and returns
    no history update (:172-173).  The scale 1/sqrt(dh) is applied AFTER the mask fill."""
    B, Lv, d = vid.shape
    dh = d // nhead
    scale = 1.0 / math.sqrt(dh)
    key_side = {"v2v": (vid, vid_mask), "t2v": (usr, usr_mask), "v2t": (vid, vid_mask), "t2t": (usr, usr_mask)}
    names_v, names_t = ATTN_BLOCKS[ablation]

    def val(name, x):
        return F.linear(x, sd[p + name + "_proj.2.weight"], sd[p + name + "_proj.2.bias"])

    def scaled_logits(names, q, q_mask, side):
        blocks = [attn_logits(sd, p + n + "_proj.", key_side[n][0], key_side[n][1], q, q_mask, nhead) for n in names]
        logits = torch.cat(blocks, -1)
        if drop is not None:
            logits = drop(DROP_ATTN, where[0], where[1], side, logits, blocks=[b.shape[-1] for b in blocks])
        return logits * scale

    def attend(names, logits, Lq):
        value = torch.cat([val(n, key_side[n][0]) for n in names], 1).view(B, -1, nhead, dh)
        return torch.einsum("bhqk,bkhd->bqhd", F.softmax(logits, -1), value).reshape(B, Lq, d)

    # the reference's order of dropout draws (:145-150,163-164): candidate logits, history logits, history output
    # projection, candidate output projection -- kept so that torch_dropout() replays the reference's generator stream
    want_usr = need_usr and names_t is not None
    lv = scaled_logits(names_v, vid, vid_mask, 0)
    lt = scaled_logits(names_t, usr, usr_mask, 1) if want_usr else None
    usr_out = None
    if want_usr:
        usr_ = F.linear(attend(names_t, lt, usr.shape[1]), sd[p + "ff_usr.weight"], sd[p + "ff_usr.bias"])
        if drop is not None:
            usr_ = drop(DROP_ATTN_OUT, where[0], where[1], 1, usr_)
        usr_out = F.layer_norm(usr + usr_, (d,), sd[p + "ln_usr.weight"], sd[p + "ln_usr.bias"], 1e-12)
    vid_ = F.linear(attend(names_v, lv, Lv), sd[p + "ff_vid.weight"], sd[p + "ff_vid.bias"])
    if drop is not None:
        vid_ = drop(DROP_ATTN_OUT, where[0], where[1], 0, vid_)
    vid_out = F.layer_norm(vid + vid_, (d,), sd[p + "ln_vid.weight"], sd[p + "ln_vid.bias"], 1e-12)
    return vid_out, usr_out


# --------------------------------------------------------------------------
# a-7  FFN + residual LN
# --------------------------------------------------------------------------
def ffn(sd, p, side, x, drop=None, where=(0, 0)):
    """models/encoder.py:198-206 + kn_util/nn_utils/layers/mlp.py:17-24:
    LN(x + do(W2 do(gelu_erf(W1 x)))), eps 1e-12."""
    d = x.shape[-1]
    si = 0 if side == "vid" else 1
    h = F.gelu(F.linear(x, sd[p + f"ff_{side}.layers.0.weight"], sd[p + f"ff_{side}.layers.0.bias"]))
    if drop is not None:
        h = drop(DROP_MLP1, where[0], where[1], si, h)
    h = F.linear(h, sd[p + f"ff_{side}.layers.1.weight"], sd[p + f"ff_{side}.layers.1.bias"])
    if drop is not None:
        h = drop(DROP_MLP2, where[0], where[1], si, h)
    return F.layer_norm(x + h, (d,), sd[p + f"ln_{side}.weight"], sd[p + f"ln_{side}.bias"], 1e-12)


# --------------------------------------------------------------------------
# a-8  encoder stack; output = INPUT of the last layer
# --------------------------------------------------------------------------
def mlp_block(sd, prefix, x, drop=None, tower=0):
    """MLP_Block (models/encoder.py:210-254, after FuxiCTR) as SegFormerX builds it for the MLP ablations (:392-400):
    [Linear(d, d) -> ReLU -> Dropout] per hidden unit, then Linear(d, d); nn.Sequential indices 0, 3, 6, ... are the Linears."""
    idx = sorted(int(k[len(prefix + "encoder_mlp.mlp."):].split(".")[0]) for k in sd if k.startswith(prefix + "encoder_mlp.mlp.") and k.endswith(".weight"))
    for n, i in enumerate(idx):
        x = F.linear(x, sd[f"{prefix}encoder_mlp.mlp.{i}.weight"], sd[f"{prefix}encoder_mlp.mlp.{i}.bias"])
        if n < len(idx) - 1:
            x = F.relu(x)
            if drop is not None:
                x = drop(DROP_BLOCK, tower, n, 0, x)
    return x


def backbone(sd, prefix, usr, usr_mask, vid, vid_mask, nhead, num_layers, use_pe=True, ablation="ours", drop=None, full_usr=False,
             no_pos=False):
    """models/encoder.py:302-324,475-520.  intermediate_states records vid_feat
    BEFORE each layer and the caller takes [-1], so layer N-1 never reaches the
    output, nor does the history side of layer N-2."""
    v, u = embed(sd, prefix, usr, vid, use_pe, drop, no_pos)
    tw = _tower(prefix)
    # the MLP ablations replace the attention encoder (encoder.py:503-511); none of them looks at the masks
    if ablation == "CrossMLP":     # MLP over [history ; candidate] tokens, then AdaptiveAvgPool1d(40) along the token axis
        x = mlp_block(sd, prefix, torch.cat([u, v], dim=-2), drop, tw)
        return F.adaptive_avg_pool1d(x.permute(0, 2, 1), 40).permute(0, 2, 1)
    if ablation == "SelfMLP":      # MLP over the candidate tokens
        return mlp_block(sd, prefix, v, drop, tw)
    if ablation == "w/oAtt":       # the embedded candidate tokens go straight to the head (encoder_mlp is built but never called)
        return v
    if usr.ndim == 1:   # ID user: one token, mask of ones (encoder.py:478-481)
        usr_mask = torch.ones(usr.shape[0], 1, dtype=torch.bool)
    for i in range(num_layers - 1):
        p = f"{prefix}encoder.layers.{i}."
        need_usr = full_usr or i < num_layers - 2      # full_usr: also the dead history side of layer N-2 (its dropout DRAWS
                                                       # precede live ones in the reference's generator stream)
        v, u2 = cross_attention(sd, p + "cross_attn.", v, vid_mask, u, usr_mask, nhead, need_usr, ablation, drop, (tw, i))
        v = ffn(sd, p, "vid", v, drop, (tw, i))
        if u2 is not None:          # SelfAtt: the history tokens are never updated (encoder.py:320-321)
            u = ffn(sd, p, "usr", u2, drop, (tw, i))
    return v


# --------------------------------------------------------------------------
# a-9/a-10  head (+ optional learnable position bias)
# --------------------------------------------------------------------------
def _position_bias(sd, logits):
    if "bias_weight" in sd:
        pos = torch.arange(logits.shape[1], dtype=logits.dtype)
        logits = logits + (pos + 1) * sd["bias_weight"].reshape(1, -1) + sd["bias_bias"].reshape(1, -1)
    return logits


def head_logits(sd, x):
    """models/decoder_leave_focal.py:451,596 and :497-504."""
    logits = F.linear(x, sd["stage_mlp1.weight"], sd["stage_mlp1.bias"]).squeeze(-1)
    return _position_bias(sd, logits)


def fusion_logits(sd, x, y, num_heads):
    """InteractionAggregation.forward, models/decoder_leave_focal.py:411-423 (output_dim 1):
    w_x.x + w_y.y + sum_h x_h^T W_h y_h with W stored flat [H*dx*dy, 1] viewed [H, dx, dy]."""
    B, L, d = x.shape
    out = F.linear(x, sd["fusion_module.w_x.weight"], sd["fusion_module.w_x.bias"]) + \
        F.linear(y, sd["fusion_module.w_y.weight"], sd["fusion_module.w_y.bias"])
    dx = d // num_heads
    W = sd["fusion_module.w_xy"].view(num_heads, dx, dx)
    xh = x.reshape(B * L, num_heads, dx)
    yh = y.reshape(B * L, num_heads, dx)
    xy = torch.einsum("rhi,hij,rhj->r", xh, W, yh)
    return _position_bias(sd, out.squeeze(-1) + xy.view(B, L))


# --------------------------------------------------------------------------
# a-11/a-12  loss (focal) + diagnostics
# --------------------------------------------------------------------------
def focal_loss_sum(logits, gt_mut, mask, exposure_prob, bsz):
    """models/decoder_leave_focal.py:35-59,533-538: alpha 0.5, gamma 2, masked
    sum / batch size.  ``gt_mut`` is gt after the in-place rewrite
    (gt>0 -> 1, gt==-1 -> 0; -2 stays)."""
    t = gt_mut.to(logits.dtype)
    ep = torch.as_tensor(exposure_prob, dtype=logits.dtype)[None, : logits.shape[1]]
    p = torch.sigmoid(logits) * ep
    ce = F.binary_cross_entropy_with_logits(logits, t, reduction="none")
    p_t = p * t + (1 - p) * (1 - t)
    loss = ce * (1 - p_t) ** 2
    alpha_t = 0.5 * t + 0.5 * (1 - t)
    loss = alpha_t * loss
    return (loss * mask.to(loss.dtype)).sum() / bsz


def interest_bpr_all(logits, gt):
    """models/decoder_leave_focal.py:163-221 (default loss `interestBPR`)."""
    view = (gt == 1).sum(1)
    rows = view < 40
    lg = logits[rows]
    vl = view[rows]
    n = lg.shape[0]
    pos = lg[torch.arange(n), vl]
    neg_mask = torch.ones_like(lg, dtype=torch.bool)
    neg_mask[torch.arange(n), vl] = False
    neg = lg[neg_mask].view(n, -1)
    neg_softmax = (neg - neg.max()).softmax(1)
    soft = (neg - pos[:, None]).sigmoid() * neg_softmax
    return -(soft.sum(1)).clamp(min=1e-8, max=1 - 1e-8).log().mean()


def huber_broadcast(hazard_masked, view_lengths, delta=1.0):
    """`huber` (models/decoder_leave_focal.py:61-66, call :539-540): the prediction is [B] (expected leave position =
    sum of masked hazards) and the target [B,1], so -- like `mse` -- the error broadcasts to [B,B]: err[i,j] = H_j - v_i."""
    err = hazard_masked.sum(1)[None, :] - view_lengths
    a = err.abs()
    return torch.where(a < delta, 0.5 * err * err, delta * (a - 0.5 * delta)).mean()


def partial_likelihood(hazard_masked, view_lengths):
    """`hazard` (compute_partial_likelihood_loss :273-286): rows whose observed time (= number of watched segments)
    equals 40 are skipped; -(log(h[ot]+1e-6) - log(sum_{t>=ot} h[t] + 1e-6)) summed, / n_samples (all rows)."""
    B, L = hazard_masked.shape
    ot = view_lengths.view(-1).long()
    rows = ot != 40
    idx = torch.arange(L)[None, :]
    at = (hazard_masked * (idx == ot[:, None])).sum(1)
    risk = (hazard_masked * (idx >= ot[:, None])).sum(1)
    ll = torch.log(at + 1e-6) - torch.log(risk + 1e-6)
    return -(ll * rows.to(ll.dtype)).sum() / B


def leave_prob_ce(h_t, gt_binary, mask):
    """`surviveCE` (compute_leave_prob_CE :68-97): BCE-with-logits fed exp(h_t) -- the survival probability -- as the
    *logit*, masked, sum / number of valid positions in the batch."""
    ce = F.binary_cross_entropy_with_logits(torch.exp(h_t), gt_binary, reduction="none")
    m = mask.to(ce.dtype)
    return (ce * m).sum() / m.sum()


def interest_leave_ce(logits, gt_cur, mask, kind, use_mask):
    """`interestCE` / `interestKL` (compute_interest_leave_CE :99-161).  gt_cur is gt *as it is when the loss runs*
    (after focal's in-place rewrite if 'focal' precedes it in loss_type_list)."""
    nonleave = (gt_cur != 0).to(logits.dtype)
    log_q = torch.log(logits.softmax(1))
    tgt = nonleave.softmax(1)
    m = mask.to(logits.dtype)
    if kind == "CE":
        if use_mask:
            return (-(m * tgt * log_q).sum(1) / m.sum(1)).mean()
        return (-(tgt * log_q).sum(1)).mean()
    kl = tgt * (torch.log(tgt) - log_q)
    if use_mask:
        return ((kl * m).sum(1) / m.sum(1)).mean()
    return kl.sum() / logits.shape[0]


def compute_loss(logits, gt, exposure_prob, loss_type_list=("focal",), loss_weight=None, mask_loss=0):
    """models/decoder_leave_focal.py:490-572.  Returns the same dict (tensors).
    ``gt`` is NOT modified here; the returned ``gt`` is the rewritten copy the
    reference leaves behind when 'focal' is in the list."""
    loss_weight = loss_weight or {}
    B = gt.shape[0]
    mask = gt != -2
    p = torch.sigmoid(logits)
    h_t = torch.cumsum(torch.log(p), 1)
    survival = torch.exp(h_t)
    view_lengths = (gt == 1).to(logits.dtype).sum(1, keepdim=True)
    durations = mask.sum(1)
    survival_masked = torch.where(mask, survival, torch.zeros_like(survival))
    hazard_masked = torch.where(mask, 1 - survival, torch.zeros_like(survival))
    out = {}
    gt_cur = gt.clone()
    for name in loss_type_list:
        if name == "focal":
            gt_cur = torch.where(gt_cur > 0, torch.ones_like(gt_cur), gt_cur)
            gt_cur = torch.where(gt_cur == -1, torch.zeros_like(gt_cur), gt_cur)
            out["focal"] = focal_loss_sum(logits, gt_cur, mask, exposure_prob, B)
        elif name == "interestBPR":
            out["interestBPR"] = interest_bpr_all(logits, gt)
        elif name == "huber":
            out["huber"] = huber_broadcast(hazard_masked, view_lengths)
        elif name == "hazard":
            out["hazard"] = partial_likelihood(hazard_masked, view_lengths)
        elif name == "surviveCE":
            out["surviveCE"] = leave_prob_ce(h_t, (gt == 1).to(logits.dtype), mask)
        elif name in ("interestCE", "interestKL"):
            out[name] = interest_leave_ce(logits, gt_cur, mask, name[-2:], mask_loss)
        else:
            raise NotImplementedError(name)
    s = survival_masked.sum(1)  # [B]
    # nn.MSELoss()([B], [B,1]) broadcasts to [B,B] (reference quirk, :552)
    out["mse"] = ((s[None, :] - view_lengths) ** 2).mean()
    sm2 = survival_masked.clone()
    sm2[torch.arange(B), (durations - 1) % sm2.shape[1]] = 1.0  # :554-555 (index -1 wraps)
    view_lengths2 = (gt_cur >= 0).sum(1, keepdim=True).to(logits.dtype)
    out["mse2"] = ((sm2.sum(1)[None, :] - view_lengths2) ** 2).mean()
    loss = 0.0
    for name in loss_type_list:
        loss = loss + out[name] * loss_weight.get("mse" if name == "huber" else name, 1.0)   # :563-564
    out["loss"] = loss
    out["logits"] = logits
    out["gt"] = gt_cur
    return out


# --------------------------------------------------------------------------
# full forward (image modality, single backbone)
# --------------------------------------------------------------------------
def forward(sd, usr_image, usr_mask, vid_image, vid_mask, gt, *, nhead, num_layers,
            exposure_prob=None, loss_type_list=("focal",), use_pe=True, mode="train", loss_weight=None,
            usr_id=None, vid_id=None, input_type=None, fusion_heads=2, mask_loss=0, ablation_type="ours", drop=None,
            full_usr=False):
    """models/decoder_leave_focal.py:574-658.  input_type {'user': image|id|both, 'photo': image|id|both}
    (default image/image, single backbone, Linear head); with a 'both' entry there are two backbones
    (main...SegMM.py:63-106: backbone1 takes the image side of a 'both' input, backbone2 the id side) fused by
    InteractionAggregation (fusion_heads > 0)."""
    it = input_type or {"user": "image", "photo": "image"}
    two = it["user"] == "both" or it["photo"] == "both"

    def pick(kind, image, ident, which):
        if kind == "both":
            return image if which == 1 else ident
        return image if kind == "image" else ident

    abl = ablation_type if ablation_type in MLP_ABLATIONS else attn_ablation(ablation_type)
    no_pos = "noPos" in (ablation_type or "")
    x1 = backbone(sd, "backbone1.", pick(it["user"], usr_image, usr_id, 1), usr_mask.bool(),
                  pick(it["photo"], vid_image, vid_id, 1), vid_mask.bool(), nhead, num_layers, use_pe, abl, drop, full_usr, no_pos)
    if two:
        x2 = backbone(sd, "backbone2.", pick(it["user"], usr_image, usr_id, 2), usr_mask.bool(),
                      pick(it["photo"], vid_image, vid_id, 2), vid_mask.bool(), nhead, num_layers, use_pe, abl, drop, full_usr, no_pos)
        if fusion_heads > 0:
            logits = fusion_logits(sd, x1, x2, fusion_heads)
        elif fusion_heads == 0:      # models/decoder_leave_focal.py:630-631: stage_mlp1(x1) + stage_mlp2(x2)
            logits = _position_bias(sd, (F.linear(x1, sd["stage_mlp1.weight"], sd["stage_mlp1.bias"]) +
                                         F.linear(x2, sd["stage_mlp2.weight"], sd["stage_mlp2.bias"])).squeeze(-1))
        elif fusion_heads == -1:     # :627-629: Linear(2d, 1) over cat([x1, x2], -1)
            logits = head_logits(sd, torch.cat([x1, x2], dim=-1))
        elif fusion_heads == -2:     # :624-626: Linear(d, 1) over x1 + x2
            logits = head_logits(sd, x1 + x2)
        elif fusion_heads == -3:     # :621-623: `vid_feat_lvls1 + vid_feat_lvls2` concatenates two LISTS; [-1] is backbone2's output
            logits = head_logits(sd, x2)
        else:
            raise NotImplementedError(f"fusion_heads={fusion_heads}")
    else:
        logits = head_logits(sd, x1)
    if mode == "inference":
        return dict(logits=logits, gt=gt)
    exposure_prob = exposure_prob if exposure_prob is not None else [1.0] * logits.shape[1]
    return compute_loss(logits, gt, exposure_prob, loss_type_list, loss_weight, mask_loss=mask_loss)


def live_param_names(sd_keys, num_layers, ablation_type="ours", fusion_heads=2):
    """Names of parameters that receive a gradient in the reference (SURVEY
    section 0 fact 5): everything except layer N-1, the history side of layer
    N-2 (its v2t/t2t projections, ff_usr, ln_usr), and the never-called
    pe_lns / txt_lvl_projs / patch_merge.  Ablations: 'CrossAtt' never uses v2v / t2t; 'SelfAtt' never uses the history
    tokens at all (t2v / v2t / t2t, every *_usr module, usr_proj / usr_pe / usr_ln).  fusion_heads == -3 scores backbone2
    alone (decoder_leave_focal.py:621-623), so nothing of backbone1 is trained."""
    live = []
    N = num_layers
    if ablation_type in MLP_ABLATIONS:      # no attention encoder: embeddings (+ encoder_mlp) + head; SelfMLP / w/oAtt never read the history
        for k in sd_keys:
            if ablation_type != "CrossMLP" and any(s in k for s in ("usr_proj", "usr_pe", "usr_ln")):
                continue
            if ablation_type == "w/oAtt" and "encoder_mlp" in k:
                continue
            live.append(k)
        return live
    abl = attn_ablation(ablation_type)
    for k in sd_keys:
        if any(s in k for s in ("pe_lns", "txt_lvl_projs", "patch_merge")):
            continue
        if fusion_heads == -3 and k.startswith("backbone1."):
            continue
        if abl == "SelfAtt" and any(s in k for s in ("usr_proj", "usr_pe", "usr_ln", "t2v_proj", "v2t_proj", "t2t_proj", "ff_usr", "ln_usr")):
            continue
        if abl == "CrossAtt" and any(s in k for s in ("v2v_proj", "t2t_proj")):
            continue
        if ".encoder.layers." in k:
            i = int(k.split(".encoder.layers.")[1].split(".")[0])
            if i == N - 1:
                continue
            if i == N - 2 and any(s in k for s in ("v2t_proj", "t2t_proj", "ff_usr", "ln_usr")):
                continue
        live.append(k)
    return live


def clip_and_adamw(params, grads, exp_avg, exp_avg_sq, step, lr=1e-3, wd=1e-4, betas=(0.9, 0.999),
                   eps=1e-8, max_norm=None):
    """main_for_seq_leave_earlystop_SegMM.py:298-299: clip_grad_norm_(param_dict, 10.0) then
    torch.optim.AdamW(lr, weight_decay).step().  Lists of tensors, updated in
    place; returns the global gradient norm.

    max_norm=None (default) is what the reference driver actually does: `param_dict = model.parameters()` (:224) is a
    GENERATOR, AdamW's constructor consumes it (:225), so `clip_grad_norm_(param_dict, 10.0)` (:298) walks an exhausted
    generator, returns 0 and clips nothing (torch only warns).  A float turns real clipping on (explicit opt-in)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0) if max_norm else torch.ones((), dtype=grads[0].dtype)
    b1, b2 = betas
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g * coef
        p.mul_(1 - lr * wd)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** step
        bc2 = 1 - b2 ** step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)
    return total


# --------------------------------------------------------------------------
# a-15  parity metric: per-segment skip AUC
# --------------------------------------------------------------------------
def prob_auc_batch(logits, gt, exposure_prob=None):
    """models/my_evaluation.py:73-80 (ProbAUC_batch) fed as main_eval_batch does (:264-284) from
    main_for_seq_leave_earlystop_SegMM.py:402-403: interests = sigmoid(logits) * exposure_prob,
    survival = exp(cumsum(log interests)); positions with gt != -2; label -1 -> 0;
    sklearn.metrics.roc_auc_score on the flattened batch."""
    from sklearn.metrics import roc_auc_score
    logits = torch.as_tensor(logits, dtype=torch.float32)
    gt = torch.as_tensor(gt)
    ep = torch.ones(logits.shape[1]) if exposure_prob is None else torch.as_tensor(exposure_prob, dtype=torch.float32)
    interests = torch.sigmoid(logits) * ep[None, :]
    survival = torch.exp(torch.cumsum(torch.log(interests), dim=1))
    mask = gt != -2
    labels = gt[mask]
    labels = torch.where(labels == -1, torch.zeros_like(labels), labels)
    return float(roc_auc_score(labels.numpy().ravel(), survival[mask].numpy().ravel()))


def eval_rows(interests, gt):
    """Per-row validation metrics of main_eval_batch (models/my_evaluation.py:264-357, test_type 'new'):
    columns pred_view_length (predict_view_length :82-85), view_length, duration, LeaveCTR, LeaveCTR_view (:87-90;
    index view_length - 1 wraps for view_length 0), JaccardSim (IoU_Sim length_aware :37-57)."""
    interests = torch.as_tensor(interests, dtype=torch.float32)
    gt = torch.as_tensor(gt)
    B, L = gt.shape
    survival = torch.exp(torch.cumsum(torch.log(interests), dim=1))
    mask = gt != -2
    view = (gt == 1).sum(1)
    dur = mask.sum(1)
    pred = (survival * mask).sum(1)
    at = (view - 1) % L
    idx = torch.arange(B)
    ctr, ctr_view = 1 - interests[idx, at], 1 - survival[idx, at]
    watched = torch.arange(L)[None, :] < view[:, None]
    inter = ((1 - (gt.to(torch.float32) - survival).abs()) * watched).sum(1)
    jac = (inter + (dur - view).to(torch.float32)) / dur.to(torch.float32)
    return torch.stack([pred, view.float(), dur.float(), ctr, ctr_view, jac], dim=1)
