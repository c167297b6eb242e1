"""Training-mode parity: every nn.Dropout(0.1) site of the reference (models/encoder.py:145-150,163-164,198-202,386,472;
kn_util/nn_utils/layers/mlp.py:21-22) through the CUDA path, against the oracle running the reference's arithmetic with
EXACTLY the masks the kernels generate (oracle/dropout_ref.py is the numpy twin of csrc/dropout.cuh).  The oracle's
dropout placement is pinned to the unmodified reference by tests/test_oracle_golden.py (generator-stream replay).
Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import json
import math
import os

import numpy as np
import pytest
import torch

from types import SimpleNamespace

pytestmark = pytest.mark.gpu


def make_args(**over):
    a = dict(debug=0, input_type={"user": "image", "photo": "image"}, d_model=512, nhead=16, learnable_bias=0,
             exposure_prob=[1.0] * 40, fusion_heads=2, loss_type_list=["focal"],
             loss_weight={"focal": 1.0, "mse": 1.0, "hazard": 1.0, "surviveCE": 1.0, "interestBPR": 1.0,
                          "interestCE": 1.0, "interestKL": 1.0},
             mask_loss=0, num_layers_enc=6, ablation_type="ours", use_pe=1, mmi_precision="fp32")
    a.update(over)
    return SimpleNamespace(**a)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "needs a CUDA device"
    from segmminterest_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


# ----------------------------------------------------------------------------- the generator itself
@pytest.mark.parametrize("key,thr8,row0,rows,cols,group0", [
    (0x12345678, 26, 0, 300, 512, 0), (0xDEADBEEF, 26, (1 << 33) + 12345, 64, 96, 0), (7, 128, 5, 33, 40, 1 << 20),
    (0xFFFFFFFF, 255, 0, 17, 3072, 0), (1, 1, 0, 9, 31, 3)])
def test_mask_kernel_matches_numpy_twin(dev, key, thr8, row0, rows, cols, group0):
    from oracle import dropout_ref
    from segmminterest_b200 import ops
    from segmminterest_b200.dropout import DropSite
    mask = torch.empty(rows, cols, dtype=torch.uint8, device=dev)
    ops.dropout_mask(DropSite(key, thr8, 256.0 / (256 - thr8)), row0, rows, cols, mask, group0=group0)
    want = dropout_ref.keep_mask(key, thr8, np.arange(rows, dtype=np.uint64) + np.uint64(row0), cols, group0=group0)
    assert np.array_equal(mask.cpu().numpy().astype(bool), want)


# ----------------------------------------------------------------------------- element-wise sites
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("d,rows", [(64, 300), (512, 777), (768, 130)])
def test_layernorm_dropout_sites(dev, d, rows, dtype):
    """y = dropout(LN(x)) and its backward (embedding), and the two-output backward of LN(x + dropout(Linear))."""
    from oracle import dropout_ref
    from segmminterest_b200 import _lib, ops
    from segmminterest_b200.dropout import DropSite, quantise
    torch.manual_seed(5)
    thr8, scale = quantise(0.1)
    s1, s2 = DropSite(0xA5A5A5A5, thr8, scale), DropSite(0x0BADF00D, thr8, scale)
    m1 = torch.from_numpy(dropout_ref.keep_mask(s1.key, thr8, np.arange(rows), d)).to(dev)
    m2 = torch.from_numpy(dropout_ref.keep_mask(s2.key, thr8, np.arange(rows), d)).to(dev)
    x = (torch.randn(rows, d, device=dev) * 2 + 0.5).to(dtype)
    g, b = torch.randn(d, device=dev), torch.randn(d, device=dev)
    dy = torch.randn(rows, d, device=dev).to(dtype)
    y, st = torch.empty_like(x), torch.empty(rows, 2, device=dev)
    ops.layernorm_fwd(x, rows, d, g, b, y, st, drop=s1)
    xr = x.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-12) * m1 * scale
    tol = 2e-6 if dtype == torch.float32 else 6e-3
    assert _rel(y.double().cpu(), yr.detach().cpu()) < tol
    assert bool(((y == 0) | m1).all())                   # dropped elements are exactly zero
    yr.backward(dy.double())
    ws = torch.empty(int(_lib.load().mmi_layernorm_bwd_workspace(d)), device=dev)
    dx, dg, db = torch.empty_like(x), torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    ops.layernorm_bwd(dy, x, rows, d, g, st, None, dx, dg, db, ws, dy_drop=s1)
    assert _rel(dx.double().cpu(), xr.grad.cpu()) < tol
    assert _rel(dg.cpu(), gr.grad.cpu()) < tol and _rel(db.cpu(), br.grad.cpu()) < tol
    # residual site: dx stays the residual-branch gradient, dx_dropped = mask * scale * dx, dxsum sums dx_dropped
    dx2, dxm, dxs = torch.empty_like(x), torch.empty_like(x), torch.full((d,), 2.0, device=dev)
    dg2, db2 = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    xr2 = x.double().requires_grad_(True)
    torch.nn.functional.layer_norm(xr2, (d,), g.double(), b.double(), 1e-12).backward(dy.double())
    ops.layernorm_bwd(dy, x, rows, d, g, st, None, dx2, dg2, db2, ws, dxsum=dxs, dx_drop=s2, dx_dropped=dxm)
    assert _rel(dx2.double().cpu(), xr2.grad.cpu()) < tol
    assert _rel(dxm.double().cpu(), (xr2.grad * m2 * scale).cpu()) < tol
    assert _rel((dxs - 2).cpu(), dxm.double().sum(0).cpu()) < 1e-5


@pytest.mark.parametrize("impl,dtype", [("simt", torch.float32), ("tc", torch.bfloat16)])
def test_gemm_epilogue_dropout(dev, impl, dtype):
    """Linear + dropout + residual, and Linear + GELU + dropout with the saved gelu' carrying the mask."""
    from oracle import dropout_ref
    from segmminterest_b200 import ops
    from segmminterest_b200.dropout import DropSite, quantise
    torch.manual_seed(6)
    M, N, K = 700, 512, 256
    thr8, scale = quantise(0.1)
    site = DropSite(0x51F15EED, thr8, scale)
    m = torch.from_numpy(dropout_ref.keep_mask(site.key, thr8, np.arange(M), N)).to(dev)
    A = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
    W = (torch.randn(N, K, device=dev) * 0.1).to(dtype)
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).to(dtype)
    code = ops.IMPL_TC if impl == "tc" else ops.IMPL_SIMT
    tol = 3e-6 if dtype == torch.float32 else 8e-3
    z = A.double() @ W.double().T + bias.double()
    out = torch.empty(M, N, device=dev, dtype=dtype)
    ops.gemm(ops.GEMM_NT, code, A, K, W, K, out, N, M, N, K, bias=bias, add=res, add_mod=M, ld_add=N, drop=site)
    assert _rel(out.double().cpu(), (z * m * scale + res.double()).cpu()) < tol
    out2, pre = torch.empty_like(out), torch.empty_like(out)
    ops.gemm(ops.GEMM_NT, code, A, K, W, K, out2, N, M, N, K, bias=bias, act=ops.ACT_GELU, preact=pre, save_act_grad=True, drop=site)
    zr = z.clone().requires_grad_(True)
    gz = torch.nn.functional.gelu(zr)
    gz.sum().backward()
    assert _rel(out2.double().cpu(), (gz.detach() * m * scale).cpu()) < tol
    assert _rel(pre.double().cpu(), (zr.grad * m * scale).cpu()) < tol


# ----------------------------------------------------------------------------- attention logits dropout
def _ref_attention_drop(qa, ka, va, mka, qb, kb, vb, mkb, mq, H, keep, scale):
    B, Lq, d = qa.shape
    dh = d // H

    def logits(q, k, mk):
        s = torch.einsum("bqhd,bkhd->bhqk", q.view(B, Lq, H, dh), k.view(B, -1, H, dh))
        m = (mq[:, :, None] & mk[:, None, :])[:, None].expand_as(s)
        return torch.where(m, s, torch.full_like(s, -10000.0))

    S = torch.cat([logits(qa, ka, mka), logits(qb, kb, mkb)], -1)
    S = torch.where(keep, S * scale, torch.zeros_like(S)) / math.sqrt(dh)     # models/encoder.py:144-146
    V = torch.cat([va, vb], 1).view(B, -1, H, dh)
    return torch.einsum("bhqk,bkhd->bqhd", S.softmax(-1), V).reshape(B, Lq, d)


@pytest.mark.parametrize("dh,dtype,impl,shape,ragged", [
    (32, torch.float32, "simt", (3, 2, 70, 40, 150), True), (16, torch.float32, "simt", (2, 2, 33, 40, 70), True),
    (32, torch.bfloat16, "tc", (3, 2, 70, 40, 150), True), (32, torch.bfloat16, "tc", (2, 4, 500, 40, 500), True),
    (32, torch.bfloat16, "tc", (2, 4, 500, 40, 500), False), (32, torch.bfloat16, "tc", (2, 4, 40, 40, 500), True),
    (32, torch.bfloat16, "tc", (1, 4, 129, 64, 65), True), (32, torch.bfloat16, "tc", (1, 2, 300, 33, 1030), False)])
def test_attention_logits_dropout(dev, dh, dtype, impl, shape, ragged):
    from oracle import dropout_ref
    from segmminterest_b200 import ops
    from segmminterest_b200.dropout import DropSite, quantise
    torch.manual_seed(4)
    B, H, Lq, La, Lb = shape
    d = H * dh
    code = ops.IMPL_TC if impl == "tc" else ops.IMPL_SIMT
    thr8, scale = quantise(0.1)
    site = DropSite(0xC0FFEE11, thr8, scale)
    keep = torch.from_numpy(dropout_ref.attn_keep_mask(site.key, thr8, B, H, Lq, [La, Lb]))

    def mk(L):
        n = torch.randint(1, L + 1, (B,)) if ragged else torch.full((B,), L)
        return (torch.arange(L)[None] < n[:, None])

    mq, mka, mkb = mk(Lq), mk(La), mk(Lb)
    mq[0, :] = True
    t = [(torch.randn(B, L, d) * 0.7).to(dtype).to(dev) for L in (Lq, La, La, Lq, Lb, Lb)]
    qa, ka, va, qb, kb, vb = t
    ref_in = [x.double().cpu().requires_grad_(True) for x in t]
    ref = _ref_attention_drop(ref_in[0], ref_in[1], ref_in[2], mka, ref_in[3], ref_in[4], ref_in[5], mkb, mq, H, keep, scale)
    out = torch.empty(B * Lq, d, device=dev, dtype=dtype)
    lse = torch.empty(B, H, Lq, device=dev)
    mqd, mkad, mkbd = [m.to(dev).view(torch.uint8) for m in (mq, mka, mkb)]
    blocks = [dict(q=(qa.data_ptr(), d), k=(ka.data_ptr(), d), v=(va.data_ptr(), d), mask_k=mkad, Lk=La),
              dict(q=(qb.data_ptr(), d), k=(kb.data_ptr(), d), v=(vb.data_ptr(), d), mask_k=mkbd, Lk=Lb)]
    side = ops.AttnSide(ops.dt(out), code, B, H, dh, Lq, mqd, out, d, lse, blocks, drop=site)
    side.fwd()
    tol = 3e-6 if dtype == torch.float32 else 8e-3
    assert _rel(out.view(B, Lq, d).double().cpu(), ref.detach()) < tol
    dO = (torch.randn(B, Lq, d) * 0.5).to(dtype).to(dev)
    ref.backward(dO.double().cpu())
    grads = [torch.zeros_like(x) for x in t]
    delta = torch.empty(B, H, Lq, device=dev)
    side.set_bwd(dO, d, delta, [dict(dq=(grads[0].data_ptr(), d), dk=(grads[1].data_ptr(), d), dv=(grads[2].data_ptr(), d)),
                                dict(dq=(grads[3].data_ptr(), d), dk=(grads[4].data_ptr(), d), dv=(grads[5].data_ptr(), d))])
    side.bwd_dq()
    side.bwd_dkv(0)
    side.bwd_dkv(1)
    for g, r, name in zip(grads, ref_in, ["dqa", "dka", "dva", "dqb", "dkb", "dvb"]):
        assert _rel(g.double().cpu(), r.grad) < (5e-5 if dtype == torch.float32 else 1.5e-2), name
    if impl == "tc":                     # the one-kernel backward regenerates the same masks (thread = key, 64-query tiles)
        for g in grads:
            g.zero_()
        acc = [torch.zeros(B * Lq, d, device=dev) for _ in range(2)]
        cnt = [torch.zeros(B * H, device=dev, dtype=torch.int32) for _ in range(2)]
        side.set_fused(acc, cnt)
        assert side.bwd_fused(0) and side.bwd_fused(1)
        for g, r, name in zip(grads, ref_in, ["dqa", "dka", "dva", "dqb", "dkb", "dvb"]):
            assert _rel(g.double().cpu(), r.grad) < 1.5e-2, ("fused", name)
        assert all(float(a_.abs().max()) == 0.0 for a_ in acc) and all(int(c_.abs().max()) == 0 for c_ in cnt)
        for g in grads:
            g.zero_()
        if side.bwd_all():               # one CTA per (b, h): keep words of two key tiles per bit transpose
            for g, r, name in zip(grads, ref_in, ["dqa", "dka", "dva", "dqb", "dkb", "dvb"]):
                assert _rel(g.double().cpu(), r.grad) < 1.5e-2, ("all-keys", name)


# ----------------------------------------------------------------------------- the whole model in train() mode
def _oracle_drop(eng, H):
    """the oracle's dropout hook, fed with the masks of the forward call the engine just ran"""
    from oracle import dropout_ref
    from segmminterest_b200.dropout import quantise, site_key
    thr8, scale = quantise(eng.drop_p)
    call, seed = eng._drop_call, eng.drop_seed

    def drop(kind, tower, layer, side, x, blocks=None):
        key = site_key(seed, call, ((tower * 64 + layer) * 2 + side) * 8 + kind)
        if blocks is not None:
            B, H_, Lq, _ = x.shape
            keep = dropout_ref.attn_keep_mask(key, thr8, B, H_, Lq, blocks)
        else:
            keep = dropout_ref.keep_mask(key, thr8, np.arange(x.shape[0] * x.shape[1]), x.shape[2]).reshape(x.shape)
        keep = torch.from_numpy(keep)
        return torch.where(keep, x * scale, torch.zeros_like(x))
    return drop


@pytest.mark.parametrize("name", ["model_small_dh32", "model_small_dh16", "model_small_crossatt", "model_small_selfatt",
                                  "model_small_selfmlp", "model_small_crossmlp"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_train_mode_model_vs_oracle_with_the_same_masks(dev, name, precision):
    from oracle import mmi_oracle
    from segmminterest_b200.model import build_model
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(str(z["cfg"]))
    abl = cfg.get("ablation_type", "ours")
    args = make_args(d_model=cfg["d_model"], nhead=cfg["nhead"], num_layers_enc=cfg["num_layers_enc"], mmi_precision=precision,
                     ablation_type=abl)
    torch.manual_seed(77)
    model = build_model(args, din=cfg["din"], max_usr_len=cfg["Lt"])
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    model.load_state_dict(sd)
    model = model.cuda()
    model.train()                                        # nn.Dropout(0.1) semantics at every site
    B = z["usr_image"].shape[0]

    def run():
        return model(usr_image=torch.from_numpy(z["usr_image"]).to(dev), usr_id=torch.zeros(B, dtype=torch.long, device=dev),
                     usr_mask=torch.from_numpy(z["usr_mask"]).to(dev), vid_image=torch.from_numpy(z["vid_image"]).to(dev),
                     vid_id=torch.zeros(B, dtype=torch.long, device=dev), vid_mask=torch.from_numpy(z["vid_mask"]).to(dev),
                     gt=torch.from_numpy(z["gt_in"].copy()).to(dev), mode="train")

    first = run()["logits"].detach().cpu().numpy()
    out = run()                                          # second call: fresh masks (the call counter is part of every key)
    assert _rel(first, out["logits"].detach().cpu().numpy()) > 1e-3
    eng = model.engine()
    osd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    o = mmi_oracle.forward(osd, torch.from_numpy(z["usr_image"]), torch.from_numpy(z["usr_mask"]), torch.from_numpy(z["vid_image"]),
                           torch.from_numpy(z["vid_mask"]), torch.from_numpy(z["gt_in"].copy()), nhead=cfg["nhead"],
                           num_layers=cfg["num_layers_enc"], ablation_type=abl, drop=_oracle_drop(eng, cfg["nhead"]))
    tol = 1e-4 if precision == "fp32" else 2e-2
    assert _rel(out["logits"].detach().cpu().numpy(), o["logits"].detach().numpy()) < tol
    assert abs(out["loss"].item() - o["loss"].item()) < tol * abs(o["loss"].item())
    assert _rel(o["logits"].detach().numpy(), z["logits"]) > 1e-2          # and it is NOT the eval-mode result
    out["loss"].backward()
    o["loss"].backward()
    live = set(mmi_oracle.live_param_names(list(sd.keys()), cfg["num_layers_enc"], abl))
    gmax = max(float(osd[k].grad.norm()) for k in live)
    # same criteria as tests/test_gpu_model.py::test_small_model_vs_reference_golden: whole gradient at the north-star bar,
    # bf16 single tensors at max(4 x bar, 1.5 x amp) with amp = error of the head-bias gradient (the fp32 sum of dlogits:
    # how much this fixture's loss gradient amplifies the forward's rounding error), plus a noise floor
    floor = 0.0 if precision == "fp32" else 2e-5 * gmax
    amp = 0.0
    if precision == "bf16":
        hb = dict(model.named_parameters())["stage_mlp1.bias"].grad.double().cpu().numpy()
        hr = osd["stage_mlp1.bias"].grad.numpy()
        amp = float(np.linalg.norm(hb - hr) / np.linalg.norm(hr))
    per_tensor = tol if precision == "fp32" else max(4 * tol, 1.5 * amp)
    whole = tol if precision == "fp32" else max(tol, 0.4 * amp)
    assert amp < 0.15, amp
    err2 = ref2 = 0.0
    for k, p in model.named_parameters():
        if k not in live:
            assert p.grad is None, k
            continue
        ref = osd[k].grad.numpy()
        if np.linalg.norm(ref) < 1e-7:
            assert float(p.grad.abs().max()) < max(1e-5, floor), k
            continue
        err = float(np.linalg.norm(p.grad.double().cpu().numpy() - ref))
        err2 += err * err
        ref2 += float(np.linalg.norm(ref)) ** 2
        assert err < per_tensor * float(np.linalg.norm(ref)) + floor, (k, err, float(np.linalg.norm(ref)), amp)
    assert err2 ** 0.5 < whole * ref2 ** 0.5, ("whole gradient", err2 ** 0.5, ref2 ** 0.5, amp)
    # eval() switches every site off again: the fixture's eval-mode logits come back
    model.eval()
    ev = run()
    assert _rel(ev["logits"].detach().cpu().numpy(), z["logits"]) < tol


@pytest.mark.parametrize("name", ["model_both_small", "model_id_small"])
def test_train_mode_two_towers_and_id_inputs_vs_oracle(dev, name):
    """train() mode of the reference's default multi-modal configuration (image tower + ID tower fused by
    InteractionAggregation) and of the ID-only configuration: every dropout site of BOTH towers (the ID tower has a
    one-token history) against the oracle under the kernels' masks, strict fp32 bar."""
    from oracle import mmi_oracle
    from segmminterest_b200.model import build_model
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(str(z["cfg"]))
    args = make_args(d_model=cfg["d_model"], nhead=cfg["nhead"], num_layers_enc=cfg["num_layers_enc"], input_type=cfg["input_type"],
                     fusion_heads=cfg["fusion_heads"], loss_type_list=list(cfg["loss_types"]), mmi_precision="fp32")
    torch.manual_seed(5)
    model = build_model(args, din=cfg["din"], max_usr_len=100, n_users=cfg["n_users"], n_items=cfg["n_items"])
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    model.load_state_dict(sd)
    model = model.cuda().train()
    out = model(usr_image=torch.from_numpy(z["usr_image"]).to(dev), usr_id=torch.from_numpy(z["usr_id"]).to(dev),
                usr_mask=torch.from_numpy(z["usr_mask"]).to(dev), vid_image=torch.from_numpy(z["vid_image"]).to(dev),
                vid_id=torch.from_numpy(z["vid_id"]).to(dev), vid_mask=torch.from_numpy(z["vid_mask"]).to(dev),
                gt=torch.from_numpy(z["gt_in"].copy()).to(dev), mode="train")
    osd = {k: v.clone() for k, v in sd.items()}
    for v in osd.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    o = mmi_oracle.forward(osd, torch.from_numpy(z["usr_image"]), torch.from_numpy(z["usr_mask"]), torch.from_numpy(z["vid_image"]),
                           torch.from_numpy(z["vid_mask"]), torch.from_numpy(z["gt_in"].copy()), nhead=cfg["nhead"],
                           num_layers=cfg["num_layers_enc"], loss_type_list=tuple(cfg["loss_types"]), usr_id=torch.from_numpy(z["usr_id"]),
                           vid_id=torch.from_numpy(z["vid_id"]), input_type=cfg["input_type"], fusion_heads=cfg["fusion_heads"],
                           drop=_oracle_drop(model.engine(), cfg["nhead"]))
    assert _rel(out["logits"].detach().cpu().numpy(), o["logits"].detach().numpy()) < 1e-4
    assert _rel(o["logits"].detach().numpy(), z["logits"]) > 1e-3                            # not the eval-mode result
    assert abs(out["loss"].item() - o["loss"].item()) < 1e-4 * abs(o["loss"].item()) + 2e-6
    out["loss"].backward()
    o["loss"].backward()
    dead = set(json.loads(str(z["dead_params"])))
    n = 0
    for k, p in model.named_parameters():
        if k in dead:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        ref = osd[k].grad.numpy().astype(np.float64)
        assert np.linalg.norm(p.grad.double().cpu().numpy() - ref) < 3e-4 * np.linalg.norm(ref) + 1e-7, k
        n += 1
    assert n > 40


def test_train_step_dropout_follows_module_mode(dev):
    """TrainStep: dropout on under train(), off under eval(), overridable; masks differ from step to step."""
    from segmminterest_b200 import synth
    from segmminterest_b200.model import build_model
    from segmminterest_b200.train import TrainStep
    torch.manual_seed(3)
    B, Lt, din = 8, 24, 64
    model = build_model(make_args(d_model=64, nhead=2, num_layers_enc=3, mmi_precision="bf16"), din=din, max_usr_len=Lt).cuda()
    rng = np.random.default_rng(2)
    table = torch.from_numpy(rng.standard_normal((512, din)).astype(np.float32)).to(dev)
    usr_idx = torch.from_numpy(rng.integers(0, 512, (B, Lt)).astype(np.int32)).to(dev)
    vid_idx = torch.from_numpy(rng.integers(0, 512, (B, 40)).astype(np.int32)).to(dev)
    _, _, _, _, gt = synth.make_dense_batch(rng, B, Lt, din)
    losses = {}
    for mode, override in (("eval", None), ("train", None), ("train0", 0.0)):
        model.train(mode != "eval")
        ts = TrainStep(model, table, lr=0.0, weight_decay=0.0, global_batch=B, dropout=override)
        losses[mode] = [float(ts.step(usr_idx, vid_idx, torch.from_numpy(gt.copy()).to(dev))[3].item()) for _ in range(3)]
    assert len(set(losses["eval"])) == 1 and losses["eval"] == losses["train0"]          # lr = 0: deterministic without dropout
    assert len(set(losses["train"])) == 3 and all(abs(a - losses["eval"][0]) > 1e-6 for a in losses["train"])
