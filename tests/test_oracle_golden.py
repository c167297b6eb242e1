"""Pins the oracle restatement (oracle/) against fixtures produced by the
unmodified reference (oracle/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gather_oracle, mmi_oracle
from segmminterest_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("name", ["model_small_dh32", "model_small_dh16", "model_small_crossatt", "model_small_selfatt"])
def test_model_small_matches_reference(name):
    z = _load(name)
    cfg = json.loads(str(z["cfg"]))
    abl = cfg.get("ablation_type", "ours")
    sd = {k[3:]: torch.from_numpy(z[k]).clone().requires_grad_(True) for k in z.files if k.startswith("sd/")}
    out = mmi_oracle.forward(sd, torch.from_numpy(z["usr_image"]), torch.from_numpy(z["usr_mask"]),
                             torch.from_numpy(z["vid_image"]), torch.from_numpy(z["vid_mask"]),
                             torch.from_numpy(z["gt_in"]), nhead=cfg["nhead"], num_layers=cfg["num_layers_enc"], ablation_type=abl)
    assert _rel(out["logits"].detach().numpy(), z["logits"]) < 1e-5
    assert abs(out["loss"].item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    assert abs(out["mse"].item() - float(z["mse"])) <= 1e-5 * abs(float(z["mse"]))
    assert abs(out["mse2"].item() - float(z["mse2"])) <= 1e-5 * abs(float(z["mse2"]))
    assert np.array_equal(out["gt"].numpy(), z["gt_out"])
    out["loss"].backward()
    dead = set(json.loads(str(z["dead_params"])))
    live = set(mmi_oracle.live_param_names(list(sd.keys()), cfg["num_layers_enc"], abl))
    named_params = {k for k in sd if ("grad/" + k) in z.files} | dead
    assert live == {k for k in named_params if k not in dead}
    for k in sorted(live):
        g = sd[k].grad
        assert g is not None, k
        # single-block ablations: a key-projection bias shifts every logit of a query row by the same amount, so its
        # gradient is exactly zero in exact arithmetic -- both sides hold ~1e-9 rounding noise there (absolute floor)
        assert np.linalg.norm(g.numpy() - z["grad/" + k]) < 2e-5 * np.linalg.norm(z["grad/" + k]) + 1e-8, k
    for k in dead:
        assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
    inf = mmi_oracle.forward({k: v.detach() for k, v in sd.items()}, torch.from_numpy(z["usr_image"]),
                             torch.from_numpy(z["usr_mask"]), torch.from_numpy(z["vid_image"]),
                             torch.from_numpy(z["vid_mask"]), torch.from_numpy(z["gt_in"]), nhead=cfg["nhead"],
                             num_layers=cfg["num_layers_enc"], mode="inference", ablation_type=abl)
    assert _rel(inf["logits"].numpy(), z["logits_inference"]) < 1e-5


@pytest.mark.parametrize("name", ["model_small_selfmlp", "model_small_crossmlp", "model_small_woatt", "model_small_selfmlp_dropout"])
def test_mlp_ablations_match_reference(name):
    """the MLP ablations of the encoder (SelfMLP / CrossMLP / w/oAtt, encoder.py:392-400,503-511: MLP_Block over
    the candidate tokens, over [history ; candidate] tokens + AdaptiveAvgPool1d(40), or no encoder at all) -- oracle only so
    far: the CUDA path raises NotImplementedError for them.  The `_dropout` fixture is the reference in train() mode; the
    oracle replays its generator stream (dropout after every hidden ReLU of MLP_Block)."""
    z = _load(name)
    cfg = json.loads(str(z["cfg"]))
    abl = cfg["ablation_type"]
    sd = {k[3:]: torch.from_numpy(z[k]).clone().requires_grad_(True) for k in z.files if k.startswith("sd/")}
    kw = dict(nhead=cfg["nhead"], num_layers=cfg["num_layers_enc"], ablation_type=abl)
    args = (torch.from_numpy(z["usr_image"]), torch.from_numpy(z["usr_mask"]), torch.from_numpy(z["vid_image"]),
            torch.from_numpy(z["vid_mask"]), torch.from_numpy(z["gt_in"]))
    if cfg.get("train_seed") is not None:
        torch.manual_seed(cfg["train_seed"])
        out = mmi_oracle.forward(sd, *args, drop=mmi_oracle.torch_dropout(0.1), **kw)
    else:
        out = mmi_oracle.forward(sd, *args, **kw)
    assert _rel(out["logits"].detach().numpy(), z["logits"]) < 1e-5
    assert abs(out["loss"].item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    out["loss"].backward()
    dead = set(json.loads(str(z["dead_params"])))
    live = set(mmi_oracle.live_param_names(list(sd.keys()), cfg["num_layers_enc"], abl))
    assert live == {k for k in sd if k not in dead}
    for k in sorted(live):
        assert np.linalg.norm(sd[k].grad.numpy() - z["grad/" + k]) < 2e-5 * np.linalg.norm(z["grad/" + k]) + 1e-8, k
    for k in dead:
        assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
    inf = mmi_oracle.forward({k: v.detach() for k, v in sd.items()}, *args, mode="inference", **kw)
    assert _rel(inf["logits"].numpy(), z["logits_inference"]) < 1e-5


@pytest.mark.parametrize("name", ["model_small_dropout", "model_small_dropout_crossatt"])
def test_dropout_sites_match_reference_generator_stream(name):
    """train() mode of the UNMODIFIED reference (nn.Dropout(0.1) at every site, torch generator seeded right before the
    forward) against the oracle drawing from the same generator through its dropout hook: identical only if every site
    sits at the same place, sees the same tensor shape and is drawn in the same order as in the reference -- including
    the logits dropout between the -10000 fill and the scale.  This pins the PLACEMENT of the oracle's dropout sites;
    the CUDA path is then compared with the oracle under its own counter-based masks (tests/test_gpu_dropout.py)."""
    z = _load(name)
    cfg = json.loads(str(z["cfg"]))
    abl = cfg.get("ablation_type", "ours")
    sd = {k[3:]: torch.from_numpy(z[k]).clone().requires_grad_(True) for k in z.files if k.startswith("sd/")}
    torch.manual_seed(cfg["train_seed"])
    out = mmi_oracle.forward(sd, torch.from_numpy(z["usr_image"]), torch.from_numpy(z["usr_mask"]),
                             torch.from_numpy(z["vid_image"]), torch.from_numpy(z["vid_mask"]),
                             torch.from_numpy(z["gt_in"]), nhead=cfg["nhead"], num_layers=cfg["num_layers_enc"], ablation_type=abl,
                             drop=mmi_oracle.torch_dropout(0.1), full_usr=True)
    assert _rel(out["logits"].detach().numpy(), z["logits"]) < 1e-5
    assert abs(out["loss"].item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    out["loss"].backward()
    dead = set(json.loads(str(z["dead_params"])))
    n = 0
    for k in sd:
        if ("grad/" + k) not in z.files:
            continue
        assert k not in dead
        g = sd[k].grad
        assert g is not None, k
        assert np.linalg.norm(g.numpy() - z["grad/" + k]) < 2e-5 * np.linalg.norm(z["grad/" + k]) + 1e-8, k
        n += 1
    assert n > 20
    # and the masks matter: the eval-mode forward of the same weights differs visibly
    ev = mmi_oracle.forward({k: v.detach() for k, v in sd.items()}, torch.from_numpy(z["usr_image"]), torch.from_numpy(z["usr_mask"]),
                            torch.from_numpy(z["vid_image"]), torch.from_numpy(z["vid_mask"]), torch.from_numpy(z["gt_in"]),
                            nhead=cfg["nhead"], num_layers=cfg["num_layers_enc"], mode="inference", ablation_type=abl)
    assert _rel(ev["logits"].numpy(), z["logits_inference"]) < 1e-5
    assert _rel(ev["logits"].numpy(), z["logits"]) > 1e-2


@pytest.mark.parametrize("name", ["model_both_small", "model_id_small", "model_both_fh0", "model_both_fh-1", "model_both_fh-2", "model_both_fh-3",
                                  "model_both_bias", "model_image_bias", "model_both_nopos", "model_id_nopos"])
def test_general_config_matches_reference(name):
    """ID-embedding inputs, two backbones + InteractionAggregation (the reference default 'both'),
    interestBPR: the oracle against the unmodified reference."""
    z = _load(name)
    cfg = json.loads(str(z["cfg"]))
    sd = {k[3:]: torch.from_numpy(z[k]).clone() for k in z.files if k.startswith("sd/")}
    for k, v in sd.items():
        if v.is_floating_point():
            v.requires_grad_(True)
    kw = dict(nhead=cfg["nhead"], num_layers=cfg["num_layers_enc"], loss_type_list=tuple(cfg["loss_types"]),
              usr_id=torch.from_numpy(z["usr_id"]), vid_id=torch.from_numpy(z["vid_id"]), input_type=cfg["input_type"],
              fusion_heads=cfg["fusion_heads"], ablation_type=cfg.get("ablation_type", "ours"))
    args = (torch.from_numpy(z["usr_image"]), torch.from_numpy(z["usr_mask"]), torch.from_numpy(z["vid_image"]),
            torch.from_numpy(z["vid_mask"]), torch.from_numpy(z["gt_in"]))
    if cfg.get("draw_seed") is not None:      # 'noPos': replay the reference's torch.randperm stream (training forward, then inference)
        torch.manual_seed(cfg["draw_seed"])
    out = mmi_oracle.forward(sd, *args, **kw)
    assert _rel(out["logits"].detach().numpy(), z["logits"]) < 1e-5
    assert abs(out["loss"].item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    for lt in cfg["loss_types"]:
        assert abs(out[lt].item() - float(z[lt])) <= 1e-5 * abs(float(z[lt]))
    out["loss"].backward()
    dead = set(json.loads(str(z["dead_params"])))
    live = set(mmi_oracle.live_param_names([k for k in sd if ("grad/" + k) in z.files or k in dead], cfg["num_layers_enc"],
                                           fusion_heads=cfg["fusion_heads"]))
    assert live == {k for k in sd if ("grad/" + k) in z.files}
    for k in sorted(live):
        assert sd[k].grad is not None, k
        ref = z["grad/" + k].astype(np.float64)
        # interestBPR is invariant to a common shift of the logits: the head biases get an analytically zero
        # gradient (1e-10 rounding noise in both implementations), hence the absolute term
        assert np.linalg.norm(sd[k].grad.numpy() - ref) < 5e-5 * np.linalg.norm(ref) + 1e-7, k
    for k in dead:
        assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
    inf = mmi_oracle.forward({k: v.detach() for k, v in sd.items()}, *args, mode="inference", **kw)
    assert _rel(inf["logits"].numpy(), z["logits_inference"]) < 1e-5


def test_model_full_matches_reference():
    z = _load("model_full_b4")
    cfg = json.loads(str(z["cfg"]))
    from segmminterest_b200.model import reference_state_shapes
    shapes = reference_state_shapes(d_model=cfg["d_model"], num_layers=cfg["num_layers_enc"], din=cfg["din"],
                                    max_usr_len=cfg["Lt"], max_vid_len=40)
    sd = {k: torch.from_numpy(v).requires_grad_(True) for k, v in synth.fill_state_dict(shapes, cfg["fill_seed"]).items()}
    rng = np.random.default_rng(int(z["data_seed"]))
    usr, usr_mask, vid, vid_mask, gt = synth.make_dense_batch(rng, cfg["B"], cfg["Lt"], cfg["din"])
    assert np.array_equal(gt, z["gt_in"]) and np.array_equal(usr_mask, z["usr_mask"])
    out = mmi_oracle.forward(sd, torch.from_numpy(usr), torch.from_numpy(usr_mask), torch.from_numpy(vid),
                             torch.from_numpy(vid_mask), torch.from_numpy(gt), nhead=cfg["nhead"],
                             num_layers=cfg["num_layers_enc"])
    assert _rel(out["logits"].detach().numpy(), z["logits"]) < 1e-5
    assert abs(out["loss"].item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    out["loss"].backward()
    for k in z.files:
        if k.startswith("gradnorm/"):
            name = k[len("gradnorm/"):]
            g = sd[name].grad.numpy()
            gn = np.sqrt((g.astype(np.float64) ** 2).sum())
            assert abs(gn - float(z[k])) <= 1e-4 * float(z[k]) + 1e-12, name
            assert np.allclose(g.reshape(-1)[:16], z["gradhead/" + name], rtol=1e-3, atol=1e-7 + 1e-4 * np.abs(g).max()), name


def test_loss_cases_match_reference():
    z = _load("loss_cases")
    logits = torch.from_numpy(z["logits"]).double().requires_grad_(True)
    gt = torch.from_numpy(z["gt_in"])
    out = mmi_oracle.compute_loss(logits, gt, list(z["exposure_prob"]), ("focal", "interestBPR"))
    for k in ("focal", "interestBPR", "mse", "mse2", "loss"):
        assert abs(out[k].item() - float(z[k])) <= 2e-6 * abs(float(z[k])), k
    assert np.array_equal(out["gt"].numpy(), z["gt_out"])
    (gf,) = torch.autograd.grad(out["focal"], logits, retain_graph=True)
    (gb,) = torch.autograd.grad(out["interestBPR"], logits)
    assert _rel(gf.numpy(), z["grad_focal"]) < 1e-5
    assert _rel(gb.numpy(), z["grad_bpr"]) < 1e-5


def _loss_variants():
    z = _load("loss_cases_all")
    return z, json.loads(str(z["variants"])), json.loads(str(z["weights"]))


@pytest.mark.parametrize("i", range(10))
def test_every_selectable_loss_matches_reference(i):
    """huber / hazard / surviveCE / interestCE / interestKL (models/decoder_leave_focal.py:528-551), alone and mixed
    with focal (whose in-place gt rewrite later losses see) and interestBPR, against the unmodified reference."""
    z, variants, weights = _loss_variants()
    base = _load("loss_cases")
    tag, lst, mask_loss = variants[i]
    logits = torch.from_numpy(z[f"{tag}/logits"]).requires_grad_(True)
    out = mmi_oracle.compute_loss(logits, torch.from_numpy(base["gt_in"]), list(base["exposure_prob"]), tuple(lst), weights,
                                  mask_loss=mask_loss)
    for k in lst + ["mse", "mse2", "loss"]:
        assert abs(out[k].item() - float(z[f"{tag}/{k}"])) <= 2e-5 * abs(float(z[f"{tag}/{k}"])), (tag, k)
    assert np.array_equal(out["gt"].numpy(), z[f"{tag}/gt_out"])
    (g,) = torch.autograd.grad(out["loss"], logits)
    assert _rel(g.numpy(), z[f"{tag}/grad"]) < 1e-5, tag


def test_eval_metrics_match_reference_main_eval_batch():
    """oracle restatement of main_eval_batch's metrics vs the lists the unmodified reference collected"""
    z = _load("eval_cases")
    base = _load("loss_cases")
    rows = mmi_oracle.eval_rows(torch.from_numpy(z["interests"]), torch.from_numpy(base["gt_in"])).numpy()
    for name, col in (("LeaveMSE", 0), ("view_lengths", 1), ("duration_lengths", 2), ("LeaveCTR", 3), ("LeaveCTR_view", 4), ("JaccardSim", 5)):
        assert np.allclose(rows[:, col], z[name], rtol=1e-5, atol=1e-6, equal_nan=True), name
    auc = mmi_oracle.prob_auc_batch(base["logits"], base["gt_in"], base["exposure_prob"])
    assert abs(auc - float(z["ProbAUC"][0])) < 1e-9


def test_gather_oracle_matches_reference_dataloader():
    z = _load("gather_small")
    table = z["table"]
    lineid = json.loads(str(z["lineid_json"]))
    uin = json.loads(str(z["user_input_json"]))
    rows = json.loads(str(z["rows_json"]))
    for b, (uid, pid, t, dur, play, lab, hi, hp, hl) in enumerate(rows):
        cand = gather_oracle.candidate_rows(pid, dur, lineid)
        photo, pmask = gather_oracle.gather_pad_mask(table, cand, 40)
        assert np.array_equal(photo, z["out/photo"][b]) and np.array_equal(pmask, z["out/photo_mask"][b])
        hist = gather_oracle.history_rows(uid, gather_oracle.parse_int_list(hi) if hl > 0 else [],
                                          gather_oracle.parse_int_list(hp) if hl > 0 else [], lineid, uin)
        assert len(hist) <= 100
        user, umask = gather_oracle.gather_pad_mask(table, hist, 100)
        assert np.array_equal(user, z["out/user"][b]) and np.array_equal(umask, z["out/user_mask"][b])
        assert np.array_equal(gather_oracle.pad_labels(lab), z["out/label"][b])
        assert int(play / 5000) == int(z["out/play_time"][b]) and int(dur / 5000) == int(z["out/duration"][b])
    with pytest.raises(ValueError):
        gather_oracle.candidate_rows(104, 9000, lineid)  # 2 segments wanted, only 1 in the map


def test_gather_dense_and_l1():
    rng = np.random.default_rng(0)
    table = rng.standard_normal((50, 24), dtype=np.float32)
    idx = rng.integers(-1, 50, size=(3, 7)).astype(np.int32)
    out, m = gather_oracle.gather_dense(table, idx)
    assert np.array_equal(m, idx >= 0)
    assert np.array_equal(out[m], table[idx[m]]) and not out[~m].any()
    x = torch.from_numpy(out)
    ref = (x / (x.norm(p=1, dim=-1, keepdim=True) + 1e-6)).numpy()
    assert np.allclose(gather_oracle.l1_normalise(out), ref, rtol=1e-6, atol=0)  # sum order of the norm differs by ulps


def test_clip_adamw_oracle_matches_torch():
    torch.manual_seed(0)
    ps = [torch.randn(7, 5), torch.randn(11)]
    gs = [torch.randn(7, 5) * 30, torch.randn(11) * 30]
    ref_p = [torch.nn.Parameter(p.clone()) for p in ps]
    opt = torch.optim.AdamW(ref_p, lr=1e-3, weight_decay=1e-4)
    mine = [p.clone() for p in ps]
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    for step in (1, 2, 3):
        for p, g in zip(ref_p, gs):
            p.grad = g.clone() * step
        n_ref = torch.nn.utils.clip_grad_norm_(ref_p, 10.0)
        opt.step()
        n = mmi_oracle.clip_and_adamw(mine, [g.clone() * step for g in gs], m, v, step, max_norm=10.0)   # clipping: explicit opt-in
        assert abs(n.item() - n_ref.item()) < 1e-4 * n_ref.item()
        for a, b in zip(mine, ref_p):
            assert torch.allclose(a, b.detach(), rtol=1e-6, atol=1e-7)


def test_reference_driver_clip_is_a_no_op_and_the_oracle_default_follows_it():
    """The driver's own statements (main_for_seq_leave_earlystop_SegMM.py:224-225,298): `param_dict = model.parameters()`
    is a generator, AdamW's constructor consumes it, so `clip_grad_norm_(param_dict, 10.0)` walks nothing, returns 0 and
    leaves gradients of norm >> 10 untouched.  mmi_oracle.clip_and_adamw / TrainStep default to exactly that."""
    import inspect
    import warnings
    torch.manual_seed(1)
    model = torch.nn.Linear(5, 7)
    param_dict = model.parameters()
    opt = torch.optim.AdamW(param_dict, lr=1e-3, weight_decay=1e-4)
    mine = [p.detach().clone() for p in model.parameters()]
    m = [torch.zeros_like(p) for p in mine]
    v = [torch.zeros_like(p) for p in mine]
    for step in (1, 2):
        gs = [torch.randn_like(p) * 50 for p in model.parameters()]
        for p, g in zip(model.parameters(), gs):
            p.grad = g.clone()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert float(torch.nn.utils.clip_grad_norm_(param_dict, 10.0)) == 0.0      # exhausted generator
        for p, g in zip(model.parameters(), gs):
            assert torch.equal(p.grad, g)                                            # nothing was clipped
        opt.step()
        n = mmi_oracle.clip_and_adamw(mine, [g.clone() for g in gs], m, v, step)      # default: no clipping
        assert n.item() > 10.0
        for a, b in zip(mine, model.parameters()):
            assert torch.allclose(a, b.detach(), rtol=1e-6, atol=1e-7)
    from segmminterest_b200.train import TrainStep
    assert inspect.signature(TrainStep.__init__).parameters["max_norm"].default is None
