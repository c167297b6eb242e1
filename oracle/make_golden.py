"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, via oracle/ref_shim.py).  Build-container only; the fixtures
are committed so the GPU box never needs the reference tree.

    python -m oracle.make_golden

torch 2.11.0+cu128 (CPU), numpy 2.3, eval() mode (dropout off) -- see SURVEY
section 8c for why bit-level comparison is only meaningful with dropout off.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gather_oracle, ref_shim  # noqa: E402
from segmminterest_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run_model_case(name, d_model, nhead, nlayers, din, Lt, B, seed, store_sd=True, loss_types=("focal",),
                   fill_seed=None, ablation_type="ours", train_seed=None):
    """train_seed: run the training forward in train() mode (nn.Dropout(0.1) live at every site) right after
    torch.manual_seed(train_seed); the oracle replays the same generator stream (tests/test_oracle_golden.py)."""
    args = ref_shim.make_args(d_model=d_model, nhead=nhead, num_layers_enc=nlayers,
                              loss_type_list=list(loss_types), ablation_type=ablation_type)
    model = ref_shim.build_reference_model(args, din=din, max_usr_len=Lt, seed=seed)
    if fill_seed is not None:
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        sd = {k: torch.from_numpy(v) for k, v in synth.fill_state_dict(shapes, fill_seed).items()}
        model.load_state_dict(sd)
    else:
        # the reference initialises biases / LN to 0 / 1; perturb so those paths are exercised
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for k, p in model.named_parameters():
                if p.ndim == 1:
                    p.add_(torch.randn(p.shape, generator=g) * 0.05)
                if k == "stage_mlp1.weight":
                    p.mul_(3.0)
    model.eval()
    rng = np.random.default_rng(seed + 7)
    usr, usr_mask, vid, vid_mask, gt = synth.make_dense_batch(rng, B, Lt, din)
    batch = dict(usr_image=torch.from_numpy(usr), usr_id=torch.zeros(B, dtype=torch.long),
                 usr_mask=torch.from_numpy(usr_mask), vid_image=torch.from_numpy(vid),
                 vid_id=torch.zeros(B, dtype=torch.long), vid_mask=torch.from_numpy(vid_mask),
                 gt=torch.from_numpy(gt.copy()))
    if train_seed is not None:
        model.train()
        torch.manual_seed(train_seed)
    out = ref_shim.run_reference(model, batch, mode="train")
    out["loss"].backward()
    model.eval()
    save = dict(cfg=json.dumps(dict(d_model=d_model, nhead=nhead, num_layers_enc=nlayers, din=din, Lt=Lt, B=B,
                                    seed=seed, loss_types=list(loss_types), fill_seed=fill_seed, ablation_type=ablation_type,
                                    train_seed=train_seed)),
                usr_mask=usr_mask, vid_mask=vid_mask, gt_in=gt,
                logits=out["logits"].detach().numpy(), gt_out=out["gt"].numpy(),
                loss=np.float64(out["loss"].item()), mse=np.float64(out["mse"].item()),
                mse2=np.float64(out["mse2"].item()))
    for lt_ in loss_types:
        save[lt_] = np.float64(out[lt_].item())
    with torch.no_grad():
        batch["gt"] = torch.from_numpy(gt.copy())
        inf = ref_shim.run_reference(model, batch, mode="inference")
    save["logits_inference"] = inf["logits"].numpy()
    dead = []
    if store_sd:
        save["usr_image"] = usr
        save["vid_image"] = vid
        for k, v in model.state_dict().items():
            save["sd/" + k] = v.numpy()
        for k, p in model.named_parameters():
            if p.grad is None:
                dead.append(k)
            else:
                save["grad/" + k] = p.grad.numpy()
    else:
        save["data_seed"] = np.int64(seed + 7)
        for k, p in model.named_parameters():
            if p.grad is None:
                dead.append(k)
            else:
                g = p.grad.numpy()
                save["gradnorm/" + k] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
                save["gradhead/" + k] = g.reshape(-1)[:16].copy()
    save["dead_params"] = json.dumps(dead)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "loss", save["loss"], "n_dead", len(dead))


@contextlib.contextmanager
def _cuda_is_identity():
    """The position bias hard-codes `.cuda()` on a fresh arange (decoder_leave_focal.py:498,651).  On this CPU-only container
    the unmodified reference runs with Tensor.cuda patched to the identity -- the arithmetic is untouched."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def run_general_case(name, input_type, d_model, nhead, nlayers, din, B, seed, loss_types, n_users=23, n_items=57, fusion_heads=2,
                     learnable_bias=0, ablation_type="ours", draw_seed=None):
    """draw_seed: torch.manual_seed(draw_seed) right before the training forward -- the 'noPos' ablation draws one
    torch.randperm(40) per interaction and ID tower per forward call (encoder.py:428-429); the test replays the stream."""
    """SURVEY 8f-1: ID inputs / two backbones + InteractionAggregation (the reference's default 'both' config), history
    padded to the reference's 100 tokens.  Stores the full state_dict, inputs, outputs and gradients."""
    args = ref_shim.make_args(d_model=d_model, nhead=nhead, num_layers_enc=nlayers, loss_type_list=list(loss_types),
                              input_type=input_type, fusion_heads=fusion_heads, learnable_bias=learnable_bias,
                              ablation_type=ablation_type)
    model = ref_shim.build_reference_model_general(args, din=din, n_users=n_users, n_items=n_items, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if k in ("bias_weight", "bias_bias"):      # [1, 40], initialised to ones: make every position's pair distinct
                p.copy_(torch.randn(p.shape, generator=g) * (0.02 if k == "bias_weight" else 0.3))
                continue
            if p.ndim == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
            if k in ("stage_mlp1.weight", "stage_mlp2.weight"):
                p.mul_(3.0)
            # the fusion head keeps its init scale: interestBPR = -log(A) must stay away from A ~ 1, where the loss
            # amplifies any logit noise by 1 / (1 - A) and a bf16 comparison would only measure that amplification
    model.eval()
    rng = np.random.default_rng(seed + 7)
    Lt = 100
    usr, usr_mask, vid, vid_mask, gt = synth.make_dense_batch(rng, B, Lt, din)
    usr_id = rng.integers(0, n_users + 1, size=B)
    vid_id = rng.integers(0, n_items + 1, size=B)
    vid_id[1] = vid_id[0]          # two interactions with the same video: embedding gradients must add up
    batch = dict(usr_image=torch.from_numpy(usr), usr_id=torch.from_numpy(usr_id), usr_mask=torch.from_numpy(usr_mask),
                 vid_image=torch.from_numpy(vid), vid_id=torch.from_numpy(vid_id), vid_mask=torch.from_numpy(vid_mask),
                 gt=torch.from_numpy(gt.copy()))
    if draw_seed is not None:
        torch.manual_seed(draw_seed)
    with _cuda_is_identity():
        out = ref_shim.run_reference(model, batch, mode="train")
    out["loss"].backward()
    save = dict(cfg=json.dumps(dict(d_model=d_model, nhead=nhead, num_layers_enc=nlayers, din=din, Lt=Lt, B=B, seed=seed,
                                    loss_types=list(loss_types), input_type=input_type, n_users=n_users, n_items=n_items,
                                    fusion_heads=fusion_heads, learnable_bias=learnable_bias, ablation_type=ablation_type,
                                    draw_seed=draw_seed)),
                usr_image=usr, vid_image=vid, usr_id=usr_id, vid_id=vid_id, usr_mask=usr_mask, vid_mask=vid_mask, gt_in=gt,
                logits=out["logits"].detach().numpy(), gt_out=out["gt"].numpy(), loss=np.float64(out["loss"].item()),
                mse=np.float64(out["mse"].item()), mse2=np.float64(out["mse2"].item()))
    for lt_ in loss_types:
        save[lt_] = np.float64(out[lt_].item())
    with torch.no_grad(), _cuda_is_identity():
        batch["gt"] = torch.from_numpy(gt.copy())
        save["logits_inference"] = ref_shim.run_reference(model, batch, mode="inference")["logits"].numpy()
    dead = []
    for k, v in model.state_dict().items():
        save["sd/" + k] = v.numpy()
    for k, p in model.named_parameters():
        if p.grad is None:
            dead.append(k)
        else:
            save["grad/" + k] = p.grad.numpy()
    save["dead_params"] = json.dumps(dead)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "loss", save["loss"], "n_dead", len(dead), "n_params", sum(p.numel() for p in model.parameters()))


def run_loss_cases():
    """compute_loss alone (decoder_leave_focal.py:490-572) on hand-made edge cases."""
    args = ref_shim.make_args(d_model=32, nhead=2, num_layers_enc=2, loss_type_list=["focal", "interestBPR"])
    model = ref_shim.build_reference_model(args, din=8, max_usr_len=4)
    rng = np.random.default_rng(99)
    B = 24
    nv = rng.integers(1, 41, size=B)
    nv[:3] = [40, 1, 2]
    gt = synth.make_labels(rng, nv)
    gt[3] = 1  # fully watched 40-segment video: view_len == 40 row (excluded from BPR)
    gt[4, :] = -2
    gt[4, :5] = 1  # fully watched short video -> BPR positive is the first pad position
    logits = (rng.standard_normal((B, 40)) * 3).astype(np.float32)
    logits[5, :4] = [30.0, -30.0, 60.0, -60.0]  # saturating sigmoids
    lt = torch.from_numpy(logits).requires_grad_(True)
    ep = list(np.linspace(1.0, 0.6, 40).astype(np.float64))
    model.exposure_prob = ep
    with contextlib.redirect_stdout(io.StringIO()):
        out = model.compute_loss(stage_logits=lt[..., None], gt=torch.from_numpy(gt.copy()))
    (g_focal,) = torch.autograd.grad(out["focal"], lt, retain_graph=True)
    (g_bpr,) = torch.autograd.grad(out["interestBPR"], lt, retain_graph=True)
    np.savez_compressed(os.path.join(OUT, "loss_cases.npz"), logits=logits, gt_in=gt,
                        exposure_prob=np.array(ep), gt_out=out["gt"].numpy(),
                        focal=np.float64(out["focal"].item()), interestBPR=np.float64(out["interestBPR"].item()),
                        mse=np.float64(out["mse"].item()), mse2=np.float64(out["mse2"].item()),
                        loss=np.float64(out["loss"].item()), grad_focal=g_focal.numpy(), grad_bpr=g_bpr.numpy())
    print("loss_cases focal", out["focal"].item(), "bpr", out["interestBPR"].item())


ALL_LOSS_VARIANTS = [  # (tag, loss_type_list, mask_loss): order matters -- focal rewrites gt in place before later losses see it
    ("huber", ["huber"], 0), ("hazard", ["hazard"], 0), ("surviveCE", ["surviveCE"], 0),
    ("interestCE", ["interestCE"], 0), ("interestCE_m", ["interestCE"], 1),
    ("interestKL", ["interestKL"], 0), ("interestKL_m", ["interestKL"], 1),
    ("focal_then_CE_KL", ["focal", "interestCE", "interestKL"], 0),
    ("CE_then_focal_KL_m", ["interestCE", "focal", "interestKL"], 1),
    ("all", ["interestBPR", "huber", "hazard", "surviveCE", "focal", "interestCE", "interestKL"], 0),
]


def run_loss_cases_all():
    """Every other selectable loss of compute_loss (decoder_leave_focal.py:528-551: huber, hazard, surviveCE,
    interestCE, interestKL, each alone and mixed with focal / interestBPR) on the inputs of loss_cases.npz."""
    base = np.load(os.path.join(OUT, "loss_cases.npz"))
    logits, gt, ep = base["logits"], base["gt_in"], list(base["exposure_prob"])
    save = {}
    weights = {"focal": 0.7, "mse": 0.3, "hazard": 0.9, "surviveCE": 1.1, "interestBPR": 1.3, "interestCE": 0.8, "interestKL": 1.7}
    for tag, lst, mask_loss in ALL_LOSS_VARIANTS:
        args = ref_shim.make_args(d_model=32, nhead=2, num_layers_enc=2, loss_type_list=list(lst), mask_loss=mask_loss,
                                  loss_weight=dict(weights))
        model = ref_shim.build_reference_model(args, din=8, max_usr_len=4)
        model.exposure_prob = ep
        lg = logits.copy()
        if any(n.startswith("interestCE") or n.startswith("interestKL") for n in lst):
            # the reference takes log(softmax(x)) literally (:102,112): the +-60 logits of row 5 underflow it to -inf
            lg[5, :4] = [8.0, -8.0, 12.0, -12.0]
        lt = torch.from_numpy(lg).requires_grad_(True)
        with contextlib.redirect_stdout(io.StringIO()):
            out = model.compute_loss(stage_logits=lt[..., None], gt=torch.from_numpy(gt.copy()))
        (g,) = torch.autograd.grad(out["loss"], lt, retain_graph=True)
        save[f"{tag}/logits"] = lg
        save[f"{tag}/loss"] = np.float64(out["loss"].item())
        save[f"{tag}/grad"] = g.numpy()
        save[f"{tag}/gt_out"] = out["gt"].numpy()
        for name in lst + ["mse", "mse2"]:
            save[f"{tag}/{name}"] = np.float64(out[name].item())
        print("loss_cases_all", tag, {n: round(float(out[n].item()), 6) for n in lst}, "loss", out["loss"].item())
    save["variants"] = json.dumps(ALL_LOSS_VARIANTS)
    save["weights"] = json.dumps(weights)
    np.savez_compressed(os.path.join(OUT, "loss_cases_all.npz"), **save)


def run_eval_case():
    """main_eval_batch of models/my_evaluation.py (:264-357) on the loss_cases inputs: ProbAUC of the batch and the
    per-row JaccardSim / LeaveMSE / LeaveCTR / LeaveCTR_view lists, exactly as the driver's validation loop collects them
    (main...SegMM.py:396-432)."""
    ev = ref_shim.load_evaluation()
    base = np.load(os.path.join(OUT, "loss_cases.npz"))
    logits = torch.from_numpy(base["logits"])
    gt = torch.from_numpy(base["gt_in"])
    ep = torch.tensor(base["exposure_prob"], dtype=torch.float32)
    interests = torch.sigmoid(logits) * ep                     # main...SegMM.py:402-403
    results = {k: [] for k in ("ProbAUC", "JaccardSim", "LeaveMSE", "view_lengths", "duration_lengths", "LeaveCTR", "LeaveCTR_view")}
    args = ref_shim.make_args(TOP_K_mask=0, TOP_K_permutation=1, draw_case=0)
    with contextlib.redirect_stdout(io.StringIO()):
        results = ev.main_eval_batch(args, interests, gt, torch.zeros_like(gt), results, type="inference")
    save = {k: np.asarray(v, dtype=np.float64) for k, v in results.items()}
    save["interests"] = interests.numpy()
    # test_type 'old' (:270-271: the interests ARE the survival probabilities) and the `logits=` MAES bookkeeping (:309-320)
    old = {k: [] for k in ("ProbAUC", "JaccardSim", "LeaveMSE", "view_lengths", "LeaveCTR", "LeaveCTR_view")}
    old["MAES"] = 0.0
    old["pred_leave"] = []
    with contextlib.redirect_stdout(io.StringIO()):
        old = ev.main_eval_batch(args, interests, gt, torch.zeros_like(gt), old, type="inference", test_type="old", logits=logits * 0.125)   # scaled: the +-60 logits of row 5 would turn 1 / softmax into inf
    for k, v in old.items():
        if k == "pred_leave":
            save["old/pred_leave"] = torch.stack(v).numpy()
        else:
            save["old/" + k] = np.asarray(v, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "eval_cases.npz"), **save)
    print("eval_cases", {k: (len(v), float(np.mean(v))) for k, v in results.items()})


def run_topk_case():
    """TOP_K_leave / TOP_K_leave_mask of models/my_evaluation.py (:137-231) as the driver's validation loop calls them
    (main...SegMM.py:155-167) on the loss_cases inputs, with np.random seeded so the per-row permutations can be replayed."""
    ev = ref_shim.load_evaluation()
    base = np.load(os.path.join(OUT, "loss_cases.npz"))
    logits = torch.from_numpy(base["logits"])
    gt = torch.from_numpy(base["gt_in"])
    ep = torch.tensor(base["exposure_prob"], dtype=torch.float32)
    interests = (torch.sigmoid(logits) * ep).numpy()
    interests[1, 3] = interests[1, 7]                          # ties: only the permutation decides their order
    interests[2, :5] = interests[2, 5]
    view_lengths = (gt == 1).sum(dim=1, keepdim=True).numpy()
    mask_batch = (gt != -2).numpy()
    save = dict(interests=interests, view_lengths=view_lengths, mask_batch=mask_batch, seed=np.int64(2024))
    for fn in ("TOP_K_leave", "TOP_K_leave_mask"):
        for perm in (1, 0):
            np.random.seed(2024)
            with contextlib.redirect_stdout(io.StringIO()):
                res = getattr(ev, fn)(interests.copy(), view_lengths.copy(), mask_batch.copy(), permutation=perm)
            for k, v in res.items():
                save[f"{fn}/{perm}/{k}"] = np.float64(v)
    np.random.seed(2024)
    with contextlib.redirect_stdout(io.StringIO()):
        _, mins = ev.TOP_K_leave(interests.copy(), view_lengths.copy(), mask_batch.copy(), permutation=1, test=1)
    save["TOP_K_leave/min_indices"] = np.asarray(mins)
    np.savez_compressed(os.path.join(OUT, "topk_cases.npz"), **save)
    print("topk_cases", {k: float(v) for k, v in save.items() if k.startswith("TOP_K_leave/1/")})


def run_gather_case():
    """FrameDatasetSeq_SegMM + DataCollator (utils/dataloader_SegMM.py:186-382) on a
    5-video fixture; stores the inputs in index form plus the dense outputs."""
    import pandas as pd
    dl = ref_shim.load_dataloader()
    rng = np.random.default_rng(5)
    din = 1024
    # videos 101..105 with 3,2,40,1,6 segments
    nseg = {101: 3, 102: 2, 103: 40, 104: 1, 105: 6}
    lineid, r = {}, 0
    for pid, n in nseg.items():
        for i in range(n):
            lineid[f"{pid}-{i}"] = r
            r += 1
    lineid["105-25"] = 5  # reachable only through user_input_dict ("pid_sec")
    table = rng.standard_normal((r, din), dtype=np.float32)
    rows = [
        # user, video, time, duration_ms, playing, label, hist items, hist playing, hist len
        (7, 101, 1000, 14999, 6000, "[ 1 0 -1]", "[102 103]", "[ 9000 12000]", 2),
        (7, 103, 2000, 200000, 200000, "[" + " ".join(["1"] * 40) + "]", "[101 102 105]", "[15000  5000 31000]", 3),
        (8, 104, 3000, 4000, 1000, "[0]", "[]", "[]", 0),
        (9, 105, 4000, 27000, 27000, "[1 1 1 1 1 1]", "[103 999]", "[100000 5000]", 2),
    ]
    df = pd.DataFrame(rows, columns=["user_id", "video_id", "time_ms", "duration_ms", "playing_time_x", "label_1D",
                                     "history_items", "history_playing", "history_lengths"])
    uid_dict = {"7": ["105_25", "555_0"], "8": ["104_0"], "9": ["101_0"]}  # NB: a user with zero rows crashes the reference (IndexError at :259)

    class Corpus:
        data_df = {"test": df}
        user_input_dict = uid_dict

    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "SegMM"))
        json.dump({"7": 1, "8": 2, "9": 3}, open(os.path.join(td, "SegMM", "second_map_user2id.json"), "w"))
        json.dump({str(p): i + 1 for i, p in enumerate(nseg)}, open(os.path.join(td, "SegMM", "second_map_item2id.json"), "w"))
        os.chdir(td)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                ds = dl.FrameDatasetSeq_SegMM(corpus=Corpus(), lineid_map=lineid, feat_memmap=table, phase="test",
                                              shuffle=False, verbose=False)
                batch = dl.DataCollator()(list(ds))
        finally:
            os.chdir(cwd)
    save = {"out/" + k: v.numpy() for k, v in batch.items()}
    save["table"] = table
    save["lineid_json"] = json.dumps(lineid)
    save["user_input_json"] = json.dumps(uid_dict)
    save["rows_json"] = json.dumps(rows)
    save["user2id_json"] = json.dumps({"7": 1, "8": 2, "9": 3})
    save["item2id_json"] = json.dumps({str(p): i + 1 for i, p in enumerate(nseg)})
    np.savez_compressed(os.path.join(OUT, "gather_small.npz"), **save)
    print("gather_small", {k: tuple(v.shape) for k, v in batch.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    assert ref_shim.available(), "reference tree missing"
    run_gather_case()
    run_loss_cases()
    run_loss_cases_all()
    run_eval_case()
    run_topk_case()
    run_model_case("model_small_dh32", d_model=64, nhead=2, nlayers=3, din=48, Lt=12, B=5, seed=11)
    run_model_case("model_small_dh16", d_model=64, nhead=4, nlayers=4, din=40, Lt=20, B=4, seed=12)
    run_model_case("model_full_b4", d_model=512, nhead=16, nlayers=6, din=1024, Lt=100, B=4, seed=13,
                   store_sd=False, fill_seed=42)
    run_general_cases()
    run_ablation_cases()
    run_dropout_cases()
    run_mlp_ablation_cases()


def run_general_cases():
    run_general_case("model_both_small", {"user": "both", "photo": "both"}, d_model=64, nhead=2, nlayers=3, din=24, B=4, seed=21,
                     loss_types=("interestBPR",))
    run_general_case("model_id_small", {"user": "id", "photo": "id"}, d_model=64, nhead=2, nlayers=3, din=24, B=4, seed=22,
                     loss_types=("focal",))
    run_fusion_variants()
    run_bias_cases()
    run_nopos_cases()


def run_nopos_cases():
    """the 'noPos' ablation (encoder.py:428-429): random frame positions into the ID tower's frame projection, per call"""
    run_general_case("model_both_nopos", {"user": "both", "photo": "both"}, d_model=64, nhead=2, nlayers=3, din=24, B=4, seed=81,
                     loss_types=("interestBPR",), ablation_type="noPos", draw_seed=4321)
    run_general_case("model_id_nopos", {"user": "id", "photo": "id"}, d_model=64, nhead=2, nlayers=3, din=24, B=4, seed=82,
                     loss_types=("focal",), ablation_type="noPos", draw_seed=99)


def run_bias_cases():
    """the learnable position bias (decoder_leave_focal.py:442-444,497-504,650-658) in the reference's default
    'both' configuration with its default loss, and in the image-only focal configuration"""
    run_general_case("model_both_bias", {"user": "both", "photo": "both"}, d_model=64, nhead=2, nlayers=3, din=24, B=4, seed=71,
                     loss_types=("interestBPR",), learnable_bias=1)
    run_general_case("model_image_bias", {"user": "image", "photo": "image"}, d_model=64, nhead=2, nlayers=3, din=24, B=4, seed=72,
                     loss_types=("focal", "interestBPR"), learnable_bias=1)


def run_ablation_cases():
    """encoder ablations that select one attention block per query side (encoder.py:108-135,172-175)"""
    run_model_case("model_small_crossatt", d_model=64, nhead=2, nlayers=4, din=48, Lt=12, B=5, seed=41, ablation_type="CrossAtt")
    run_model_case("model_small_selfatt", d_model=64, nhead=2, nlayers=3, din=48, Lt=12, B=5, seed=42, ablation_type="SelfAtt")


def run_dropout_cases():
    """train() mode: nn.Dropout(0.1) live at every site, torch generator seeded right before the forward"""
    run_model_case("model_small_dropout", d_model=64, nhead=2, nlayers=4, din=48, Lt=12, B=5, seed=51, train_seed=1234)
    run_model_case("model_small_dropout_crossatt", d_model=64, nhead=2, nlayers=3, din=48, Lt=12, B=5, seed=52, train_seed=99,
                   ablation_type="CrossAtt")


def run_mlp_ablation_cases():
    """the MLP ablations of the encoder (encoder.py:392-400,503-511): MLP_Block instead of attention; one of them also in
    train() mode for the dropout placement inside MLP_Block"""
    run_model_case("model_small_selfmlp", d_model=64, nhead=2, nlayers=6, din=48, Lt=12, B=5, seed=61, ablation_type="SelfMLP")
    run_model_case("model_small_crossmlp", d_model=64, nhead=2, nlayers=6, din=48, Lt=12, B=5, seed=62, ablation_type="CrossMLP")
    run_model_case("model_small_woatt", d_model=64, nhead=2, nlayers=6, din=48, Lt=12, B=5, seed=63, ablation_type="w/oAtt")
    run_model_case("model_small_selfmlp_dropout", d_model=64, nhead=2, nlayers=6, din=48, Lt=12, B=5, seed=64, ablation_type="SelfMLP",
                   train_seed=777)


def run_fusion_variants():
    """the other fusions of forward (decoder_leave_focal.py:624-634): fusion_heads 0 (two Linear heads summed), -1 (one
    Linear(2d, 1) over the concatenation), -2 (one Linear head over the sum of the two backbones), -3 (list concatenation:
    backbone2's output alone through one Linear head; backbone1 is dead)"""
    for fh in (0, -1, -2, -3):
        run_general_case(f"model_both_fh{fh}", {"user": "both", "photo": "both"}, d_model=64, nhead=2, nlayers=3, din=24, B=4,
                         seed=30 - fh, loss_types=("focal",), fusion_heads=fh)


if __name__ == "__main__":
    if "--general-only" in sys.argv:
        run_general_cases()
    elif "--losses-only" in sys.argv:
        run_loss_cases_all()
    elif "--eval-only" in sys.argv:
        run_eval_case()
    elif "--topk-only" in sys.argv:
        run_topk_case()
    elif "--bias-only" in sys.argv:
        run_bias_cases()
    elif "--nopos-only" in sys.argv:
        run_nopos_cases()
    elif "--fusion-only" in sys.argv:
        run_fusion_variants()
    elif "--ablation-only" in sys.argv:
        run_ablation_cases()
    elif "--dropout-only" in sys.argv:
        run_dropout_cases()
    elif "--mlp-ablations-only" in sys.argv:
        run_mlp_ablation_cases()
    else:
        main()
