"""SURVEY 8b import surface: `main_for_seq_leave_earlystop_SegMM.py` against segmminterest_b200/dropin.

* With the reference tree present (build container) the driver's OWN source is executed -- unchanged except for its syntax
  error at :32 (a stray full-width parenthesis after `torch.cuda.manual_seed(seed_value)`) -- with `dropin/` in front of
  sys.path: its imports resolve, `load_data(args)` builds the three loaders, `init_model(args, reader)` builds our modules
  for the image-only and the default 'both' input types with the reference's state_dict schema.  On a machine that has
  both the reference and a GPU, `main(args)` itself runs in --debug mode.
* On the GPU box (no reference tree) the same statements, restated, run end to end on the config-1 fixture: loaders ->
  model -> the driver's training statements -> validation statements -> CheckPointer round trip -> inference loop with
  main_eval_batch.
"""
import json
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "segmminterest_b200", "dropin")
REF = os.environ.get("MMI_REFERENCE_ROOT", "/root/reference")
DRIVER = os.path.join(REF, "MMinterest", "main_for_seq_leave_earlystop_SegMM.py")
GOLD = os.path.join(ROOT, "tests", "golden", "config1.npz")


def _args(**over):
    """the argparse defaults of main...SegMM.py:458-576 after its post-processing"""
    a = dict(train_batch_size=64, valid_batch_size=64, test_batch_size=64, threshold=0.5, learnable_bias=0, wandb=0, exp="", logging_step=1,
             valid_step=2, ckpt_dir="ckpts_SegMM", d_model=64, ff_dim=64, nhead=2, num_query=1, num_clips=1, num_layers_enc=3,
             num_layers_dec=0, dropout=0.1, epochs=1, exposure_prob_type="ones", debug=1, learning_rate=1e-3, weight_decay=1e-4,
             input_type={"user": "image", "photo": "image"}, user_input_type="image", photo_input_type="image", loss_type="interestBPR",
             loss_type_list=["interestBPR"], loss_weight={"focal": 1.0, "mse": 1.0, "hazard": 1.0, "surviveCE": 1.0, "interestBPR": 1.0,
                                                           "interestCE": 1.0, "interestKL": 1.0},
             loss_weight_surviveCE=1.0, loss_weight_interestBPR=1.0, loss_weight_interestCE=1.0, use_pe=1, test_model=1, save_logits=0,
             eval_type_list=["JaccardSim", "LeaveMSE", "LeaveCTR", "LeaveCTR_view", "TOP_K"], draw_case=0, early_stop=20, main_metrics="HR@5",
             TOP_K_permutation=1, record_train_detail=0, mask_loss=0, count_view_completion=0, TOP_K_mask=0, fusion_heads=2, eval_cold="",
             ablation_type="ours", exposure_prob=[1.0] * 40, sep="\t", path="SegMM/", data="inter", dict_path="user_input_dict.json",
             history_max=50)
    a.update(over)
    return SimpleNamespace(**a)


def _workdir(td):
    """the files the driver opens with relative paths (main...SegMM.py:35-40, utils/dataloader_SegMM.py:59,207-210)"""
    gold = np.load(GOLD)
    seg = os.path.join(td, "SegMM")
    os.makedirs(seg, exist_ok=True)
    for n in ("train_his.csv", "dev_his.csv", "test_his.csv", "user_input_dict.json", "second_map_user2id.json", "second_map_item2id.json"):
        with open(os.path.join(seg, n), "wb") as f:
            f.write(gold["files/" + n].tobytes())
    with open(os.path.join(td, "SegMM_photoidframeid2lineid.json"), "wb") as f:
        f.write(gold["files/SegMM_photoidframeid2lineid.json"].tobytes())
    meta = json.loads(str(gold["meta"]))
    table = np.random.default_rng(meta["table_seed"]).standard_normal((meta["n_rows"], meta["din"]), dtype=np.float32)
    table.tofile(os.path.join(td, "SegMM_feat_memmap.dat"))                 # float32 rows of 1024, as the driver declares (:39)
    for d in ("logs_new/SegMM", "ckpts_SegMM", "eval_results_new/SegMM/results_all_points", "pics/SegMM", "DebugAndCheck/SegMM"):
        os.makedirs(os.path.join(td, d), exist_ok=True)
    return meta


class _Surface:
    """puts dropin/ in front of sys.path (and stubs the plotting / tracking packages that are not ours) for one test"""
    NAMES = ("model", "utils", "kn_util", "kn_util.nn_utils")

    def __enter__(self):
        self.saved_path = list(sys.path)
        self.saved_mods = {n: sys.modules.pop(n, None) for n in self.NAMES}
        self.stubbed = []
        sys.path.insert(0, DROPIN)
        for name in ("wandb", "matplotlib", "matplotlib.pyplot"):
            try:
                __import__(name)
            except ImportError:
                m = types.ModuleType(name)
                for fn in ("figure", "plot", "title", "savefig", "close"):
                    setattr(m, fn, lambda *a, **k: None)
                sys.modules[name] = m
                self.stubbed.append(name)
        if "matplotlib" in self.stubbed:
            sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        return self

    def __exit__(self, *exc):
        sys.path[:] = self.saved_path
        for n in self.NAMES:
            sys.modules.pop(n, None)
            if self.saved_mods[n] is not None:
                sys.modules[n] = self.saved_mods[n]
        for n in self.stubbed:
            sys.modules.pop(n, None)


def _load_driver():
    src = open(DRIVER, encoding="utf-8").read()
    bad = "torch.cuda.manual_seed(seed_value) ）"
    assert bad in src, "the driver's syntax error at :32 moved"
    src = src.replace(bad, "torch.cuda.manual_seed(seed_value)")
    ns = {"__name__": "mmi_reference_driver", "__file__": DRIVER}
    exec(compile(src, DRIVER, "exec"), ns)
    return ns


def test_import_surface_resolves_to_the_b200_path():
    import segmminterest_b200 as pkg
    with _Surface():
        import model
        import utils
        from kn_util.nn_utils import CheckPointer
        assert model.MultiScaleTemporalDetrLeaveFocal is pkg.MultiScaleTemporalDetrLeaveFocal and model.SegFormerX is pkg.SegFormerX
        assert model.main_eval_batch is pkg.main_eval_batch and model.TOP_K_leave is pkg.TOP_K_leave and model.TOP_K_leave_mask is pkg.TOP_K_leave_mask
        assert isinstance(model.QueryBasedDecoder, type)
        from segmminterest_b200 import checkpoint, dataset, reader
        assert utils.BaseReaderSeq_SegMM is reader.BaseReaderSeq_SegMM and utils.FrameDatasetSeq_SegMM is dataset.FrameDatasetSeq_SegMM
        assert utils.DataCollator is dataset.DataCollator and utils.DataLoader is dataset.DataLoader
        assert hasattr(utils, "BaseReaderSeq_SegMM_sampled") and hasattr(utils, "FrameDatasetSeq_SegMM_sampled")
        assert CheckPointer is checkpoint.CheckPointer
        ck = CheckPointer("main_metric", "some/dir", mode="max", cur_time="2026-01-01-00:00:00")     # main...SegMM.py:217
        assert ck.better(2.0, 1.0) and not ck.better(1.0, 2.0)


def test_checkpointer_round_trip(tmp_path):
    """kn_util/nn_utils/checkpoint.py:11-77 as the driver uses it (:333,366-367)"""
    from segmminterest_b200.checkpoint import CheckPointer
    net = torch.nn.Linear(3, 2)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    ck = CheckPointer("main_metric", str(tmp_path / "ckpts" / "run"), mode="max", cur_time="t0")
    assert ck.save_checkpoint(model=net, optimizer=opt, num_epochs=0, metric_vals={"main_metric": 0.25}) is True
    w_best = net.weight.detach().clone()
    with torch.no_grad():
        net.weight.add_(1.0)
    assert ck.save_checkpoint(model=net, optimizer=opt, num_epochs=1, metric_vals={"main_metric": 0.2}) is False     # not better
    assert sorted(os.listdir(ck.work_dir)) == ["ckpt-best-ep0-0.25.pth", "ckpt-latest.pth"]
    load_dict = ck.load_checkpoint(net, opt, mode="best")
    assert torch.equal(net.weight, w_best) and load_dict["num_epochs"] == 0 and load_dict["metrics"] == {"main_metric": 0.25}
    assert set(load_dict) >= {"model", "optimizer", "num_epochs", "metrics"}
    assert ck.save_checkpoint(model=net, optimizer=opt, num_epochs=2, metric_vals={"main_metric": 0.5}) is True
    assert sorted(os.listdir(ck.work_dir)) == ["ckpt-best-ep2-0.5.pth", "ckpt-latest.pth"]              # the old best is removed


@pytest.mark.skipif(not os.path.isfile(DRIVER), reason="reference tree not present (GPU box)")
def test_unmodified_driver_functions_run_against_the_dropin(tmp_path):
    meta = _workdir(str(tmp_path))
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        with _Surface():
            drv = _load_driver()
            import model as model_pkg
            assert drv["MultiScaleTemporalDetrLeaveFocal"] is model_pkg.MultiScaleTemporalDetrLeaveFocal
            args = _args()
            reader, train_dl, valid_dl, test_dl = drv["load_data"](args)              # main...SegMM.py:42-58, unmodified
            assert (reader.n_users, reader.n_items) == (1903, 352494)
            assert len(train_dl) == (len(reader.data_df["train"]) + 63) // 64 and len(valid_dl) > 0 and len(test_dl) > 0
            assert train_dl.dataset.feat_memmap.shape == (meta["n_rows"], 1024)
            m = drv["init_model"](args, reader)                                       # :60-130, unmodified
            from segmminterest_b200.model import reference_state_shapes
            assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == list(reference_state_shapes(64, 3, 1024, 100, 40).items())
            args2 = _args(input_type={"user": "both", "photo": "both"}, user_input_type="both", photo_input_type="both", d_model=32)
            m2 = drv["init_model"](args2, SimpleNamespace(n_users=23, n_items=57))
            keys = set(m2.state_dict().keys())
            assert {"backbone2.vid_proj.weight", "backbone2.frameid_proj.weight", "fusion_module.w_xy", "backbone1.usr_proj.bias"} <= keys
            assert m2.state_dict()["backbone2.vid_proj.weight"].shape == (58, 16) and m2.state_dict()["backbone2.usr_proj.weight"].shape == (24, 32)
            assert drv["compute_final_result"]({"JaccardSim": [0.5, 1.0], "LeaveMSE": [1.0, 2.0], "view_lengths": [1.0, 2.0]}) == \
                {"LeaveMSE": 0.0, "JaccardSim": 0.75}
            if not torch.cuda.is_available():
                with pytest.raises(Exception, match="CUDA|cuda|no CPU fallback"):
                    next(iter(train_dl))                                              # fails loudly, no CPU path
            else:                                                                     # reference AND a GPU: the whole main()
                drv["main"](_args(debug=1, epochs=1))
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
def test_driver_statements_end_to_end_on_config1(tmp_path):
    """The driver's statements (load_data :42-58, init_model :60-130, the loop body :264-300, valid_model :132-186, the
    checkpoint calls :333,366-367 and the test loop :388-405), restated because the reference tree is not on the GPU box,
    against the drop-in names only."""
    import random
    _workdir(str(tmp_path))
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        with _Surface():
            from kn_util.nn_utils import CheckPointer
            from model import MultiScaleTemporalDetrLeaveFocal, SegFormerX, TOP_K_leave, main_eval_batch
            from utils import BaseReaderSeq_SegMM, DataCollator, DataLoader, FrameDatasetSeq_SegMM
            args = _args(loss_type_list=["interestBPR", "focal"], eval_type_list=["JaccardSim", "ProbAUC", "LeaveMSE", "LeaveCTR", "LeaveCTR_view"])
            random.seed(42); np.random.seed(42); torch.manual_seed(42)
            lineid = json.load(open("SegMM_photoidframeid2lineid.json"))
            feat = np.memmap("SegMM_feat_memmap.dat", dtype="float32", mode="r", shape=(len(lineid), 1024))
            reader = BaseReaderSeq_SegMM(args)
            mk = lambda phase, sh: FrameDatasetSeq_SegMM(corpus=reader, lineid_map=lineid, feat_memmap=feat, phase=phase, shuffle=sh,  # noqa: E731
                                                         do_scale_image_to_01=True, image_resize=True, verbose=False)
            train_dl = DataLoader(mk("train", True), args.train_batch_size, collate_fn=DataCollator())
            valid_dl = DataLoader(mk("dev", False), args.valid_batch_size, collate_fn=DataCollator())
            n = args.num_layers_enc
            bb = SegFormerX(d_model_in=args.d_model, d_model_lvls=[args.d_model] * n, num_head_lvls=[args.nhead] * n, ff_dim_lvls=[args.d_model] * n,
                            input_vid_dim=1024, input_usr_dim=1024, max_vid_len=40, max_usr_len=100, sr_ratio_lvls=[1] * n,
                            use_patch_merge=[False] * n, output_layers=[-1], model_cfg=args, user_id_max=-1, video_id_max=-1, use_pe=args.use_pe)
            model = MultiScaleTemporalDetrLeaveFocal(bb, None, None, torch.nn.Identity(), args).cuda()
            ckpt = CheckPointer("main_metric", os.path.join(args.ckpt_dir, "run"), mode="max", cur_time="t")
            param_dict = model.parameters()
            optimizer = torch.optim.AdamW(param_dict, lr=args.learning_rate, weight_decay=args.weight_decay)
            losses = []
            for local_step, batch in enumerate(train_dl):
                if local_step > 3:
                    break
                optimizer.zero_grad()
                model.train()
                batch = {k: v.cuda() for k, v in batch.items()}
                usr_feat = batch["user"] / (batch["user"].norm(p=1, dim=-1, keepdim=True) + 1e-6)
                vid_feat = batch["photo"] / (batch["photo"].norm(p=1, dim=-1, keepdim=True) + 1e-6)
                out = model(usr_image=usr_feat, usr_id=batch["user_identity_id"], usr_mask=batch["user_mask"], vid_image=vid_feat,
                            vid_id=batch["photo_identity_id"], vid_mask=batch["photo_mask"], gt=batch["label"], mode="train")
                loss = out["loss"]
                loss.backward()
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    torch.nn.utils.clip_grad_norm_(param_dict, 10.0).item()
                optimizer.step()
                losses.append(loss.item())
                assert set(out) >= {"loss", "interestBPR", "focal", "mse", "mse2", "logits", "gt"} and out["logits"].shape == (64, 40)
            assert all(np.isfinite(losses))
            # valid_model (:132-186): mode="train" under eval(), TOP_K on the host
            model.eval()
            for local_valid_step, vb in enumerate(valid_dl):
                if local_valid_step > 1:
                    break
                vb = {k: v.cuda() for k, v in vb.items()}
                usr_feat = vb["user"] / (vb["user"].norm(p=1, dim=-1, keepdim=True) + 1e-6)
                vid_feat = vb["photo"] / (vb["photo"].norm(p=1, dim=-1, keepdim=True) + 1e-6)
                vo = model(usr_image=usr_feat, usr_id=vb["user_identity_id"], usr_mask=vb["user_mask"], vid_image=vid_feat,
                           vid_id=vb["photo_identity_id"], vid_mask=vb["photo_mask"], gt=vb["label"], mode="train")
                interests = torch.sigmoid(vo["logits"]) * torch.tensor(args.exposure_prob).cuda()[None]
                gt = vo["gt"]
                ev = TOP_K_leave(interests.cpu().detach().numpy(), (gt == 1).sum(dim=1, keepdim=True).cpu().numpy(),
                                 (gt != -2).cpu().detach().numpy(), permutation=args.TOP_K_permutation)
                assert "HR@5" in ev and np.isfinite(float(vo["loss"].item()))
            assert ckpt.save_checkpoint(model=model, optimizer=optimizer, num_epochs=0, metric_vals={"main_metric": float(ev["HR@5"])}) is True
            before = {k: v.clone() for k, v in model.state_dict().items()}
            load_dict = ckpt.load_checkpoint(model, optimizer, mode="best")
            model.load_state_dict(load_dict["model"])
            model = model.cuda()
            model.eval()
            assert all(torch.equal(v, model.state_dict()[k]) for k, v in before.items())
            results = {k: [] for k in args.eval_type_list + ["view_lengths"]}
            for i, tb in enumerate(valid_dl):
                if i > 0:
                    break
                tb = {k: v.cuda() for k, v in tb.items()}
                with torch.no_grad():
                    usr_feat = tb["user"] / (tb["user"].norm(p=1, dim=-1, keepdim=True) + 1e-6)
                    vid_feat = tb["photo"] / (tb["photo"].norm(p=1, dim=-1, keepdim=True) + 1e-6)
                    o = model(usr_image=usr_feat, usr_id=tb["user_identity_id"], usr_mask=tb["user_mask"], vid_image=vid_feat,
                              vid_id=tb["photo_identity_id"], vid_mask=tb["photo_mask"], gt=tb["label"], mode="inference")
                interests = torch.sigmoid(o["logits"]) * torch.tensor(args.exposure_prob).cuda()[None]
                pred_label = torch.where(interests > args.threshold, 1.0, 0.0)
                results = main_eval_batch(args, interests, o["gt"], pred_label, results, type="inference")
            assert len(results["JaccardSim"]) == 64 and 0.0 <= results["ProbAUC"][0] <= 1.0
    finally:
        os.chdir(cwd)
