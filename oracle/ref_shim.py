"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference model files from
/root/reference behind import stubs (SURVEY.md section 8c recipe).

This module only works inside the build container (the GPU box has no
/root/reference).  It is used by ``oracle/make_golden.py`` to generate the
golden fixtures under ``tests/golden/`` and by CPU tests that pin the oracle
restatement (``oracle/mmi_oracle.py``) against the real reference.

Nothing in the product package imports this file.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types
from types import SimpleNamespace

REF_ROOT = os.environ.get("MMI_REFERENCE_ROOT", "/root/reference")
_MM = os.path.join(REF_ROOT, "MMinterest")

_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(_MM, "models", "encoder.py"))


def _load_by_path(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []  # behave like a package
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference's encoder / decoder modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    kn = os.path.join(_MM, "models", "kn_util")
    # empty package stubs so the reference's broken __init__ files never run
    _stub("kn_util")
    basic = _stub("kn_util.basic")
    nn_utils = _stub("kn_util.nn_utils")
    layers = _stub("kn_util.nn_utils.layers")
    # real leaf files, loaded by path
    init_m = _load_by_path("kn_util.nn_utils.init", os.path.join(kn, "nn_utils", "init.py"))
    ops_m = _load_by_path("kn_util.nn_utils.ops", os.path.join(kn, "nn_utils", "ops.py"))
    mlp_m = _load_by_path("kn_util.nn_utils.layers.mlp", os.path.join(kn, "nn_utils", "layers", "mlp.py"))
    _load_by_path("kn_util.nn_utils.math", os.path.join(kn, "nn_utils", "math.py"))
    bops = _load_by_path("kn_util.basic.ops", os.path.join(kn, "basic", "ops.py"))
    nn_utils.clones = ops_m.clones
    nn_utils.init_module = init_m.init_module
    layers.MLP = mlp_m.MLP
    basic.eval_env = bops.eval_env
    # names imported by decoder_leave_focal.py that only dead code uses
    _stub("model")
    _stub("model.ms_temporal_detr")
    _stub("model.ms_temporal_detr.ms_pooler", MultiScaleRoIAlign1D=object)
    _stub("misc", cw2se=None, calc_iou=None)
    models_pkg = _stub("models")
    models_pkg.__path__ = [os.path.join(_MM, "models")]
    _stub("models.loss", l1_loss=None, iou_loss=None)
    with contextlib.redirect_stdout(io.StringIO()):
        enc = _load_by_path("models.encoder", os.path.join(_MM, "models", "encoder.py"))
        dec = _load_by_path("models.decoder_leave_focal", os.path.join(_MM, "models", "decoder_leave_focal.py"))
    _loaded = SimpleNamespace(encoder=enc, decoder=dec,
                              SegFormerX=enc.SegFormerX,
                              MultiScaleTemporalDetrLeaveFocal=dec.MultiScaleTemporalDetrLeaveFocal)
    return _loaded


def load_dataloader():
    """Loads the reference's utils/dataloader_SegMM.py by path (np.int shim)."""
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # removed in numpy 1.24; the reference still uses it
    name = "_ref_dataloader_SegMM"
    if name in sys.modules:
        return sys.modules[name]
    with contextlib.redirect_stdout(io.StringIO()):
        return _load_by_path(name, os.path.join(_MM, "utils", "dataloader_SegMM.py"))


def load_evaluation():
    """Loads the reference's models/my_evaluation.py by path; matplotlib (absent here, only used by draw_hotmap) is
    replaced by empty stubs."""
    name = "_ref_my_evaluation"
    if name in sys.modules:
        return sys.modules[name]
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors"):
        if m not in sys.modules:
            try:
                importlib.import_module(m)
            except ImportError:
                _stub(m, Normalize=object)
    with contextlib.redirect_stdout(io.StringIO()):
        return _load_by_path(name, os.path.join(_MM, "models", "my_evaluation.py"))


def make_args(**over):
    """The argparse namespace fields the reference model reads."""
    a = dict(debug=0, input_type={"user": "image", "photo": "image"}, d_model=512, nhead=16,
             learnable_bias=0, exposure_prob=[1.0] * 40, fusion_heads=2,
             loss_type_list=["focal"],
             loss_weight={"focal": 1.0, "mse": 1.0, "hazard": 1.0, "surviveCE": 1.0,
                          "interestBPR": 1.0, "interestCE": 1.0, "interestKL": 1.0},
             mask_loss=0, num_layers_enc=6, ablation_type="ours", use_pe=1)
    a.update(over)
    return SimpleNamespace(**a)


def build_reference_model(args, din=1024, max_usr_len=100, max_vid_len=40, seed=42):
    """Mirrors init_model() of main_for_seq_leave_earlystop_SegMM.py:60-130 for the
    image-only single-backbone configuration."""
    import torch
    import torch.nn as nn
    ref = load()
    torch.manual_seed(seed)
    n = args.num_layers_enc
    with contextlib.redirect_stdout(io.StringIO()):
        bb = ref.SegFormerX(d_model_in=args.d_model, d_model_lvls=[args.d_model] * n,
                            num_head_lvls=[args.nhead] * n, ff_dim_lvls=[args.d_model] * n,
                            input_vid_dim=din, input_usr_dim=din,
                            max_vid_len=max_vid_len, max_usr_len=max_usr_len,
                            sr_ratio_lvls=[1] * n, use_patch_merge=[False] * n,
                            output_layers=[-1], model_cfg=args,
                            user_id_max=-1, video_id_max=-1, use_pe=args.use_pe)
        model = ref.MultiScaleTemporalDetrLeaveFocal(bb, None, None, nn.Identity(), args)
    return model


def build_reference_model_general(args, din=1024, n_users=0, n_items=0, max_vid_len=40, seed=42):
    """init_model() of main_for_seq_leave_earlystop_SegMM.py:60-130 for any input_type (image / id / both);
    n_users / n_items stand for reader.n_users / reader.n_items."""
    import torch
    import torch.nn as nn
    ref = load()
    torch.manual_seed(seed)
    n = args.num_layers_enc
    it = args.input_type

    def bb(user_id_max, video_id_max, max_usr_len):
        return ref.SegFormerX(d_model_in=args.d_model, d_model_lvls=[args.d_model] * n, num_head_lvls=[args.nhead] * n,
                              ff_dim_lvls=[args.d_model] * n, input_vid_dim=din, input_usr_dim=din, max_vid_len=max_vid_len,
                              max_usr_len=max_usr_len, sr_ratio_lvls=[1] * n, use_patch_merge=[False] * n, output_layers=[-1],
                              model_cfg=args, user_id_max=user_id_max, video_id_max=video_id_max, use_pe=args.use_pe)

    with contextlib.redirect_stdout(io.StringIO()):
        if it["user"] == "both" or it["photo"] == "both":
            u1, l1, u2, l2 = {"both": (-1, 100, n_users, 1), "id": (n_users, 1, n_users, 1), "image": (-1, 100, -1, 100)}[it["user"]]
            v1, v2 = {"both": (-1, n_items), "id": (n_items, n_items), "image": (-1, -1)}[it["photo"]]
            model = ref.MultiScaleTemporalDetrLeaveFocal(bb(u1, v1, l1), bb(u2, v2, l2), None, nn.Identity(), args)
        else:
            u1, l1 = (n_users, 1) if it["user"] == "id" else (-1, 100)
            v1 = n_items if it["photo"] == "id" else -1
            model = ref.MultiScaleTemporalDetrLeaveFocal(bb(u1, v1, l1), None, None, nn.Identity(), args)
    return model


def run_reference(model, batch, mode="train"):
    """Calls the reference forward with stdout swallowed (it prints each call)."""
    with contextlib.redirect_stdout(io.StringIO()):
        return model(usr_image=batch["usr_image"], usr_id=batch["usr_id"], usr_mask=batch["usr_mask"],
                     vid_image=batch["vid_image"], vid_id=batch["vid_id"], vid_mask=batch["vid_mask"],
                     gt=batch["gt"], mode=mode)
