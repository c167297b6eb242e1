"""TEST INFRASTRUCTURE ONLY -- what bf16 does to the UNMODIFIED reference itself, fixture by fixture.

    python -m oracle.make_autocast_errors      (build container: needs /root/reference)

For every model fixture of tests/golden the reference module is rebuilt from the stored state_dict and run twice on the
stored inputs: in fp32 (must reproduce the fixture) and under `torch.autocast("cpu", dtype=torch.bfloat16)` -- the
reference's only route to bf16 (SURVEY 8c measured 4.4e-3 / 5.1e-3 logits / gradient error for it on the default model).
The relative L2 error of the autocast run against the fp32 run is stored per gradient tensor, for the logits, for the loss
and for the whole gradient in tests/golden/bf16_autocast_errors.json.  tests/test_gpu_model.py takes its bf16 bars from
these numbers: a tensor may miss the north-star 2e-2 only where the reference's own bf16 run misses it too.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.make_golden import _cuda_is_identity  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SIMPLE = ["model_small_dh32", "model_small_dh16", "model_small_crossatt", "model_small_selfatt", "model_small_selfmlp",
          "model_small_crossmlp", "model_small_woatt"]
GENERAL = ["model_both_small", "model_id_small", "model_both_fh0", "model_both_fh-1", "model_both_fh-2", "model_both_fh-3",
           "model_both_bias", "model_image_bias", "model_both_nopos", "model_id_nopos"]


def _build(z, cfg, general):
    if general:
        args = ref_shim.make_args(d_model=cfg["d_model"], nhead=cfg["nhead"], num_layers_enc=cfg["num_layers_enc"],
                                  loss_type_list=list(cfg["loss_types"]), input_type=cfg["input_type"], fusion_heads=cfg["fusion_heads"],
                                  learnable_bias=cfg.get("learnable_bias", 0), ablation_type=cfg.get("ablation_type", "ours"))
        model = ref_shim.build_reference_model_general(args, din=cfg["din"], n_users=cfg["n_users"], n_items=cfg["n_items"], seed=cfg["seed"])
    else:
        args = ref_shim.make_args(d_model=cfg["d_model"], nhead=cfg["nhead"], num_layers_enc=cfg["num_layers_enc"],
                                  loss_type_list=list(cfg["loss_types"]), ablation_type=cfg.get("ablation_type", "ours"))
        model = ref_shim.build_reference_model(args, din=cfg["din"], max_usr_len=cfg["Lt"], seed=cfg["seed"])
    model.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")})
    return model.eval()


def _run(model, z, general, autocast, draw_seed=None):
    B = z["usr_mask"].shape[0]
    if draw_seed is not None:          # 'noPos': the same frame permutations in both runs
        torch.manual_seed(draw_seed)
    for p in model.parameters():
        p.grad = None
    batch = dict(usr_image=torch.from_numpy(z["usr_image"]), usr_mask=torch.from_numpy(z["usr_mask"]), vid_image=torch.from_numpy(z["vid_image"]),
                 vid_mask=torch.from_numpy(z["vid_mask"]), gt=torch.from_numpy(z["gt_in"].copy()),
                 usr_id=torch.from_numpy(z["usr_id"]) if general else torch.zeros(B, dtype=torch.long),
                 vid_id=torch.from_numpy(z["vid_id"]) if general else torch.zeros(B, dtype=torch.long))
    with _cuda_is_identity(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        out = ref_shim.run_reference(model, batch, mode="train")
    out["loss"].float().backward()
    grads = {k: p.grad.detach().double().numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
    return out["logits"].detach().double().numpy(), float(out["loss"]), grads


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    res = {}
    for name in SIMPLE + GENERAL:
        path = os.path.join(GOLD, name + ".npz")
        if not os.path.exists(path):
            continue
        z = np.load(path)
        cfg = json.loads(str(z["cfg"]))
        general = name in GENERAL
        model = _build(z, cfg, general)
        lg32, loss32, g32 = _run(model, z, general, False, cfg.get("draw_seed"))
        assert rel(lg32, z["logits"].astype(np.float64)) < 1e-6, name          # the fp32 run IS the fixture
        lg16, loss16, g16 = _run(model, z, general, True, cfg.get("draw_seed"))
        e2 = r2 = 0.0
        per = {}
        for k, g in g32.items():
            d = float(np.linalg.norm(g16[k] - g))
            e2 += d * d
            r2 += float(np.linalg.norm(g)) ** 2
            per[k] = d / max(float(np.linalg.norm(g)), 1e-300)
        res[name] = dict(logits=rel(lg16, lg32), loss=abs(loss16 - loss32) / abs(loss32), whole_gradient=(e2 / r2) ** 0.5, grads=per)
        worst = sorted(((v, k) for k, v in per.items() if np.linalg.norm(g32[k]) > 1e-6 * r2 ** 0.5), reverse=True)[:3]
        print(f"{name:26s} logits {res[name]['logits']:.2e} loss {res[name]['loss']:.2e} whole grad {res[name]['whole_gradient']:.2e}  worst "
              + ", ".join(f"{k.split('.', 1)[-1]} {v:.2e}" for v, k in worst))
    res["_how"] = "unmodified reference, eval(), torch.autocast('cpu', bfloat16) vs its own fp32 run; torch " + torch.__version__
    json.dump(res, open(os.path.join(GOLD, "bf16_autocast_errors.json"), "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
