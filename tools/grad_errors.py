#!/usr/bin/env python
"""Per-tensor and whole-gradient relative errors of the CUDA path against a golden fixture (diagnostic)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from test_gpu_model import GOLDEN, _run, make_args  # noqa: E402
from segmminterest_b200.model import build_model  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    for name in sys.argv[1:] or ["model_small_dh32", "model_small_crossatt", "model_small_selfatt"]:
        for precision in ("fp32", "bf16"):
            z = np.load(os.path.join(GOLDEN, name + ".npz"))
            cfg = json.loads(str(z["cfg"]))
            args = make_args(d_model=cfg["d_model"], nhead=cfg["nhead"], num_layers_enc=cfg["num_layers_enc"], mmi_precision=precision,
                             ablation_type=cfg.get("ablation_type", "ours"))
            model = build_model(args, din=cfg["din"], max_usr_len=cfg["Lt"])
            model.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")})
            model = model.cuda().eval()
            out = _run(model, z["usr_image"], z["usr_mask"], z["vid_image"], z["vid_mask"], z["gt_in"], dev)
            out["loss"].backward()
            rows, e2, r2 = [], 0.0, 0.0
            for k, p in model.named_parameters():
                if p.grad is None or ("grad/" + k) not in z.files:
                    continue
                ref = z["grad/" + k]
                err = float(np.linalg.norm(p.grad.double().cpu().numpy() - ref)); rn = float(np.linalg.norm(ref))
                e2 += err * err; r2 += rn * rn
                rows.append((err / max(rn, 1e-30), k, err, rn))
            lg = float(np.linalg.norm(out["logits"].detach().cpu().numpy() - z["logits"]) / np.linalg.norm(z["logits"]))
            print(f"== {name} {precision}: logits rel {lg:.2e}  whole-gradient rel {e2 ** 0.5 / r2 ** 0.5:.2e}")
            for r in sorted(rows, reverse=True)[:8]:
                print(f"   {r[0]:.3e}  err {r[2]:.3e}  ||ref|| {r[3]:.3e}  {r[1]}")


if __name__ == "__main__":
    main()
