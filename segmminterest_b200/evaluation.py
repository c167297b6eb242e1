"""Drop-in for the part of MMinterest/models/my_evaluation.py the driver's validation loop uses
(main_for_seq_leave_earlystop_SegMM.py:5,396-432): `main_eval_batch` with the metrics computed on the device by
`mmi_eval_metrics` -- one kernel pass and one small D2H copy per batch instead of the reference's per-row Python loop
(`.item()` / `.cpu()` per row, sklearn on host copies).

`results_list` keeps the reference's protocol: a dict whose KEYS select the metrics and whose values are lists that get
one entry per batch (ProbAUC) or per row (JaccardSim, LeaveMSE + view_lengths [+ duration_lengths], LeaveCTR,
LeaveCTR_view).  The TOP_K family shuffles every row with np.random.permutation (my_evaluation.py:92-231) and stays on
the host in the reference; it is not built here.
"""
from __future__ import annotations

import torch

from . import _lib, ops

_ROW = {"pred": 0, "view": 1, "duration": 2, "LeaveCTR": 3, "LeaveCTR_view": 4, "JaccardSim": 5}


class DeviceMetrics:
    """Workspace holder for mmi_eval_metrics.  __call__(logits, gt, exposure_prob) -> (rows [B,6], out [4]) device tensors."""

    def __init__(self):
        self._buf = {}

    def __call__(self, logits: torch.Tensor, gt: torch.Tensor, exposure_prob=None, interests=False):
        if not logits.is_cuda:
            raise _lib.MMIError("DeviceMetrics needs CUDA tensors; there is no CPU fallback")
        B, L = logits.shape
        key = (B, L, logits.device)
        if key not in self._buf:
            nbytes = int(_lib.load().mmi_eval_metrics_workspace(B, L))
            self._buf[key] = (torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=logits.device),
                              torch.empty(B, 6, dtype=torch.float32, device=logits.device),
                              torch.empty(4, dtype=torch.float32, device=logits.device))
        ws, rows, out = self._buf[key]
        ep = None
        if not interests:
            ep = exposure_prob if torch.is_tensor(exposure_prob) else torch.tensor(list(exposure_prob)[:L], dtype=torch.float32)
            ep = ep.to(logits.device, torch.float32).contiguous()
        ops.eval_metrics(logits.float().contiguous(), gt.to(torch.int64).contiguous(), ep, rows, out, ws, interests=interests)
        return rows, out


_METRICS = DeviceMetrics()


def prob_auc_batch(logits, gt, exposure_prob) -> torch.Tensor:
    """ProbAUC_batch (my_evaluation.py:73-80) of survival = exp(cumsum(log(sigmoid(logits) * exposure_prob))): a 0-d device
    tensor (no sync)."""
    _, out = _METRICS(logits, gt, exposure_prob)
    return out[0].clone()


def main_eval_batch(args, interests, ground_truths, pred_labels, results_list, type="inference", test_type="new", logits=None):
    """my_evaluation.py:264-357 for test_type 'new'.  `interests` = sigmoid(logits) * exposure_prob as the driver builds
    them (main...SegMM.py:402-403)."""
    if test_type != "new":
        raise NotImplementedError("test_type 'old' (interests already are survival probabilities) is not built")
    if "TOP_K" in results_list or "TOP1MSE" in results_list:
        raise NotImplementedError("TOP_K_leave* shuffle every row on the host with np.random.permutation "
                                  "(my_evaluation.py:92-231); not part of the device path")
    if logits is not None:
        raise NotImplementedError("the `logits=` branch (MAES, my_evaluation.py:309-320) is never taken by the SegMM driver")
    if getattr(args, "draw_case", 0):
        raise NotImplementedError("draw_case needs matplotlib on the host (my_evaluation.py:233-262)")
    rows, out = _METRICS(interests, ground_truths, interests=True)
    host = torch.cat([rows.reshape(-1), out]).cpu()          # the one D2H copy (and sync) of the batch
    r, o = host[:-4].view(-1, 6), host[-4:]
    if "ProbAUC" in results_list:
        if not (o[1] > 0 and o[2] > 0):
            raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")   # sklearn's error
        results_list["ProbAUC"].append(float(o[0]))
    for i in range(r.shape[0]):                               # per-row appends, in the reference's row-major order (:324-355)
        for eval_type in results_list:
            if eval_type == "JaccardSim":
                results_list[eval_type].append(float(r[i, _ROW["JaccardSim"]]))
            elif eval_type == "LeaveMSE":
                results_list[eval_type].append(float(r[i, _ROW["pred"]]))
                results_list["view_lengths"].append(float(r[i, _ROW["view"]]))
                if "duration_lengths" in results_list:
                    results_list["duration_lengths"].append(float(r[i, _ROW["duration"]]))
            elif eval_type == "LeaveCTR":
                results_list[eval_type].append(float(r[i, _ROW["LeaveCTR"]]))
            elif eval_type == "LeaveCTR_view":
                results_list[eval_type].append(float(r[i, _ROW["LeaveCTR_view"]]))
    return results_list


def TOP_K_leave(*a, **k):
    """Imported by the driver (main...SegMM.py:5); host-side metric with per-row np.random.permutation -- not built."""
    raise NotImplementedError("TOP_K_leave is a host-side metric (my_evaluation.py:180-231); not part of the device path")


def TOP_K_leave_mask(*a, **k):
    raise NotImplementedError("TOP_K_leave_mask is a host-side metric (my_evaluation.py:137-178); not part of the device path")
