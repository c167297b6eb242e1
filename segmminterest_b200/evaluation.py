"""Drop-in for the part of MMinterest/models/my_evaluation.py the driver's validation loop uses
(main_for_seq_leave_earlystop_SegMM.py:5,396-432): `main_eval_batch` with the metrics computed on the device by
`mmi_eval_metrics` -- one kernel pass and one small D2H copy per batch instead of the reference's per-row Python loop
(`.item()` / `.cpu()` per row, sklearn on host copies).

`results_list` keeps the reference's protocol: a dict whose KEYS select the metrics and whose values are lists that get
one entry per batch (ProbAUC) or per row (JaccardSim, LeaveMSE + view_lengths [+ duration_lengths], LeaveCTR,
LeaveCTR_view).  TOP_K_leave / TOP_K_leave_mask -- called by the driver's validation loop on host numpy copies
(main...SegMM.py:164-167) and shuffling every row with np.random.permutation -- are host numpy functions here as well,
drawing their permutations in the reference's order (a seeded run reproduces the reference's HR@k / NDCG@k exactly).
"""
from __future__ import annotations

import torch

from . import _lib, ops

_ROW = {"pred": 0, "view": 1, "duration": 2, "LeaveCTR": 3, "LeaveCTR_view": 4, "JaccardSim": 5}


class DeviceMetrics:
    """Workspace holder for mmi_eval_metrics.  __call__(logits, gt, exposure_prob) -> (rows [B,6], out [4]) device tensors."""

    def __init__(self):
        self._buf = {}

    def __call__(self, logits: torch.Tensor, gt: torch.Tensor, exposure_prob=None, interests=False, old=False):
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
        ops.eval_metrics(logits.float().contiguous(), gt.to(torch.int64).contiguous(), ep, rows, out, ws, interests=interests, old=old)
        return rows, out


_METRICS = DeviceMetrics()


def prob_auc_batch(logits, gt, exposure_prob) -> torch.Tensor:
    """ProbAUC_batch (my_evaluation.py:73-80) of survival = exp(cumsum(log(sigmoid(logits) * exposure_prob))): a 0-d device
    tensor (no sync)."""
    _, out = _METRICS(logits, gt, exposure_prob)
    return out[0].clone()


def main_eval_batch(args, interests, ground_truths, pred_labels, results_list, type="inference", test_type="new", logits=None):
    """my_evaluation.py:264-357.  `interests` = sigmoid(logits) * exposure_prob as the driver builds them
    (main...SegMM.py:402-403); test_type 'old' takes them as survival probabilities directly (:270-271), anything else
    builds exp(cumsum(log interests)) (:273-274).  `logits=` adds the MAES bookkeeping of :309-320."""
    if getattr(args, "draw_case", 0):
        raise NotImplementedError("draw_case needs matplotlib on the host (my_evaluation.py:233-262)")
    rows, out = _METRICS(interests, ground_truths, interests=True, old=test_type == "old")
    host = torch.cat([rows.reshape(-1), out]).cpu()          # the one D2H copy (and sync) of the batch
    r, o = host[:-4].view(-1, 6), host[-4:]
    if "ProbAUC" in results_list:
        if not (o[1] > 0 and o[2] > 0):
            raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")   # sklearn's error
        results_list["ProbAUC"].append(float(o[0]))
    if "TOP_K" in results_list:                               # my_evaluation.py:287-303: host metrics on numpy copies, like the reference
        view_lengths = (ground_truths == 1).sum(dim=1, keepdim=True).cpu().numpy()
        mask_np = (ground_truths != -2).cpu().numpy()
        inter_np = interests.detach().float().cpu().numpy()
        if getattr(args, "TOP_K_mask", 0):
            evaluations = TOP_K_leave_mask(inter_np, view_lengths, mask_np, permutation=args.TOP_K_permutation)
        elif "TOP1MSE" in results_list:
            evaluations, top1 = TOP_K_leave(inter_np, view_lengths, mask_np, permutation=args.TOP_K_permutation, test=1)
            results_list["TOP1MSE"].append(top1)
        else:
            evaluations = TOP_K_leave(inter_np, view_lengths, mask_np, permutation=args.TOP_K_permutation)
        for name, value in evaluations.items():
            results_list.setdefault(name, []).append(float(value))
    if logits is not None:
        # :309-320 -- expected leave position under softmax(1 / softmax(logits)); a [B, 40] host computation in the reference
        # too (its `pos` lives on the CPU), accumulated as a running sum in results_list['MAES']
        lg = logits.detach().float().cpu()
        inv = 1 / torch.nn.functional.softmax(lg, dim=1)
        leave_p = inv / inv.sum(dim=1).unsqueeze(1)
        pred_leave = torch.sum(leave_p * torch.linspace(0, 39, 40), dim=1).int()
        views = (ground_truths == 1).sum(dim=1).cpu()
        mae = float((views - pred_leave).abs().double().mean())          # sklearn.metrics.mean_absolute_error
        results_list["MAES"] += mae * interests.shape[0]
        results_list["pred_leave"].append(pred_leave)
    for i in range(r.shape[0]):                               # per-row appends, in the reference's row-major order (:324-355)
        for eval_type in list(results_list):
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


def _rank_metrics(scores, targets, permutation):
    """HR@k / NDCG@k (k = 1, 3, 5, 10) of the rank of position `targets[i]` in the ascending order of row i of `scores`.
    permutation != 0 breaks ties by a fresh np.random.permutation per row, drawn in row order exactly like the reference
    (so a seeded validation run reproduces the reference's numbers); 0 leaves ties to np.argsort."""
    import numpy as np
    n, L = scores.shape
    if permutation:
        perms = np.stack([np.random.permutation(L) for _ in range(n)]) if n else np.zeros((0, L), dtype=np.int64)
        shuffled = np.take_along_axis(scores, perms, axis=1)
        where = (perms == targets[:, None]).argmax(axis=1)          # where the target went
        order = np.argsort(shuffled, axis=1)
    else:
        where = targets
        order = np.argsort(scores, axis=1)
    rank = (order == where[:, None]).argmax(axis=1) + 1
    out = {}
    for k in (1, 3, 5, 10):
        hit = (rank <= k).astype(np.float32)
        out[f"HR@{k}"] = hit.mean()
        out[f"NDCG@{k}"] = (hit / np.log2(rank + 1)).mean()
    return out


def TOP_K_leave(interests, view_lengths, mask_batch, permutation=1, test=0):
    """models/my_evaluation.py:180-231 -- the metric the driver's validation loop calls on every batch
    (main...SegMM.py:164-167) with host numpy arrays (`interests.cpu().detach().numpy()`), so it is host numpy here too:
    among rows that were not watched to the end of the 40 positions, the rank of the skip position (index view_length) in
    the ascending order of the interests; HR@k and NDCG@k.  test != 0 also returns argmin(interests) per (unfiltered) row."""
    import numpy as np
    interests = np.asarray(interests)
    first_min = np.argmin(interests, axis=1)
    views = np.asarray(view_lengths).astype(np.int64).reshape(-1)
    rows = views < 40
    ev = _rank_metrics(interests[rows], views[rows], permutation)
    return (ev, first_min) if test else ev


def TOP_K_leave_mask(interests, view_lengths, mask_batch, permutation=1):
    """models/my_evaluation.py:137-178: like TOP_K_leave over the rows whose view length differs from the number of real
    segments (the video was skipped), with the interests of padded positions replaced by 1.1 so they rank last."""
    import numpy as np
    interests, mask_batch = np.asarray(interests), np.asarray(mask_batch)
    views = np.asarray(view_lengths).astype(np.int64).reshape(-1)
    rows = views != mask_batch.sum(axis=1)
    return _rank_metrics(np.where(mask_batch[rows], interests[rows], 1.1), views[rows], permutation)
