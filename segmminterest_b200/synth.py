"""Seeded synthetic workloads of the shapes BASELINE.json names (SURVEY.md
section 8d).  numpy Generators only, so the streams are identical in the build
container and on the GPU box regardless of torch version.

Vocabulary follows the reference: an *interaction* is one (user history,
candidate video) pair; a video is cut into 5-second *segments*; the embedding
*table* holds one row per (video, segment) (reference: SegMM_feat_memmap.dat,
main_for_seq_leave_earlystop_SegMM.py:35-40).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

PHOTO_MAX = 40  # utils/dataloader_SegMM.py:198
PAD_LABEL = -2  # utils/dataloader_SegMM.py:240


@dataclass(frozen=True)
class Workload:
    name: str
    batch: int
    hist_videos: int      # H
    segs_per_video: int   # S   -> Lt = H*S
    din: int
    n_rows: int = 1 << 20

    @property
    def lt(self) -> int:
        return self.hist_videos * self.segs_per_video


def model_flops(nt: int, nv: int, din: int, d: int = 512, n_layers: int = 6, train: bool = True) -> float:
    """Algorithmic FLOPs of ONE interaction through one image backbone: live compute only, valid
    tokens only (nt history / nv candidate tokens), MAC = 2 FLOP; the dead layer N-1 and pad rows are not credited.
      MACs_fwd = (nt+nv) Din d  +  (N-2) [9 (nt+nv) d^2 + 2 (nt+nv)^2 d]  +  [(7 nv + 2 nt) d^2 + 2 nv (nv+nt) d]  +  nv d
    (input projection; N-2 full layers = 12 q/k/v projections + 2 output projections + 2 x 2 FFN + QK^T + PV; the
    candidate-only layer N-2; head).  Training = 3 x forward."""
    t = nt + nv
    macs = t * din * d + (n_layers - 2) * (9 * t * d * d + 2 * t * t * d) + ((7 * nv + 2 * nt) * d * d + 2 * nv * (nv + nt) * d) + nv * d
    return 2.0 * macs * (3.0 if train else 1.0)


WORKLOADS = {
    # configs[0]: reference shapes (Lt=100, Din=1024)
    "c1": Workload("c1_segmm_reference_shapes", 1024, 10, 10, 1024, 1 << 17),
    # configs[1]: the config the metric is quoted on at N=1
    "c2": Workload("c2_segmm_scale_h50xs10_din640_b1024", 1024, 50, 10, 640),
    "c3": Workload("c3_kuairand_h200xs20_din768_b4096", 4096, 200, 20, 768),
    "c4": Workload("c4_long_history_h1000xs30_din768_b8192", 8192, 1000, 30, 768),
}


def make_table(n_rows: int, din: int, seed: int = 1234, dtype=np.float32) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n_rows, din), dtype=np.float32).astype(dtype)


def make_labels(rng: np.random.Generator, nv: np.ndarray) -> np.ndarray:
    """construct_label_1D semantics (data_process/get_data_SegMM_public.py:45-89):
    1 = watched, 0 = the skip segment, -1 = after the skip, -2 = pad."""
    B = nv.shape[0]
    gt = np.full((B, PHOTO_MAX), PAD_LABEL, dtype=np.int64)
    view = rng.integers(0, nv + 1)
    for b in range(B):
        v, n = int(view[b]), int(nv[b])
        gt[b, :v] = 1
        if v < n:
            gt[b, v] = 0
            gt[b, v + 1:n] = -1
    return gt


def make_indices(batch: int, lt: int, nv_max: int, n_rows: int, seed: int = 2025, ragged: bool = False,
                 hist_videos: int | None = None, segs_per_video: int | None = None):
    """Row indices into the table; -1 = pad.  Throughput runs use full
    histories; parity runs use ragged ones."""
    rng = np.random.default_rng(seed)
    usr_idx = rng.integers(0, n_rows, size=(batch, lt), dtype=np.int64).astype(np.int32)
    vid_idx = np.full((batch, PHOTO_MAX), -1, dtype=np.int32)
    if ragged:
        H = hist_videos or max(1, lt // max(1, nv_max))
        S = segs_per_video or nv_max
        nt = np.zeros(batch, dtype=np.int64)
        for b in range(batch):
            h = int(rng.integers(1, H + 1))
            nt[b] = min(lt, int(rng.integers(1, S + 1, size=h).sum()))
        nv = rng.integers(min(2, nv_max), min(nv_max, PHOTO_MAX) + 1, size=batch)
        for b in range(batch):
            usr_idx[b, nt[b]:] = -1
    else:
        nv = np.full(batch, min(nv_max, PHOTO_MAX), dtype=np.int64)
    for b in range(batch):
        vid_idx[b, :nv[b]] = rng.integers(0, n_rows, size=int(nv[b]))
    gt = make_labels(rng, nv)
    return usr_idx, vid_idx, gt


def make_teacher_batch(table: np.ndarray, batch: int, lt: int, nv_max: int, seed: int, w_seed: int = 7, quantile: float = 0.25):
    """Planted-teacher labels for the AUC parity check: the viewer skips at the first candidate
    segment whose <w*, e_seg> + <w*, mean(history rows)> falls below a threshold (fixed w*, seed 7), so the
    per-segment skip AUC of a trained model is well above 0.5.  Ragged histories and candidates."""
    n_rows, din = table.shape
    usr_idx, vid_idx, _ = make_indices(batch, lt, nv_max, n_rows, seed=seed, ragged=True)
    w = np.random.default_rng(w_seed).standard_normal(din).astype(np.float32) / np.sqrt(din)
    proj = table.astype(np.float32) @ w                              # teacher score of every table row
    thr = np.quantile(proj, quantile)
    gt = np.full((batch, PHOTO_MAX), PAD_LABEL, dtype=np.int64)
    for b in range(batch):
        hist = usr_idx[b][usr_idx[b] >= 0]
        bias = 0.5 * (proj[hist].mean() if hist.size else 0.0)
        rows = vid_idx[b][vid_idx[b] >= 0]
        n = rows.size
        below = np.nonzero(proj[rows] + bias < thr)[0]
        v = int(below[0]) if below.size else n                     # first skipped segment (n = watched to the end)
        gt[b, :v] = 1
        if v < n:
            gt[b, v] = 0
            gt[b, v + 1:n] = -1
    return usr_idx, vid_idx, gt


def make_dense_batch(rng: np.random.Generator, B: int, Lt: int, din: int, full: bool = False):
    """Dense, already L1-normalised inputs in the form the reference driver hands
    to the model (main_for_seq_leave_earlystop_SegMM.py:271-284), ragged lengths."""
    usr = rng.standard_normal((B, Lt, din), dtype=np.float32)
    vid = rng.standard_normal((B, PHOTO_MAX, din), dtype=np.float32)
    nt = rng.integers(1, Lt + 1, size=B) if not full else np.full(B, Lt)
    nv = rng.integers(2, PHOTO_MAX + 1, size=B)
    nv[0] = PHOTO_MAX  # one fully-populated candidate
    usr_mask = np.arange(Lt)[None, :] < nt[:, None]
    vid_mask = np.arange(PHOTO_MAX)[None, :] < nv[:, None]
    usr = np.where(usr_mask[..., None], usr, 0).astype(np.float32)
    vid = np.where(vid_mask[..., None], vid, 0).astype(np.float32)
    usr = usr / (np.abs(usr).sum(-1, keepdims=True, dtype=np.float32) + np.float32(1e-6))
    vid = vid / (np.abs(vid).sum(-1, keepdims=True, dtype=np.float32) + np.float32(1e-6))
    gt = make_labels(rng, nv)
    return usr, usr_mask, vid, vid_mask, gt


def fill_state_dict(shapes: dict, seed: int = 42) -> dict:
    """Deterministic weights for full-size parity runs where committing a
    state_dict is too large: keys in sorted order, N(0, 0.02) for matrices,
    LayerNorm weight ~ 1 + N(0,0.02), biases N(0,0.02)."""
    rng = np.random.default_rng(seed)
    out = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        w = rng.standard_normal(shp, dtype=np.float32) * np.float32(0.02)
        if ("ln" in k.split(".")[-2] or "ln_" in k) and k.endswith(".weight") and len(shp) == 1:
            w = w + np.float32(1.0)
        if k.endswith("stage_mlp1.weight"):
            w = w * np.float32(5.0)
        out[k] = w
    return out
