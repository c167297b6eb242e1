"""Drop-in for the reference's data loader trio on the hot path (import surface 8b):
`FrameDatasetSeq_SegMM` + `DataLoader` + `DataCollator` of utils/dataloader_SegMM.py:186-382.

The reference yields one interaction at a time: ~140 string-keyed dict lookups and memmap row reads per sample in
Python, `np.stack` per key in the collator, then `{k: v.cuda()}` in the driver (main...SegMM.py:271).  `DeviceFrameLoader`
keeps the embedding table resident in HBM, resolves a whole batch to int32 row ids with the vectorised `SegmentIndex`
(one-time parse of the line-id map and of the stringified history columns), and produces the batch dict -- same twelve
keys, shapes, dtypes and values as `DataCollator` -- directly on the device with `mmi_gather_l1norm_fwd`
(gather + zero pad + mask; optionally the L1 normalisation of main...SegMM.py:272-273 fused in).

Row order: `shuffle=True` draws the permutation from numpy's global RNG exactly like the reference
(`np.random.shuffle(samples_BN)`, :278).  Users with more than 100 history tokens are sub-sampled with `random.sample`
in the reference (:346); so are they here (Python's global generator, same draw order: a seeded run reproduces the
reference's batches bit for bit, tests/test_config1.py) unless a numpy Generator is passed as `rng`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from .index import PHOTO_MAX, USER_MAX, SegmentIndex, parse_label


def _parse_int_list(s) -> np.ndarray:
    """'[101 102 105]' -> int64 array (utils/dataloader_SegMM.py:289-293); non-strings (NaN for empty histories) -> empty."""
    if not isinstance(s, str):
        return np.empty(0, dtype=np.int64)
    return np.array([int(x) for x in s.strip("[").strip("]").split(" ") if x], dtype=np.int64)


class HostFrameIndex:
    """The host half of the loader: the interaction table of one phase parsed once into arrays, and the int32 row-id form
    of any set of interactions (integers only -- no embedding row is touched, no device needed)."""

    def __init__(self, corpus, lineid_map: dict | None, phase: str = "train", user2id: dict | None = None, item2id: dict | None = None,
                 rng: np.random.Generator | None = None, index: SegmentIndex | None = None):
        self.index = index or SegmentIndex(lineid_map, corpus.user_input_dict)
        self.rng = rng
        df = corpus.data_df[phase]
        self.n = len(df)
        col = {c: df[c].to_numpy() for c in ("user_id", "video_id", "time_ms", "duration_ms", "playing_time_x", "label_1D",
                                             "history_items", "history_playing", "history_lengths")}
        self.user_id = col["user_id"].astype(np.int64)
        self.video_id = col["video_id"].astype(np.int64)
        self.time_ms = col["time_ms"].astype(np.int64)
        self.duration_ms = col["duration_ms"]
        self.playing = col["playing_time_x"]
        has_hist = col["history_lengths"].astype(np.int64) > 0                      # :288: the strings are only parsed then
        empty = np.empty(0, dtype=np.int64)
        self.hist_items = [_parse_int_list(s) if h else empty for s, h in zip(col["history_items"], has_hist)]
        self.hist_play = [_parse_int_list(s) if h else empty for s, h in zip(col["history_playing"], has_hist)]
        self.label = np.stack([parse_label(s) for s in col["label_1D"]]) if self.n else np.empty((0, PHOTO_MAX), np.int64)
        # int(x / 5000) like the reference (:296), not floor division (they differ for negative values only)
        self.play_time = np.array([int(p / 5000) for p in self.playing], dtype=np.int64)
        self.duration = np.array([int(d / 5000) for d in self.duration_ms], dtype=np.int64)
        u2i, i2i = user2id or {}, item2id or {}
        self.user_identity = np.array([int(u2i[str(u)]) for u in self.user_id], dtype=np.int64) if user2id is not None else None
        self.photo_identity = np.array([int(i2i[str(p)]) for p in self.video_id], dtype=np.int64) if item2id is not None else None

    def index_batch(self, sel: np.ndarray):
        """int32 row ids (-1 = pad) of the interactions `sel`: usr_idx [b,100], vid_idx [b,40] -- the form TrainStep and
        InferenceScorer consume directly."""
        vid = self.index.candidate_idx(self.video_id[sel], self.duration_ms[sel])
        usr = self.index.history_idx(self.user_id[sel], [self.hist_items[i] for i in sel], [self.hist_play[i] for i in sel], self.rng)
        # the reference's _pad_feature_list indexes feature.shape[1] of an EMPTY list for a user without a single token
        # (IndexError at :259); surface it the same way instead of training on an all-pad history
        if np.any((usr >= 0).sum(1) == 0):
            raise IndexError("user with no history token and no user_input_dict row (the reference raises IndexError at "
                             "utils/dataloader_SegMM.py:259)")
        return usr, vid

    def scalars(self, sel: np.ndarray) -> dict:
        """the per-interaction scalars and labels of the reference batch (host arrays, reference dtypes)"""
        out = {"play_time": self.play_time[sel], "duration": self.duration[sel], "user_id": self.user_id[sel],
               "photo_id": self.video_id[sel], "time_ms": self.time_ms[sel], "label": self.label[sel]}
        if self.user_identity is not None:
            out["user_identity_id"] = self.user_identity[sel]
        if self.photo_identity is not None:
            out["photo_identity_id"] = self.photo_identity[sel]
        return out


class DeviceFrameLoader(HostFrameIndex):
    def __init__(self, corpus, lineid_map: dict, table: torch.Tensor, phase: str = "train", batch_size: int = 512, shuffle: bool = False,
                 user2id: dict | None = None, item2id: dict | None = None, normalise: bool = False, out_dtype=torch.float32,
                 rng: np.random.Generator | None = None, index: SegmentIndex | None = None):
        if not table.is_cuda:
            raise _lib.MMIError("DeviceFrameLoader needs the embedding table in device memory; there is no CPU fallback")
        super().__init__(corpus, lineid_map, phase, user2id, item2id, rng, index)
        self.table = table.contiguous()
        if self.index.n_rows > self.table.shape[0]:
            # the reference's feat_memmap[line_id] raises IndexError for a row past the table; the gather kernel would read
            # such an id as padding, so a table loaded with the wrong n_rows must be refused here
            raise IndexError(f"line-id map refers to row {self.index.n_rows - 1} but the table has {self.table.shape[0]} rows")
        self.batch_size, self.shuffle, self.normalise, self.out_dtype = int(batch_size), shuffle, normalise, out_dtype

    def __len__(self):
        return (self.n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = np.arange(self.n)
        if self.shuffle:
            np.random.shuffle(order)
        dev = self.table.device
        D = self.table.shape[1]
        for a in range(0, self.n, self.batch_size):
            sel = order[a:a + self.batch_size]
            b = sel.size
            usr_idx, vid_idx = self.index_batch(sel)
            u_d, v_d = torch.from_numpy(usr_idx).to(dev, non_blocking=True), torch.from_numpy(vid_idx).to(dev, non_blocking=True)
            user = torch.empty(b, USER_MAX, D, device=dev, dtype=self.out_dtype)
            photo = torch.empty(b, PHOTO_MAX, D, device=dev, dtype=self.out_dtype)
            um = torch.empty(b, USER_MAX, device=dev, dtype=torch.uint8)
            pm = torch.empty(b, PHOTO_MAX, device=dev, dtype=torch.uint8)
            ops.gather_l1norm(self.table, u_d, user, um, self.normalise)
            ops.gather_l1norm(self.table, v_d, photo, pm, self.normalise)

            def t(x):
                return torch.from_numpy(np.ascontiguousarray(x)).to(dev, non_blocking=True)

            batch = {"play_time": t(self.play_time[sel]), "duration": t(self.duration[sel]),
                     "user": user, "user_mask": um.view(torch.bool), "user_id": t(self.user_id[sel])}
            if self.user_identity is not None:
                batch["user_identity_id"] = t(self.user_identity[sel])
            batch.update({"photo": photo, "photo_mask": pm.view(torch.bool), "photo_id": t(self.video_id[sel])})
            if self.photo_identity is not None:
                batch["photo_identity_id"] = t(self.photo_identity[sel])
            batch.update({"time_ms": t(self.time_ms[sel]), "label": t(self.label[sel]),
                          "usr_idx": u_d, "vid_idx": v_d})     # extra keys: the index form, for TrainStep / InferenceScorer
            yield batch
