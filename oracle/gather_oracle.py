"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's host-side
gather.  Nothing under
``segmminterest_b200/`` may import it.

Parity status: PINNED against the reference's own ``FrameDatasetSeq_SegMM`` +
``DataCollator`` (utils/dataloader_SegMM.py) run through ``oracle/ref_shim.py``
on a small fixture -> ``tests/golden/gather_small.npz``.
"""
from __future__ import annotations

import numpy as np

PHOTO_MAX = 40   # utils/dataloader_SegMM.py:198
USER_MAX = 100   # utils/dataloader_SegMM.py:199


def n_segments(duration_ms) -> int:
    """utils/dataloader_SegMM.py:213-215: one segment per started 5000 ms."""
    return len(range(0, int(duration_ms), 5000))


def parse_int_list(s) -> list:
    """'[ 1 1 0 -1]' -> [1,1,0,-1]  (utils/dataloader_SegMM.py:240-242,289-293)."""
    return [int(t) for t in str(s).strip("[").strip("]").split(" ") if t.strip()]


def pad_labels(label_str, max_length=PHOTO_MAX, pad_value=-2):
    """utils/dataloader_SegMM.py:240-249."""
    lab = parse_int_list(label_str)[:max_length]
    return np.array(lab + [pad_value] * (max_length - len(lab)), dtype=np.int64)


def candidate_rows(photo_id, duration_ms, lineid_map) -> list:
    """utils/dataloader_SegMM.py:301-309: every segment must exist."""
    rows = []
    for i in range(n_segments(duration_ms)):
        key = f"{photo_id}-{i}"
        if key not in lineid_map:
            raise ValueError(f"No key in lineid dict: {key}")
        rows.append(lineid_map[key])
    return rows


def history_rows(user_id, history_items, history_playing, lineid_map, user_input_dict) -> list:
    """utils/dataloader_SegMM.py:319-343: watched segments of past videos that
    exist in the map, then the user's `user_input_dict` entries ("pid_sec")."""
    rows = []
    for pid, playing in zip(history_items, history_playing):
        for i in range(n_segments(playing)):
            key = f"{pid}-{i}"
            if key in lineid_map:
                rows.append(lineid_map[key])
    for pf in user_input_dict.get(str(user_id), []):
        pid, sec = pf.split("_")
        key = f"{pid}-{sec}"
        if key in lineid_map:
            rows.append(lineid_map[key])
    return rows


def gather_pad_mask(table: np.ndarray, rows, max_length: int):
    """utils/dataloader_SegMM.py:251-268 for len(rows) <= max_length (the
    > max_length branch is a random sub-sample and is handled by the caller)."""
    out = np.zeros((max_length, table.shape[1]), dtype=table.dtype)
    mask = np.zeros(max_length, dtype=bool)
    n = len(rows)
    if n:
        out[:n] = table[np.asarray(rows, dtype=np.int64)]
    mask[:n] = True
    return out, mask


def gather_dense(table: np.ndarray, idx: np.ndarray):
    """Index form used by the CUDA path: idx [B, L] int32, -1 = pad.
    Returns (features [B,L,Din] with zero pad rows, mask [B,L] bool)."""
    m = idx >= 0
    out = table[np.where(m, idx, 0)]
    out = np.where(m[..., None], out, np.zeros((), dtype=table.dtype))
    return out.astype(table.dtype), m


def l1_normalise(x: np.ndarray) -> np.ndarray:
    """main_for_seq_leave_earlystop_SegMM.py:272-273 in float32:
    x / (sum|x| + 1e-6)."""
    x = x.astype(np.float32)
    n = np.abs(x).sum(-1, keepdims=True, dtype=np.float32)
    return x / (n + np.float32(1e-6))


def getitem_port(row: dict, lineid_map: dict, user_input_dict: dict, table: np.ndarray, user2id: dict, item2id: dict) -> dict:
    """One sample of FrameDatasetSeq_SegMM._getitem (utils/dataloader_SegMM.py:281-362) restated with the functions above:
    per-segment f-string keys + dict lookups + one table-row copy per hit, np.vstack, zero pad, mask; users with more than
    100 tokens are sub-sampled with random.sample exactly like the reference.  `row` holds the columns of a *_his.csv row.
    Used by bench.py's loader baseline (samples/s of the reference's Python loader on the box's host cores)."""
    import random
    hist_items = parse_int_list(row["history_items"]) if row["history_lengths"] > 0 else []
    hist_play = parse_int_list(row["history_playing"]) if row["history_lengths"] > 0 else []
    out = {"play_time": int(row["playing_time_x"] / 5000), "duration": int(row["duration_ms"] / 5000)}
    urows = history_rows(row["user_id"], hist_items, hist_play, lineid_map, user_input_dict)
    feats = np.array([table[r, :] for r in urows])                       # one row read per hit, like the memmap reads (:327,338)
    if feats.shape[0] > USER_MAX:
        feats = feats[random.sample(range(feats.shape[0]), USER_MAX)]
    user = np.zeros((USER_MAX, table.shape[1]), dtype=table.dtype)
    user[:feats.shape[0]] = feats
    umask = np.zeros(USER_MAX, dtype=bool)
    umask[:feats.shape[0]] = True
    out.update(user=user, user_mask=umask, user_id=row["user_id"], user_identity_id=int(user2id[str(row["user_id"])]))
    vrows = candidate_rows(row["video_id"], row["duration_ms"], lineid_map)
    photo, pmask = gather_pad_mask(table, vrows, PHOTO_MAX)
    out.update(photo=photo, photo_mask=pmask, photo_id=row["video_id"], photo_identity_id=int(item2id[str(row["video_id"])]),
               time_ms=row["time_ms"], label=pad_labels(row["label_1D"]))
    return out


def collate_port(samples: list) -> dict:
    """DataCollator.__call__ (utils/dataloader_SegMM.py:370-382) without the prints: np.stack per key"""
    return {k: np.stack([s_[k] for s_ in samples]) for k in samples[0]}
