"""the step immediately BEFORE the hot path.  The reference builds every batch with per-segment
Python string formatting and dict lookups (`f"{photo_id}-{frame_i}" in lineid_map`, one np.memmap row read per hit,
utils/dataloader_SegMM.py:301-350).  `SegmentIndex` parses `SegMM_photoidframeid2lineid.json` ("pid-seg" -> table row)
and `user_input_dict.json` ("pid_sec" lists per user) ONCE into sorted integer arrays, after which a whole batch of
interactions is turned into the int32 row-id tensors the device gather consumes (`DeviceGather`, -1 = pad) with a few
vectorised numpy calls -- same rows, same order, same padding as the reference loader.

Host-side index arithmetic only (integers): no embedding row is touched here; the gather itself is the CUDA kernel.
"""
from __future__ import annotations

import random as _pyrandom

import numpy as np

PHOTO_MAX = 40    # utils/dataloader_SegMM.py:198
USER_MAX = 100    # utils/dataloader_SegMM.py:199
SEG_MS = 5000     # utils/dataloader_SegMM.py:213-215: one segment per started 5000 ms


def n_segments(ms) -> np.ndarray:
    """len(range(0, int(ms), 5000)) for arrays (utils/dataloader_SegMM.py:213-215)."""
    ms = np.asarray(ms).astype(np.int64)
    return np.where(ms > 0, (ms + SEG_MS - 1) // SEG_MS, 0)


def _ramp(counts: np.ndarray) -> np.ndarray:
    """[0..c0-1, 0..c1-1, ...] for counts [c0, c1, ...]."""
    counts = np.asarray(counts, dtype=np.int64)
    total = int(counts.sum())
    starts = np.cumsum(counts) - counts
    return np.arange(total, dtype=np.int64) - np.repeat(starts, counts)


def parse_label(label_str, max_length=PHOTO_MAX, pad_value=-2) -> np.ndarray:
    """'[ 1 1 0 -1]' -> padded int64 [max_length]  (utils/dataloader_SegMM.py:240-249)."""
    lab = [int(t) for t in str(label_str).strip("[").strip("]").split(" ") if t.strip()][:max_length]
    return np.array(lab + [pad_value] * (max_length - len(lab)), dtype=np.int64)


class SegmentIndex:
    def __init__(self, lineid_map: dict, user_input_dict: dict | None = None):
        n = len(lineid_map)
        pid = np.empty(n, dtype=np.int64)
        seg = np.empty(n, dtype=np.int64)
        row = np.empty(n, dtype=np.int64)
        for i, (key, r) in enumerate(lineid_map.items()):
            p, s = key.rsplit("-", 1)
            pid[i], seg[i], row[i] = int(p), int(s), int(r)
        if n and (seg.min() < 0 or pid.min() < 0):
            raise ValueError("SegmentIndex: negative photo id / segment index in the line-id map")
        self.seg_cap = int(seg.max()) + 1 if n else 1
        if n and int(pid.max()) > (np.iinfo(np.int64).max // self.seg_cap) - 1:
            raise ValueError("SegmentIndex: photo ids too large for the composite key")
        key = pid * self.seg_cap + seg
        order = np.argsort(key, kind="stable")
        self._key = key[order]
        self._row = row[order].astype(np.int32)
        if n > 1 and np.any(self._key[1:] == self._key[:-1]):
            raise ValueError("SegmentIndex: duplicate 'pid-seg' keys")
        self.n_rows = int(row.max()) + 1 if n else 0
        # user extras ("pid_sec" -> key "pid-sec", skipped when missing: utils/dataloader_SegMM.py:331-339), resolved once
        self._user_ptr = {}
        extra_rows = []
        pos = 0
        for uid, items in (user_input_dict or {}).items():
            if items:
                ps = np.array([it.split("_") for it in items], dtype=np.int64).reshape(-1, 2)
                r = self.lookup(ps[:, 0], ps[:, 1])
                r = r[r >= 0]
            else:
                r = np.empty(0, dtype=np.int32)
            self._user_ptr[str(uid)] = (pos, pos + r.size)
            extra_rows.append(r)
            pos += r.size
        self._extra = np.concatenate(extra_rows).astype(np.int32) if extra_rows else np.empty(0, dtype=np.int32)

    # ------------------------------------------------------------------------------------------ lookups
    def lookup(self, pid, seg) -> np.ndarray:
        """table row of ("pid-seg") for arrays of ids; -1 where the key is not in the map."""
        pid = np.asarray(pid, dtype=np.int64)
        seg = np.asarray(seg, dtype=np.int64)
        ok = (seg >= 0) & (seg < self.seg_cap) & (pid >= 0)
        q = np.where(ok, pid, 0) * self.seg_cap + np.where(ok, seg, 0)
        if self._key.size == 0:
            return np.full(q.shape, -1, dtype=np.int32)
        i = np.searchsorted(self._key, q)
        i = np.minimum(i, self._key.size - 1)
        hit = ok & (self._key[i] == q)
        return np.where(hit, self._row[i], -1).astype(np.int32)

    def user_extra(self, user_id) -> np.ndarray:
        if str(user_id) not in self._user_ptr:
            raise KeyError(str(user_id))                     # the reference indexes user_input_dict[str(uid)] directly
        a, b = self._user_ptr[str(user_id)]
        return self._extra[a:b]

    # ------------------------------------------------------------------------------------------ batches
    def candidate_idx(self, video_id, duration_ms) -> np.ndarray:
        """[B, 40] int32 rows of the candidate videos' segments (utils/dataloader_SegMM.py:301-314); every segment
        i < ceil(duration / 5000) must exist (ValueError like the reference); more than 40 segments are sub-sampled
        by _pad_feature_list (:251-257) -- not reproducible across RNGs, so it is refused here."""
        video_id = np.asarray(video_id, dtype=np.int64)
        nseg = n_segments(duration_ms)
        if np.any(nseg > PHOTO_MAX):
            raise ValueError("candidate with more than 40 segments: the reference sub-samples at random (np.random.choice)")
        B = video_id.shape[0]
        pid = np.repeat(video_id, nseg)
        seg = _ramp(nseg)
        rows = self.lookup(pid, seg)
        if np.any(rows < 0):
            j = int(np.argmax(rows < 0))
            raise ValueError(f"No key in lineid dict: {pid[j]}-{seg[j]}")
        out = np.full((B, PHOTO_MAX), -1, dtype=np.int32)
        out[np.repeat(np.arange(B), nseg), seg] = rows
        return out

    def history_idx(self, user_id, history_items, history_playing, rng: np.random.Generator | None = None) -> np.ndarray:
        """[B, 100] int32 rows of the users' histories (utils/dataloader_SegMM.py:319-350): for each past video the
        watched segments i < ceil(playing / 5000) that exist in the map, in order, then the user's `user_input_dict`
        rows; more than 100 tokens are sub-sampled WITHOUT order: `random.sample` on Python's global generator like the
        reference when rng is None, otherwise from the numpy Generator `rng`."""
        B = len(user_id)
        # zip(history_items, history_playing) in the reference (:322) stops at the shorter of the two lists
        n_hist = np.array([min(len(h), len(p)) for h, p in zip(history_items, history_playing)], dtype=np.int64)
        vids = (np.concatenate([np.asarray(h, dtype=np.int64).reshape(-1)[:n] for h, n in zip(history_items, n_hist)])
                if n_hist.sum() else np.empty(0, np.int64))
        play = (np.concatenate([np.asarray(p, dtype=np.int64).reshape(-1)[:n] for p, n in zip(history_playing, n_hist)])
                if n_hist.sum() else np.empty(0, np.int64))
        nseg = n_segments(play)
        sample_of_vid = np.repeat(np.arange(B), n_hist)
        tok_sample = np.repeat(sample_of_vid, nseg)
        rows = self.lookup(np.repeat(vids, nseg), _ramp(nseg))
        keep = rows >= 0
        tok_sample, rows = tok_sample[keep], rows[keep]
        n_tok = np.bincount(tok_sample, minlength=B).astype(np.int64)
        extras = [self.user_extra(u) for u in user_id]
        n_extra = np.array([e.size for e in extras], dtype=np.int64)
        total = n_tok + n_extra
        out = np.full((B, USER_MAX), -1, dtype=np.int32)
        small = total <= USER_MAX
        # history tokens of the samples that fit: position = rank inside the sample (tokens are already sample-ordered)
        pos = _ramp(n_tok)
        m = small[tok_sample]
        out[tok_sample[m], pos[m]] = rows[m]
        for b in np.nonzero(n_extra > 0)[0]:
            if small[b]:
                out[b, n_tok[b]:n_tok[b] + n_extra[b]] = extras[b]
        if not small.all():
            # the reference draws `random.sample(range(n), 100)` from Python's global generator (:346), which the driver seeds
            # (main...SegMM.py:26-28); samples are visited in batch order, one draw per over-long user, so with rng=None a
            # seeded run picks exactly the reference's tokens in the reference's order (tests/test_config1.py)
            starts = np.cumsum(n_tok) - n_tok
            for b in np.nonzero(~small)[0]:
                allrows = np.concatenate([rows[starts[b]:starts[b] + n_tok[b]], extras[b]])
                if rng is None:
                    pick = np.asarray(_pyrandom.sample(range(allrows.size), USER_MAX), dtype=np.int64)
                else:
                    pick = rng.choice(allrows.size, USER_MAX, replace=False)
                out[b] = allrows[pick]
        return out

    def batch(self, user_id, video_id, duration_ms, history_items, history_playing, label_1d, rng=None) -> dict:
        """One collated batch in index form: what DataCollator (utils/dataloader_SegMM.py:370-382) returns with the
        `photo` / `user` feature tensors replaced by row ids (masks follow from idx >= 0)."""
        return dict(vid_idx=self.candidate_idx(video_id, duration_ms),
                    usr_idx=self.history_idx(user_id, history_items, history_playing, rng),
                    label=np.stack([parse_label(s) for s in label_1d]) if len(label_1d) else np.empty((0, PHOTO_MAX), np.int64))
