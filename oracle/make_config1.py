"""TEST INFRASTRUCTURE ONLY -- BASELINE config 1: `SegMM_inter_sample.csv` through the reference's OWN preparation and
loader code, dumped as tests/golden/config1.npz.  Build-container only (needs /root/reference).

    python -m oracle.make_config1

Pipeline (every step is the reference's code unless marked [restated]):
  1. [restated] data_process/get_data_SegMM_public.py:117-180 `split_inter_data` / `mask_ids` / `get_input_dict` -- that
     script runs at import time on files the repo does not ship (`data/inter_orignal_id_reverse.csv`), so its three
     functions are restated here on the sample CSV: per user with >= 100 interactions the first 80 become
     `user_input_dict` ("pid_sec" entries), the rest is split 81/9/10 with sklearn's train_test_split(random_state=2024);
     `second_map_{user,item}2id.json` number the ids from 1.
  2. the line-id map: one table row per (video, segment), segments = ceil(duration_ms / 5000).
  3. UNMODIFIED `BaseReaderSeq_SegMM` (utils/dataloader_SegMM.py:41-149) builds `{train,dev,test}_his.csv`.  Its
     `_get_history` fills the history columns by chained assignment, which pandas >= 3 (copy-on-write) ignores; it is run
     with `data_df[key]` wrapped in a proxy that gives `df[col][index] = value` the write-through semantics of the pandas
     the reference was written for (`_ChainedFrame` below).  Nothing else is patched.
  4. UNMODIFIED `BaseReaderSeq_SegMM` again (now reading the `_his.csv` files), `FrameDatasetSeq_SegMM(phase, shuffle=False)`
     and `DataCollator` (utils/dataloader_SegMM.py:186-382) on a synthetic table, `random.seed(42)` like the driver.
The fixture stores the batch in INDEX form (every dense feature row is matched back to its table row; the script asserts
the dense tensors equal table[rows] bit for bit), the scalars / masks / labels as they are, and the `_his.csv` files'
history columns + SHA-256 of the files, so that the GPU box can rebuild everything from the CSV text stored alongside.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import math
import os
import random
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CSV = os.path.join(ref_shim.REF_ROOT, "SegMM_inter_sample.csv")
N_BATCH = 256
DIN = 1024
TABLE_SEED = 1234


def split_inter_data(inter_df, seed=2024, n_input=80):
    """[restated] data_process/get_data_SegMM_public.py:117-160"""
    from sklearn.model_selection import train_test_split
    user_input_dict = {}
    parts = {"input": [], "train": [], "dev": [], "test": []}
    inter_df = inter_df.sort_values(by=["user_id", "time_ms"])
    for user_id, group in inter_df.groupby("user_id"):
        if len(group) < 100:
            continue
        input_df = group.iloc[:n_input]
        frames = []
        for _, row in input_df.iterrows():                       # get_input_dict (:101-112)
            playing = min(row["playing_time"], row["duration_ms"])
            frames += [f"{row['video_id']}_{int(i / 1000)}" for i in range(0, int(playing), 5000)]
        user_input_dict[int(user_id)] = frames
        rest = group.iloc[n_input:]
        train_valid, test = train_test_split(rest, test_size=0.1, random_state=seed)
        train, valid = train_test_split(train_valid, test_size=0.1, random_state=seed)
        for k, d in (("input", input_df), ("train", train), ("dev", valid), ("test", test)):
            parts[k].append(d)
    return user_input_dict, {k: pd.concat(v, ignore_index=True) for k, v in parts.items()}


class _ChainedFrame:
    """What `_get_history` (utils/dataloader_SegMM.py:97-110) touches of `self.data_df[key]`, with pandas < 2 semantics for
    `df[col][index] = value`: the column object handed out by __getitem__ is the one that is stored."""

    def __init__(self, df):
        self.df, self.cols = df, {}

    def __len__(self):
        return len(self.df)

    def __setitem__(self, col, value):
        n = len(self.df)
        if isinstance(value, list):
            arr = np.empty(n, dtype=object)
            for i, v in enumerate(value):
                arr[i] = v
        else:
            arr = np.full(n, value, dtype=np.int64)
        self.cols[col] = arr

    def __getitem__(self, col):
        return self.cols[col] if col in self.cols else self.df[col]

    def iterrows(self):
        return self.df.iterrows()

    def head(self, *a):
        return self.df.head(*a)

    def materialise(self):
        assert list(self.df.index) == list(range(len(self.df))), "labels must be positions (fresh merge result)"
        df = self.df.copy()
        for c, arr in self.cols.items():
            df[c] = arr
        return df


def build_files(td):
    """steps 1-3 inside directory `td` (becomes the cwd-relative 'SegMM/' of the reference)."""
    dl = ref_shim.load_dataloader()
    df = pd.read_csv(CSV)
    uid_dict, parts = split_inter_data(df)
    seg = os.path.join(td, "SegMM")
    os.makedirs(seg, exist_ok=True)
    for k in ("train", "dev", "test"):
        parts[k].to_csv(os.path.join(seg, f"{k}.csv"), sep="\t", index=False)
    json.dump(uid_dict, open(os.path.join(seg, "user_input_dict.json"), "w"))
    comb = pd.concat([parts[k] for k in ("input", "train", "dev", "test")], ignore_index=True)
    user2id = {int(k): v for v, k in enumerate(sorted(comb["user_id"].unique()), start=1)}       # mask_ids (:162-174)
    item2id = {int(k): v for v, k in enumerate(sorted(comb["video_id"].unique()), start=1)}
    json.dump(user2id, open(os.path.join(seg, "second_map_user2id.json"), "w"))
    json.dump(item2id, open(os.path.join(seg, "second_map_item2id.json"), "w"))
    # line-id map: rows in (video id, segment) order
    nseg = df.groupby("video_id")["duration_ms"].max().map(lambda d: len(range(0, int(d), 5000)))
    lineid, r = {}, 0
    for pid in sorted(nseg.index):
        for i in range(int(nseg[pid])):
            lineid[f"{pid}-{i}"] = r
            r += 1
    json.dump(lineid, open(os.path.join(td, "SegMM_photoidframeid2lineid.json"), "w"))
    args = SimpleNamespace(sep="\t", path="SegMM/", data="inter", dict_path="user_input_dict.json", history_max=50)

    # step 3: the unmodified reader; only `_get_history` sees the chained-assignment proxy
    class Reader(dl.BaseReaderSeq_SegMM):
        def _get_history(self):
            real = self.data_df
            self.data_df = {k: _ChainedFrame(v) for k, v in real.items()}
            try:
                dl.BaseReaderSeq_SegMM._get_history(self)
                self.data_df = {k: v.materialise() for k, v in self.data_df.items()}
            except Exception:
                self.data_df = real
                raise

    cwd = os.getcwd()
    os.chdir(td)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            Reader(args)
            reader = dl.BaseReaderSeq_SegMM(args)          # second construction: reads the *_his.csv it has just written
    finally:
        os.chdir(cwd)
    return dl, reader, lineid, uid_dict, user2id, item2id, args


def make_table(n_rows, din=DIN, seed=TABLE_SEED):
    """synthetic segment embeddings: numpy's PCG64 stream is stable across versions (NEP 19), unlike torch.randn"""
    return np.random.default_rng(seed).standard_normal((n_rows, din), dtype=np.float32)


def main():
    assert ref_shim.available() and os.path.isfile(CSV)
    with tempfile.TemporaryDirectory() as td:
        dl, reader, lineid, uid_dict, user2id, item2id, args = build_files(td)
        table = make_table(len(lineid))
        # make every table row identifiable from its first two values
        sig = table[:, 0].astype(np.float64) * 1e4 + table[:, 1]
        assert np.unique(sig).size == sig.size
        row_of = {s: i for i, s in enumerate(sig)}
        save = {}
        cwd = os.getcwd()
        os.chdir(td)
        try:
            for phase in ("train", "dev"):
                random.seed(42)
                np.random.seed(42)
                with contextlib.redirect_stdout(io.StringIO()):
                    ds = dl.FrameDatasetSeq_SegMM(corpus=reader, lineid_map=lineid, feat_memmap=table, phase=phase, shuffle=False,
                                                  verbose=False)
                    it = iter(ds)
                    samples = [next(it) for _ in range(N_BATCH)]
                    batch = dl.DataCollator()(samples)
                batch = {k: v.numpy() for k, v in batch.items()}

                def to_rows(x, m):
                    s = x[..., 0].astype(np.float64) * 1e4 + x[..., 1]
                    rows = np.full(m.shape, -1, dtype=np.int32)
                    for idx in zip(*np.nonzero(m)):
                        rows[idx] = row_of[s[idx]]
                    return rows

                usr_rows = to_rows(batch["user"], batch["user_mask"])
                vid_rows = to_rows(batch["photo"], batch["photo_mask"])
                for dense, rows in ((batch["user"], usr_rows), (batch["photo"], vid_rows)):   # the dense batch IS table[rows], zero-padded
                    ref = np.where((rows >= 0)[..., None], table[np.maximum(rows, 0)], np.float32(0))
                    assert np.array_equal(dense, ref)
                n_tok_total = []
                df = reader.data_df[phase]
                print(phase, "rows", len(df), "batch user tokens min/max", int(batch["user_mask"].sum(1).min()), int(batch["user_mask"].sum(1).max()),
                      "full-100 rows", int((batch["user_mask"].sum(1) == 100).sum()))
                save[f"{phase}/usr_rows"] = usr_rows
                save[f"{phase}/vid_rows"] = vid_rows
                for k in ("user_mask", "photo_mask", "label", "user_id", "photo_id", "user_identity_id", "photo_identity_id", "time_ms",
                          "play_time", "duration"):
                    save[f"{phase}/{k}"] = batch[k]
                save[f"{phase}/dtypes"] = json.dumps({k: str(v.dtype) for k, v in batch.items()})
                save[f"{phase}/keys"] = json.dumps(list(batch.keys()))
        finally:
            os.chdir(cwd)
        # the files the reference reader wrote (text, so that the GPU box needs nothing else) + their hashes
        seg = os.path.join(td, "SegMM")
        for k in ("train", "dev", "test"):
            raw = open(os.path.join(seg, f"{k}.csv"), "rb").read()
            his = open(os.path.join(seg, f"{k}_his.csv"), "rb").read()
            save[f"files/{k}.csv"] = np.frombuffer(raw, dtype=np.uint8)
            save[f"files/{k}_his.csv"] = np.frombuffer(his, dtype=np.uint8)
            save[f"sha256/{k}_his.csv"] = hashlib.sha256(his).hexdigest()
        save["files/user_input_dict.json"] = np.frombuffer(open(os.path.join(seg, "user_input_dict.json"), "rb").read(), dtype=np.uint8)
        save["files/second_map_user2id.json"] = np.frombuffer(open(os.path.join(seg, "second_map_user2id.json"), "rb").read(), dtype=np.uint8)
        save["files/second_map_item2id.json"] = np.frombuffer(open(os.path.join(seg, "second_map_item2id.json"), "rb").read(), dtype=np.uint8)
        save["files/SegMM_photoidframeid2lineid.json"] = np.frombuffer(open(os.path.join(td, "SegMM_photoidframeid2lineid.json"), "rb").read(),
                                                                        dtype=np.uint8)
        save["meta"] = json.dumps(dict(n_rows=len(lineid), din=DIN, table_seed=TABLE_SEED, n_batch=N_BATCH, history_max=50,
                                       n_users=int(reader.n_users), n_items=int(reader.n_items), pandas=pd.__version__,
                                       numpy=np.__version__, table_check=float(table[:64].astype(np.float64).sum())))
        np.savez_compressed(os.path.join(OUT, "config1.npz"), **save)
        print("config1.npz", os.path.getsize(os.path.join(OUT, "config1.npz")) >> 10, "KiB; table rows", len(lineid))


if __name__ == "__main__":
    main()
