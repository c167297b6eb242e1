"""BASELINE config 1: `SegMM_inter_sample.csv` through the reference's own preparation and loader
(tests/golden/config1.npz, written by oracle/make_config1.py from the UNMODIFIED reference reader / dataset / collator).

CPU: our `BaseReaderSeq_SegMM` writes byte-identical `*_his.csv` files; the vectorised index form of the reference batch
(256 interactions, every user above the 100-token cap, i.e. through `random.sample`) is reproduced exactly.
GPU: `DeviceFrameLoader` (and the drop-in `FrameDatasetSeq_SegMM` + `DataLoader`) yields the reference's twelve-key batch
bit for bit; one fp32 and one bf16 training step on that batch against the oracle at the benchmark's model size.
"""
import hashlib
import json
import os
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "config1.npz")
PHASE_KEYS = ("user_mask", "photo_mask", "label", "user_id", "photo_id", "user_identity_id", "photo_identity_id", "time_ms", "play_time",
              "duration")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _materialise(gold, td, with_his=True):
    seg = os.path.join(td, "SegMM")
    os.makedirs(seg, exist_ok=True)
    names = ["train.csv", "dev.csv", "test.csv", "user_input_dict.json", "second_map_user2id.json", "second_map_item2id.json"]
    if with_his:
        names += ["train_his.csv", "dev_his.csv", "test_his.csv"]
    for n in names:
        with open(os.path.join(seg, n), "wb") as f:
            f.write(gold["files/" + n].tobytes())
    with open(os.path.join(td, "SegMM_photoidframeid2lineid.json"), "wb") as f:
        f.write(gold["files/SegMM_photoidframeid2lineid.json"].tobytes())
    return SimpleNamespace(sep="\t", path=os.path.join(td, "SegMM") + "/", data="inter", dict_path="user_input_dict.json", history_max=50)


def _table(gold):
    meta = json.loads(str(gold["meta"]))
    t = np.random.default_rng(meta["table_seed"]).standard_normal((meta["n_rows"], meta["din"]), dtype=np.float32)
    assert abs(float(t[:64].astype(np.float64).sum()) - meta["table_check"]) < 1e-9, "numpy's generator stream changed"
    return t


def test_reader_builds_the_reference_his_files(gold, tmp_path):
    """utils/dataloader_SegMM.py:55-134: history construction + the files it caches, byte for byte"""
    from segmminterest_b200.reader import BaseReaderSeq_SegMM
    args = _materialise(gold, str(tmp_path), with_his=False)
    reader = BaseReaderSeq_SegMM(args)
    for k in ("train", "dev", "test"):
        raw = open(os.path.join(str(tmp_path), "SegMM", f"{k}_his.csv"), "rb").read()
        assert hashlib.sha256(raw).hexdigest() == str(gold[f"sha256/{k}_his.csv"]), f"{k}_his.csv differs from the reference's"
        assert raw == gold[f"files/{k}_his.csv"].tobytes()
    # built in this run: the counts of its own data (dataloader_SegMM.py:146-147); a second construction reads the cache
    n_u, n_i = reader.n_users, reader.n_items
    assert n_u == len(reader.all_df["user_id"].unique()) and n_i == len(reader.all_df["video_id"].unique())
    again = BaseReaderSeq_SegMM(args)
    assert (again.n_users, again.n_items) == (1903, 352494)          # the reference's hard-coded sizes (:78-79)
    for k in ("train", "dev", "test"):
        assert again.data_df[k].equals(reader.data_df[k])
    hl = again.data_df["train"]["history_lengths"].to_numpy()
    assert hl.max() == 50 and (hl == 0).sum() < len(hl)                # histories are really there (pandas >= 3 leaves the reference's empty)


def test_index_form_of_the_reference_batch(gold, tmp_path):
    """FrameDatasetSeq_SegMM._getitem + DataCollator (:270-382) vs SegmentIndex / HostFrameIndex on the first 256 rows of
    train and dev: row ids (order included: over-long users go through random.sample), masks, labels, scalars, dtypes."""
    from segmminterest_b200.loader import HostFrameIndex
    from segmminterest_b200.reader import BaseReaderSeq_SegMM
    args = _materialise(gold, str(tmp_path))
    reader = BaseReaderSeq_SegMM(args)
    lineid = json.load(open(os.path.join(str(tmp_path), "SegMM_photoidframeid2lineid.json")))
    u2i = json.loads(gold["files/second_map_user2id.json"].tobytes())
    i2i = json.loads(gold["files/second_map_item2id.json"].tobytes())
    n = json.loads(str(gold["meta"]))["n_batch"]
    for phase in ("train", "dev"):
        hf = HostFrameIndex(reader, lineid, phase, u2i, i2i)
        random.seed(42)                                              # the driver's seed (main...SegMM.py:26-28)
        usr, vid = hf.index_batch(np.arange(n))
        assert np.array_equal(vid, gold[f"{phase}/vid_rows"])
        assert np.array_equal(usr, gold[f"{phase}/usr_rows"]), "history tokens (or their random.sample order) differ from the reference"
        sc = hf.scalars(np.arange(n))
        dtypes = json.loads(str(gold[f"{phase}/dtypes"]))
        for k in PHASE_KEYS:
            if k.endswith("_mask"):
                continue
            assert np.array_equal(sc[k], gold[f"{phase}/{k}"]), k
            assert str(sc[k].dtype) == dtypes[k], k
        assert np.array_equal(usr >= 0, gold[f"{phase}/user_mask"]) and np.array_equal(vid >= 0, gold[f"{phase}/photo_mask"])
    assert int((gold["train/user_mask"].sum(1) == 100).sum()) > 0       # the > 100-token path is exercised


@pytest.mark.gpu
def test_device_loader_yields_the_reference_batch(gold, tmp_path):
    """the twelve-key batch dict of the reference on the device: feature rows bit-exact copies of table rows, zero pad rows,
    masks, labels, scalars -- through the drop-in names (FrameDatasetSeq_SegMM + DataLoader + DataCollator)."""
    from segmminterest_b200.dataset import DataCollator, DataLoader, FrameDatasetSeq_SegMM
    from segmminterest_b200.reader import BaseReaderSeq_SegMM
    args = _materialise(gold, str(tmp_path))
    table = _table(gold)
    lineid = json.load(open(os.path.join(str(tmp_path), "SegMM_photoidframeid2lineid.json")))
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        reader = BaseReaderSeq_SegMM(SimpleNamespace(sep="\t", path="SegMM/", data="inter", dict_path="user_input_dict.json", history_max=50))
        n = json.loads(str(gold["meta"]))["n_batch"]
        for phase in ("train", "dev"):
            ds = FrameDatasetSeq_SegMM(corpus=reader, lineid_map=lineid, feat_memmap=table, phase=phase, shuffle=False,
                                       do_scale_image_to_01=True, image_resize=True, verbose=False)
            dl = DataLoader(ds, n, collate_fn=DataCollator())
            random.seed(42)
            batch = next(iter(dl))
            keys = json.loads(str(gold[f"{phase}/keys"]))
            assert [k for k in batch if k not in ("usr_idx", "vid_idx")] == keys
            batch = {k: v.cuda() for k, v in batch.items()}           # main...SegMM.py:271
            dt = json.loads(str(gold[f"{phase}/dtypes"]))
            for k in keys:
                assert str(batch[k].dtype).replace("torch.", "") == dt[k], k
            for name, rows_key in (("user", "usr_rows"), ("photo", "vid_rows")):
                rows = gold[f"{phase}/{rows_key}"]
                ref = np.where((rows >= 0)[..., None], table[np.maximum(rows, 0)], np.float32(0))
                assert np.array_equal(batch[name].cpu().numpy(), ref), name
            for k in PHASE_KEYS:
                assert np.array_equal(batch[k].cpu().numpy(), gold[f"{phase}/{k}"]), k
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_config1_training_step_vs_oracle(gold, precision):
    """one forward + backward of the benchmark-sized model (d 512, 16 heads, 6 layers, Din 1024, Lt 100) on the first 12
    interactions of the reference's config-1 batch against the CPU oracle: fp32 1e-4, bf16 2e-2 (logits, loss, whole gradient)."""
    from oracle import gather_oracle, mmi_oracle
    from segmminterest_b200 import synth
    from segmminterest_b200.model import build_model, reference_state_shapes
    from test_gpu_model import make_args
    dev = torch.device("cuda:0")
    B = 12
    table = _table(gold)
    usr_idx, vid_idx = gold["train/usr_rows"][:B], gold["train/vid_rows"][:B]
    gt = gold["train/label"][:B]
    u, um = gather_oracle.gather_dense(table, usr_idx)
    c, cm = gather_oracle.gather_dense(table, vid_idx)
    u, c = gather_oracle.l1_normalise(u), gather_oracle.l1_normalise(c)
    args = make_args(d_model=512, nhead=16, num_layers_enc=6, mmi_precision=precision)
    model = build_model(args, din=1024, max_usr_len=100).cuda().eval()
    shapes = reference_state_shapes(512, 6, 1024, 100, 40)
    sd = {k: torch.from_numpy(v) for k, v in synth.fill_state_dict(shapes, 42).items()}
    model.load_state_dict(sd)
    live = mmi_oracle.live_param_names(list(sd.keys()), 6)
    osd = {k: v.clone().requires_grad_(k in live) for k, v in sd.items()}
    o = mmi_oracle.forward(osd, torch.from_numpy(u), torch.from_numpy(um), torch.from_numpy(c), torch.from_numpy(cm), torch.from_numpy(gt.copy()),
                           nhead=16, num_layers=6)
    o["loss"].backward()
    z = torch.zeros(B, dtype=torch.long, device=dev)
    out = model(usr_image=torch.from_numpy(u).to(dev), usr_id=z, usr_mask=torch.from_numpy(um).to(dev), vid_image=torch.from_numpy(c).to(dev),
                vid_id=z, vid_mask=torch.from_numpy(cm).to(dev), gt=torch.from_numpy(gt.copy()).to(dev), mode="train")
    out["loss"].backward()
    tol = 1e-4 if precision == "fp32" else 2e-2

    def rel(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

    assert rel(out["logits"].detach().cpu().numpy(), o["logits"].detach().numpy()) < tol
    assert abs(out["loss"].item() - o["loss"].item()) < tol * abs(o["loss"].item())
    named = dict(model.named_parameters())
    num = den = 0.0
    for k in live:
        g, r = named[k].grad.detach().cpu().numpy().astype(np.float64), osd[k].grad.numpy().astype(np.float64)
        num += float(((g - r) ** 2).sum())
        den += float((r ** 2).sum())
        if precision == "fp32":
            assert rel(g, r) < tol, k
    assert (num / den) ** 0.5 < tol, f"whole-gradient error {(num / den) ** 0.5:.3e}"
