"""The reference's loader trio under its own names and call signatures (utils/dataloader_SegMM.py:186-382), for the driver's
`load_data` (main_for_seq_leave_earlystop_SegMM.py:42-58):

    ds = FrameDatasetSeq_SegMM(corpus=reader, lineid_map=map, feat_memmap=memmap, phase='train', shuffle=True, ...)
    dl = DataLoader(ds, args.train_batch_size, collate_fn=DataCollator())
    for batch in dl: batch = {k: v.cuda() for k, v in batch.items()}      # twelve keys, already on the device here

`FrameDatasetSeq_SegMM` uploads `feat_memmap` to HBM once (shared by the train / dev / test datasets of one run), parses the
line-id map once (`SegmentIndex`) and hands batches to `DeviceFrameLoader`, which gathers on the device
(`mmi_gather_l1norm_fwd`).  `DataLoader(dataset, batch_size, collate_fn=...)` returns that loader for our datasets and falls
back to torch's DataLoader for anything else; `DataCollator` collates per-sample dicts exactly like the reference (without
its per-key prints) for callers that iterate the dataset sample by sample.
"""
from __future__ import annotations

import json

import numpy as np
import torch
import torch.utils.data

from . import _lib
from .index import SegmentIndex
from .loader import DeviceFrameLoader

_TABLES = {}     # id(feat_memmap) -> (feat_memmap, device table): the three datasets of a run share one upload
_INDEXES = {}    # (id(lineid_map), id(user_input_dict)) -> (lineid_map, user_input_dict, SegmentIndex)


def resident_table(feat_memmap, device=None, chunk_rows: int = 1 << 16) -> torch.Tensor:
    """the embedding rows of `feat_memmap` ([N, D] float32 / float64 array or memmap) as ONE float32 tensor in HBM"""
    if isinstance(feat_memmap, torch.Tensor) and feat_memmap.is_cuda:
        return feat_memmap
    if not torch.cuda.is_available():
        raise _lib.MMIError("FrameDatasetSeq_SegMM (b200) gathers on the GPU: a CUDA device is required, there is no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    key = (id(feat_memmap), str(dev))
    hit = _TABLES.get(key)
    if hit is not None and hit[0] is feat_memmap:
        return hit[1]
    n, d = feat_memmap.shape
    out = torch.empty(n, d, dtype=torch.float32, device=dev)
    for a in range(0, n, chunk_rows):
        b = min(n, a + chunk_rows)
        out[a:b].copy_(torch.from_numpy(np.array(feat_memmap[a:b])).to(dev).to(torch.float32))
    _TABLES[key] = (feat_memmap, out)
    return out


def shared_index(lineid_map, user_input_dict) -> SegmentIndex:
    key = (id(lineid_map), id(user_input_dict))
    hit = _INDEXES.get(key)
    if hit is not None and hit[0] is lineid_map and hit[1] is user_input_dict:
        return hit[2]
    idx = SegmentIndex(lineid_map, user_input_dict)
    _INDEXES[key] = (lineid_map, user_input_dict, idx)
    return idx


class FrameDatasetSeq_SegMM(torch.utils.data.IterableDataset):
    """Constructor keywords of utils/dataloader_SegMM.py:187-190 (the image options are accepted and unused, as in the
    reference, whose features are precomputed)."""

    def __init__(self, corpus, lineid_map: dict, feat_memmap, shuffle=True, phase='train', image_resize=True,
                 target_hw_shape=(224, 224), do_scale_image_to_01=True, verbose=True):
        super().__init__()
        self.corpus, self.lineid_map, self.feat_memmap = corpus, lineid_map, feat_memmap
        self.shuffle, self.phase, self.verbose = shuffle, phase, verbose
        self.photo_max_image = 40
        self.user_max_image = 100
        self.df = corpus.data_df[phase]
        self.user_input_dict = corpus.user_input_dict
        with open('SegMM/second_map_user2id.json', 'r', encoding='utf-8') as f:      # relative paths, like the reference (:207-210)
            self.user2id = json.load(f)
        with open('SegMM/second_map_item2id.json', 'r', encoding='utf-8') as f:
            self.item2id = json.load(f)
        self._loaders = {}

    def loader(self, batch_size: int, normalise: bool = False) -> DeviceFrameLoader:
        key = (int(batch_size), bool(normalise))
        if key not in self._loaders:
            self._loaders[key] = DeviceFrameLoader(self.corpus, self.lineid_map, resident_table(self.feat_memmap), phase=self.phase,
                                                   batch_size=batch_size, shuffle=self.shuffle, user2id=self.user2id, item2id=self.item2id,
                                                   normalise=normalise, index=shared_index(self.lineid_map, self.user_input_dict))
        return self._loaders[key]

    def __len__(self):
        return len(self.df)

    def __iter__(self):
        """one dict of numpy values per interaction, keys and dtypes of the reference's `_getitem` (:270-362)"""
        for batch in self.loader(256):
            host = {k: v.cpu().numpy() for k, v in batch.items() if k not in ("usr_idx", "vid_idx")}
            for i in range(next(iter(host.values())).shape[0]):
                yield {k: v[i] for k, v in host.items()}


class DataCollator(object):
    """utils/dataloader_SegMM.py:370-382 without the per-key shape / time prints"""

    def __call__(self, batch):
        assert len(batch)
        return {k: torch.from_numpy(np.stack([np.asarray(item[k]) for item in batch])) for k in batch[0].keys()}


def DataLoader(dataset, batch_size=1, *args, collate_fn=None, **kwargs):
    """`torch.utils.data.DataLoader` as the driver calls it (positional batch size, `collate_fn=DataCollator()`); for the
    datasets of this module the result is the device loader (same batches, already in HBM)."""
    if isinstance(dataset, FrameDatasetSeq_SegMM):
        return _LazyLoader(dataset, batch_size)
    return torch.utils.data.DataLoader(dataset, batch_size, *args, collate_fn=collate_fn, **kwargs)


class _LazyLoader:
    """The table upload and the index build happen at the first iteration, not inside `load_data`, so building the loaders
    costs nothing (and needs no device) -- iterating without CUDA raises MMIError."""

    def __init__(self, dataset: FrameDatasetSeq_SegMM, batch_size: int):
        self.dataset, self.batch_size = dataset, int(batch_size)

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        return iter(self.dataset.loader(self.batch_size))


class BaseReaderSeq_SegMM_sampled(object):
    """imported unconditionally by the driver (main...SegMM.py:7) and used only with `--eval_cold sampleData`; the reference
    does not ship `utils/dataloader_SegMM_sampled.py`, so there is nothing to reproduce"""

    @staticmethod
    def parse_data_args(parser):
        raise NotImplementedError("--eval_cold sampleData: utils/dataloader_SegMM_sampled.py is not part of the reference tree")

    def __init__(self, *a, **k):
        raise NotImplementedError("--eval_cold sampleData: utils/dataloader_SegMM_sampled.py is not part of the reference tree")


class FrameDatasetSeq_SegMM_sampled(BaseReaderSeq_SegMM_sampled):
    pass
