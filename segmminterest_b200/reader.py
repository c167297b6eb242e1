"""Drop-in for `BaseReaderSeq_SegMM` (utils/dataloader_SegMM.py:41-149): the reader on the driver's import surface
(main_for_seq_leave_earlystop_SegMM.py:6,219-220) that turns `SegMM/{train,dev,test}.csv` into the `*_his.csv` files with
per-interaction histories -- BASELINE config 1's input (SegMM_inter_sample.csv through the reference's own preparation).

Same constructor (`args.sep / path / data / dict_path / history_max`), same attributes (`data_df`, `user_input_dict`,
`all_df`, `n_users`, `n_items`, `max_users`, `max_items`, `history_max`), same files written, same cell format
(`str(np.array(...))`, what `to_csv` makes of the arrays the reference stores), so either implementation can read what the
other wrote.  Differences, all on the reference's failure paths:
  * the reference fills the history columns by chained assignment (`df[col][index] = value`, :108-110), which pandas >= 3
    (copy-on-write) silently ignores -- every history comes out empty there; here the columns are built directly
    (vectorised over each user's interaction sequence), i.e. what the reference computes under the pandas it was written for;
  * after building, the reference keeps numpy arrays in `data_df` and its dataset then fails on `.strip` (:289); here the
    freshly written files are read back, so `data_df` always holds the stringified form the dataset parses.
pandas does the file I/O, the merge and the sorts with the reference's own calls (same row order, ties included); this is
host-side preparation that runs once, not the hot path.
"""
from __future__ import annotations

import json
import os

import numpy as np
import pandas as pd

PHASES = ("train", "dev", "test")


class BaseReaderSeq_SegMM(object):
    @staticmethod
    def parse_data_args(parser):
        """utils/dataloader_SegMM.py:42-53"""
        parser.add_argument('--path', type=str, default='SegMM/', help='Input data dir.')
        parser.add_argument('--sep', type=str, default='\t', help='sep of csv file.')
        parser.add_argument('--data', type=str, default='inter')
        parser.add_argument('--dict_path', type=str, default='user_input_dict.json')
        parser.add_argument('--history_max', type=int, default=50, help='Maximum length of history.')
        return parser

    def __init__(self, args):
        self.sep = args.sep
        self.prefix = args.path
        self.data = args.data
        self.dict_path = args.dict_path
        with open(os.path.join(self.prefix, self.dict_path), 'r') as f:
            self.user_input_dict = json.load(f)
        self.history_max = args.history_max
        if not os.path.exists(self.prefix + 'train' + '_his.csv'):
            self._read_data()
            self._append_his_info()
            self._get_history()
            for key in PHASES:
                self.data_df[key] = self.data_df[key].sort_values(by='time_ms')
                self.data_df[key].to_csv(self.prefix + key + '_his.csv', sep=self.sep, index=False)
            built_users, built_items = self.n_users, self.n_items
            self._load_his()
            # a reader that has just built the files reports the counts of its own data (:146-147), not the constants
            self.n_users, self.n_items = built_users, built_items
        else:
            self._load_his()

    # ------------------------------------------------------------------ cached form (:64-86)
    def _load_his(self):
        self.data_df = dict()
        for key in PHASES:
            self.data_df[key] = pd.read_csv(self.prefix + key + '_his.csv', sep=self.sep).reset_index(drop=True).sort_values(by=['user_id', 'time_ms'])
        key_columns = ['user_id', 'video_id', 'time_ms', 'playing_time_x']
        self.all_df = pd.concat([self.data_df[key][key_columns] for key in PHASES])
        # the reference counts and then overrides with the full SegMM dataset's sizes (:78-79); the embedding tables of the
        # 'id' / 'both' input types are sized from these (main...SegMM.py:66,79)
        self.n_users = 1903
        self.n_items = 352494
        self.max_users = self.all_df['user_id'].max()
        self.max_items = self.all_df['video_id'].max()

    # ------------------------------------------------------------------ first run (:136-149, :112-134, :97-110)
    def _read_data(self):
        self.data_df = dict()
        for key in PHASES:
            self.data_df[key] = pd.read_csv(self.prefix + key + '.csv', sep=self.sep).reset_index(drop=True).sort_values(by=['user_id', 'time_ms'])
        key_columns = ['user_id', 'video_id', 'time_ms', 'playing_time']
        self.all_df = pd.concat([self.data_df[key][key_columns] for key in PHASES])
        self.n_users = len(self.all_df['user_id'].unique())
        self.n_items = len(self.all_df['video_id'].unique())

    def _append_his_info(self):
        """position of every interaction in its user's sequence (global order: time_ms, then user_id, stable) and the
        sequences themselves; the position is merged back into the three splits on (user_id, video_id, time_ms)."""
        sort_df = self.all_df.sort_values(by=['time_ms', 'user_id'], kind='mergesort')
        uid = sort_df['user_id'].to_numpy()
        # rank inside the user's sequence without a Python loop: stable sort by user keeps the global order per user
        order = np.argsort(uid, kind='stable')
        u_sorted = uid[order]
        starts = np.r_[0, np.flatnonzero(u_sorted[1:] != u_sorted[:-1]) + 1] if uid.size else np.empty(0, np.int64)
        counts = np.diff(np.r_[starts, uid.size])
        rank_sorted = np.arange(uid.size) - np.repeat(starts, counts)
        position = np.empty(uid.size, dtype=np.int64)
        position[order] = rank_sorted
        self._seq_user = u_sorted[starts] if uid.size else np.empty(0, np.int64)
        self._seq_start = starts
        self._seq_items = sort_df['video_id'].to_numpy()[order]
        self._seq_play = sort_df['playing_time'].to_numpy()[order]
        sort_df['position'] = position
        for key in PHASES:
            self.data_df[key] = pd.merge(left=self.data_df[key], right=sort_df, how='left', on=['user_id', 'video_id', 'time_ms'])
        del sort_df

    def _get_history(self):
        """history of interaction at position p of user u = the user's interactions [max(0, p - history_max), p)."""
        for key in PHASES:
            df = self.data_df[key]
            n = len(df)
            uid = df['user_id'].to_numpy()
            pos = df['position'].to_numpy().astype(np.int64)
            slot = np.searchsorted(self._seq_user, uid)
            base = self._seq_start[slot] if n else np.empty(0, np.int64)
            lo = base + np.maximum(pos - self.history_max, 0)
            hi = base + pos
            items = np.empty(n, dtype=object)
            play = np.empty(n, dtype=object)
            lengths = (hi - lo).astype(np.int64)
            for i in range(n):                         # one slice per row; the arrays themselves are views
                if lengths[i] > 0:
                    items[i] = self._seq_items[lo[i]:hi[i]]
                    play[i] = self._seq_play[lo[i]:hi[i]]
                else:
                    items[i] = None
                    play[i] = None
            df = df.copy()
            df['history_items'] = items
            df['history_playing'] = play
            df['history_lengths'] = lengths
            self.data_df[key] = df
