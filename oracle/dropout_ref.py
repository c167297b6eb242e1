"""TEST INFRASTRUCTURE ONLY -- numpy twin of segmminterest_b200/csrc/dropout.cuh.

The reference draws its dropout masks (nn.Dropout(0.1) at models/encoder.py:21,145-150,163-164,187,198-202,386,472 and
kn_util/nn_utils/layers/mlp.py:15-23) from torch's global generator, which no other implementation can replay.  The CUDA
path uses a counter-based generator instead: keep(row, col) is a pure function of (key, row, col).  This file restates
that function on the CPU so that the oracle can run the reference's arithmetic with EXACTLY the masks the kernels use
(tests/test_gpu_dropout.py), which turns "training with dropout" into a deterministic parity test.
"""
from __future__ import annotations

import numpy as np

_M = np.array([0x9E3779B1, 0x85EBCA6B, 0xC2B2AE35, 0x27D4EB2F, 0x165667B1, 0xD3A2646D, 0xFD7046C5, 0xB55A4F09], dtype=np.uint64)
_A = np.array([0x7F4A7C15, 0x94D049BB, 0xBF58476D, 0x1CE4E5B9, 0x133111EB, 0x2545F491, 0x4CF5AD43, 0x2127599B], dtype=np.uint64)
_MASK = np.uint64(0xFFFFFFFF)


def _u32(x):
    return np.asarray(x, dtype=np.uint64) & _MASK


def mix32(x):
    x = _u32(x)
    x = x ^ (x >> np.uint64(16)); x = _u32(x * np.uint64(0x7FEB352D))
    x = x ^ (x >> np.uint64(15)); x = _u32(x * np.uint64(0x846CA68B))
    x = x ^ (x >> np.uint64(16))
    return x


def rowhash(key, rows):
    rows = np.asarray(rows, dtype=np.uint64)
    lo, hi = rows & _MASK, rows >> np.uint64(32)
    return mix32(np.uint64(key) + _u32(lo * np.uint64(0x9E3779B1)) + _u32(hi * np.uint64(0x85EBCA77)))


def keep_word(rowh, group, thr8):
    """uint32 words (as uint64 arrays): bit c = 1 iff column 32*group + c of that row is kept."""
    rowh, group = np.broadcast_arrays(_u32(rowh), _u32(group))
    h0 = mix32(rowh + _u32(group * np.uint64(0xC2B2AE3D)))
    lt = np.zeros_like(h0)
    eq = np.full_like(h0, 0xFFFFFFFF)
    for i in range(7, -1, -1):
        b = _u32(_u32(h0 * _M[i]) + _A[i])
        b = b ^ (b >> np.uint64(16))
        t = np.uint64(0xFFFFFFFF) if (thr8 >> i) & 1 else np.uint64(0)
        lt = lt | (eq & (~b & _MASK) & t)
        eq = eq & (~(b ^ t) & _MASK)
    return (~lt) & _MASK


def keep_mask(key, thr8, rows, ncols, group0=0):
    """bool [len(rows), ncols]: keep(row, col) for col = 0..ncols-1; the keep words are groups group0, group0+1, ..."""
    rows = np.asarray(rows, dtype=np.uint64).reshape(-1)
    if thr8 == 0:
        return np.ones((rows.size, ncols), dtype=bool)
    ng = (ncols + 31) // 32
    rh = rowhash(key, rows)[:, None]
    words = keep_word(rh, np.arange(group0, group0 + ng, dtype=np.uint64)[None, :], thr8)     # [R, ng]
    bits = (words[:, :, None] >> np.arange(32, dtype=np.uint64)[None, None, :]) & np.uint64(1)
    return bits.reshape(rows.size, ng * 32)[:, :ncols].astype(bool)


ATTN_BLOCK_GROUP_SHIFT = 20    # keep-word group of key k of key block blk: (blk << 20) + (k >> 5)


def attn_keep_mask(key, thr8, B, H, Lq, Lks):
    """bool [B, H, Lq, sum(Lks)]: masks of the logits dropout (encoder.py:145-150) over the concatenated key blocks.
    row id = (b*H + h)*Lq + q; columns of block blk start a fresh group range at blk << 20."""
    rows = np.arange(B * H * Lq, dtype=np.uint64)
    parts = [keep_mask(key, thr8, rows, Lk, group0=blk << ATTN_BLOCK_GROUP_SHIFT) for blk, Lk in enumerate(Lks)]
    return np.concatenate(parts, axis=1).reshape(B, H, Lq, sum(Lks))
