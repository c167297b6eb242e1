"""CPU tests of the counter-based dropout's host side and of its numpy twin (oracle/dropout_ref.py): the properties the
CUDA kernels rely on.  The device generator itself is held to the twin bit for bit in tests/test_gpu_dropout.py."""
import numpy as np
import pytest
import torch

from oracle import dropout_ref, mmi_oracle
from segmminterest_b200 import dropout as host


def test_quantise_matches_the_reference_probability():
    thr8, scale = host.quantise(0.1)                       # nn.Dropout(0.1) everywhere in the reference
    assert thr8 == 26 and abs(scale - 256.0 / 230.0) < 1e-12
    assert host.quantise(0.0) == (0, 1.0)
    assert host.quantise(0.5) == (128, 2.0)
    with pytest.raises(ValueError):
        host.quantise(1.0)


def test_site_keys_are_distinct_per_site_call_and_seed():
    keys = {host.site_key(s, c, k) for s in (0, 42, 1 << 40) for c in range(1, 20) for k in range(0, 1100, 7)}
    assert len(keys) == 3 * 19 * len(range(0, 1100, 7))    # no collision among ~9 000 (seed, call, site) triples
    assert all(0 <= k < 1 << 32 for k in keys)
    assert host.mix32(0) == int(dropout_ref.mix32(0)) and host.mix32(0xDEADBEEF) == int(dropout_ref.mix32(0xDEADBEEF))


@pytest.mark.parametrize("thr8", [1, 26, 128, 255])
def test_keep_rate_and_independence(thr8):
    m = dropout_ref.keep_mask(0xC0FFEE, thr8, np.arange(2048), 512)
    want = 1.0 - thr8 / 256.0
    assert abs(m.mean() - want) < 4 * np.sqrt(want * (1 - want) / m.size) + 1e-4
    if thr8 in (26, 128):
        for a, b in ((m[:, :-1], m[:, 1:]), (m[:-1], m[1:]), (m[:, :-32], m[:, 32:])):     # neighbours in a word, in a column, across words
            assert abs(np.corrcoef(a.ravel(), b.ravel())[0, 1]) < 5e-3
        other = dropout_ref.keep_mask(0xC0FFEF, thr8, np.arange(2048), 512)                # next key: unrelated mask
        assert abs(np.corrcoef(m.ravel(), other.ravel())[0, 1]) < 5e-3


def test_masks_are_pure_functions_of_key_row_and_column():
    rows = np.array([0, 1, 5, (1 << 33) + 7], dtype=np.uint64)
    a = dropout_ref.keep_mask(7, 26, rows, 96)
    b = dropout_ref.keep_mask(7, 26, rows[::-1].copy(), 96)[::-1]
    assert np.array_equal(a, b)                                                            # row order does not matter
    assert np.array_equal(a[:, :40], dropout_ref.keep_mask(7, 26, rows, 40))               # nor the tensor width
    assert np.array_equal(a[:, 32:64], dropout_ref.keep_mask(7, 26, rows, 32, group0=1))   # group offset = column offset / 32
    assert dropout_ref.keep_mask(7, 0, rows, 96).all()                                     # thr8 = 0: off


def test_attention_mask_layout_matches_the_kernels_indexing():
    """row = (b*H + h)*Lq + q; the columns of key block i start a fresh group range at i << 20."""
    B, H, Lq, Lks = 2, 3, 5, [40, 70]
    m = dropout_ref.attn_keep_mask(99, 26, B, H, Lq, Lks)
    assert m.shape == (B, H, Lq, 110)
    b, h, q = 1, 2, 3
    row = np.array([(b * H + h) * Lq + q], dtype=np.uint64)
    assert np.array_equal(m[b, h, q, :40], dropout_ref.keep_mask(99, 26, row, 40, group0=0)[0])
    assert np.array_equal(m[b, h, q, 40:], dropout_ref.keep_mask(99, 26, row, 70, group0=1 << 20)[0])


def test_oracle_hook_sites_and_eval_mode():
    """every site of the reference is visited exactly once per layer and side, in the reference's order; drop=None is eval()"""
    torch.manual_seed(0)
    from segmminterest_b200 import synth
    from segmminterest_b200.model import reference_state_shapes
    shapes = reference_state_shapes(64, 3, 16, 8, 40)
    sd = {k: torch.from_numpy(v) for k, v in synth.fill_state_dict(shapes, 1).items()}
    rng = np.random.default_rng(0)
    usr, um, vid, vm, gt = synth.make_dense_batch(rng, 2, 8, 16)
    seen = []

    def drop(kind, tower, layer, side, x, blocks=None):
        seen.append((kind, tower, layer, side, tuple(x.shape), None if blocks is None else tuple(blocks)))
        return x

    args = (sd, torch.from_numpy(usr), torch.from_numpy(um), torch.from_numpy(vid), torch.from_numpy(vm), torch.from_numpy(gt))
    a = mmi_oracle.forward(*args, nhead=2, num_layers=3, drop=drop)
    b = mmi_oracle.forward(*args, nhead=2, num_layers=3)
    assert torch.equal(a["logits"], b["logits"])                                           # identity hook == eval mode
    E, A, O, M1, M2 = (mmi_oracle.DROP_EMB, mmi_oracle.DROP_ATTN, mmi_oracle.DROP_ATTN_OUT, mmi_oracle.DROP_MLP1, mmi_oracle.DROP_MLP2)
    kinds = [(s[0], s[2], s[3]) for s in seen]
    assert kinds == [(E, 63, 0), (E, 63, 1),
                     (A, 0, 0), (A, 0, 1), (O, 0, 1), (O, 0, 0), (M1, 0, 0), (M2, 0, 0), (M1, 0, 1), (M2, 0, 1),   # full layer
                     (A, 1, 0), (O, 1, 0), (M1, 1, 0), (M2, 1, 0)]                                                  # layer N-2: candidate side only
    assert seen[2][4] == (2, 2, 40, 48) and seen[2][5] == (40, 8)                          # logits [B,H,Lq,Lv+Lt], blocks (v2v, t2v)
