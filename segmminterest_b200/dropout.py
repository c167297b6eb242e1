"""Host side of the counter-based dropout (csrc/dropout.cuh): probability quantisation and per-site keys.

The reference applies nn.Dropout(0.1) at five kinds of sites per backbone (models/encoder.py:386,472 embedding;
:145-150 attention logits, BEFORE the 1/sqrt(dh) scale; :163-164 attention output projection; kn_util/nn_utils/layers/
mlp.py:21-22 after the FFN's GELU; encoder.py:198-202 FFN output).  SegFormerX is built without a dropout argument
(main_for_seq_leave_earlystop_SegMM.py:88-104), so p is always the constructor default 0.1 while the model is in
train() mode and 0 under eval().  Each site gets its own 32-bit key = mix(seed, call counter, site id); kernels
regenerate the mask from (key, row, col) in forward and backward, so no mask is ever stored."""
from __future__ import annotations

import warnings

_WARNED = set()
_M32 = 0xFFFFFFFF


def mix32(x: int) -> int:
    x &= _M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & _M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & _M32
    x ^= x >> 16
    return x


def quantise(p: float):
    """(thr8, scale): the element is dropped iff its 8-bit uniform number is < thr8; survivors are scaled by the reciprocal
    of the realised keep probability."""
    if not 0.0 <= p < 1.0:
        raise ValueError(f"dropout probability must be in [0, 1), got {p}")
    thr8 = min(255, int(round(p * 256.0)))
    if p > 0.0:
        realised = thr8 / 256.0
        if thr8 == 0:
            raise ValueError(f"dropout p={p} is below the 1/512 resolution of the keep words and would silently turn dropout off; "
                             "use p = 0 or p >= 1/512")
        if abs(realised - p) > 0.02 * p and p not in _WARNED:
            _WARNED.add(p)
            warnings.warn(f"dropout p={p} is realised as {thr8}/256 = {realised:.4f} (keep words quantise p to 1/256; survivors are "
                          "scaled by the realised keep probability, so the expectation stays exact)", stacklevel=2)
    return thr8, 256.0 / (256.0 - thr8)


def site_key(seed: int, call: int, site: int) -> int:
    """32-bit key of dropout site `site` in forward call number `call` of a run seeded with `seed`."""
    return mix32(mix32(seed & _M32) + mix32(((seed >> 32) & _M32) ^ 0x5BD1E995) * 3 + (call & _M32) * 0x9E3779B1 + mix32(site + 0x632BE5AB))


class DropSite:
    """What a kernel wrapper needs for one site: (key, thr8, scale)."""
    __slots__ = ("key", "thr8", "scale")

    def __init__(self, key: int, thr8: int, scale: float):
        self.key, self.thr8, self.scale = key, thr8, scale

    def __repr__(self):
        return f"DropSite(key=0x{self.key:08x}, thr8={self.thr8}, scale={self.scale:.6f})"
