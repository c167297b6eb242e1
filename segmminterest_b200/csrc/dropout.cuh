// Counter-based dropout shared by every kernel that owns an nn.Dropout site of the reference
// (models/encoder.py:21,145-150,163-164,187,198-202,386,472; kn_util/nn_utils/layers/mlp.py:15-23).
//
// There is no random state: whether element (row, col) of a site survives is a pure function of
// (key, row, col), so forward and backward -- and kernels that walk the same tensor in different
// orientations (attention dq: thread = query, dk/dv: thread = key) -- regenerate identical masks
// and nothing is stored in HBM.  `key` mixes the step seed and the site id on the host.
//
// The unit of generation is a KEEP WORD: 32 consecutive columns of one row (columns 32 g .. 32 g + 31, "group" g).
// Eight pseudo-random bit planes b7..b0 give every bit position an 8-bit uniform number u; the element is dropped
// iff u < thr8, evaluated for all 32 positions at once by a bit-sliced comparator.  The drop probability is
// therefore thr8 / 256 (nn.Dropout(0.1) -> 26 / 256 = 0.1016) and the survivors are scaled by 256 / (256 - thr8),
// the reciprocal of the REALISED keep probability, so E[dropout(x)] = x holds exactly.
// Cost: ~60 integer instructions per 32 elements, against ~9 per element for a per-element hash.
// tests/test_dropout.py holds the numpy twin (oracle/dropout_ref.py) to these functions bit for bit.
#pragma once
#include <stdint.h>

#include "../../include/mmi_b200.h"

namespace mmi {

struct DropParams {
  uint32_t key;    // mixes seed, step and site (host)
  uint32_t thr8;   // drop iff u < thr8;  0 = dropout off
  float scale;     // 256 / (256 - thr8)
};
__host__ __device__ __forceinline__ DropParams make_drop(const mmi_dropout& d) { return DropParams{d.key, d.thr8, d.scale}; }
__host__ __device__ __forceinline__ DropParams drop_off() { return DropParams{0u, 0u, 1.0f}; }

__host__ __device__ __forceinline__ uint32_t drop_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
// per-row state: computed once per thread and row
__host__ __device__ __forceinline__ uint32_t drop_rowhash(uint32_t key, uint64_t row) {
  return drop_mix32(key + static_cast<uint32_t>(row) * 0x9E3779B1u + static_cast<uint32_t>(row >> 32) * 0x85EBCA77u);
}
// bit c of the result = 1 iff the 8-bit number of position c (bit-sliced over the eight planes of h0) is >= thr8
__host__ __device__ __forceinline__ uint32_t drop_compare_planes(uint32_t h0, uint32_t thr8) {
  constexpr uint32_t M[8] = {0x9E3779B1u, 0x85EBCA6Bu, 0xC2B2AE35u, 0x27D4EB2Fu, 0x165667B1u, 0xD3A2646Du, 0xFD7046C5u, 0xB55A4F09u};
  constexpr uint32_t A[8] = {0x7F4A7C15u, 0x94D049BBu, 0xBF58476Du, 0x1CE4E5B9u, 0x133111EBu, 0x2545F491u, 0x4CF5AD43u, 0x2127599Bu};
  uint32_t lt = 0u, eq = 0xffffffffu;
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    uint32_t b = h0 * M[i] + A[i];
    b ^= b >> 16;
    const uint32_t t = 0u - ((thr8 >> i) & 1u);
    lt |= eq & ~b & t;
    eq &= ~(b ^ t);
  }
  return ~lt;
}
// bit c of the result = 1 iff element (row, 32 * group + c) is KEPT
__host__ __device__ __forceinline__ uint32_t drop_keep_word(uint32_t rowh, uint32_t group, uint32_t thr8) {
  const uint32_t h0 = drop_mix32(rowh + group * 0xC2B2AE3Du);
  // the reference's p = 0.1 (thr8 = 26 = 0b00011010) gets a constant-folded comparator: 11 logic ops instead of 24
  if (thr8 == 26u) return drop_compare_planes(h0, 26u);
  return drop_compare_planes(h0, thr8);
}

}  // namespace mmi
