// Counter-based dropout masks (Philox4x32-7, Salmon et al. SC'11 constants). One call yields 128 random bits = eight
// 16-bit lanes = the keep decisions of 8 consecutive columns of one row, so forward and backward regenerate the same mask
// from (seed, tensor coordinates) without storing it. The reference draws its masks from torch's own Philox stream
// (nn.Dropout, /root/reference/model.py:69, 91, 116, 151, 171); that stream cannot be reproduced outside torch, so
// parity is exact only at dropout = 0 and statistical (same keep probability, inverted scaling) above it.
#pragma once
#include <cstdint>

namespace ta {

struct Philox4 {
  uint32_t w[4];
};

__host__ __device__ __forceinline__ Philox4 philox4x32_7(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                                                         uint32_t c3) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint64_t p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0, n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    c1 = static_cast<uint32_t>(p1);
    c3 = static_cast<uint32_t>(p0);
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 o;
  o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
  return o;
}

// keep decisions of columns 8*col8 .. 8*col8+7 of `row` in plane `plane` (group / head index); kind separates the
// GEMM-epilogue planes (0) from the attention planes (1). Bit e of the result = column e is kept.
__host__ __device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint32_t kind, uint32_t plane, uint32_t row, uint32_t col8,
                                                           uint32_t thresh16) {
  const Philox4 r = philox4x32_7(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), col8, row, plane, kind);
  uint32_t m = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) m |= (((r.w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu) >= thresh16 ? 1u : 0u) << e;
  return m;
}

// threshold of the 16-bit lanes for drop probability p (p in [0, 1)): lane < thresh <=> dropped
__host__ __device__ __forceinline__ uint32_t dropout_thresh16(float p) {
  const float t = p * 65536.f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : static_cast<uint32_t>(t));
}

}  // namespace ta
