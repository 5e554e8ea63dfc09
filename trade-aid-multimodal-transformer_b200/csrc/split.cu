// Operand splitting for the "bf16x3" precision mode.
//
// The north star holds the logits to 2e-3 of the reference's fp32 forward (/root/reference/model.py:380-402); a single bf16
// rounding of every GEMM operand gives ~5e-3 (so does torch.autocast on the reference). Writing an fp32 value x as
// hi + lo with hi = bf16(x), lo = bf16(x - hi) keeps 16 mantissa bits, and the product of two such values is
//   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo           (lo*lo ~ 2^-18 relative is dropped)
// which is ONE ordinary bf16 GEMM over a reduction dimension three times as long:
//   [A_hi | A_lo | A_hi] x [B_hi | B_hi | B_lo]^T
// — the unmodified tcgen05 kernel (gemm.cu), fp32 accumulation in TMEM, 3x the FLOPs. This file holds the two element
// kernels that mode needs next to it: the split itself and strided fp32 <-> bf16 conversions (the attention kernels keep
// bf16 operands and outputs in every mode; the emulation showed they are not the source of the error).
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "common.cuh"

namespace ta {

struct View4 {        // [G1][G0][R][C] view with unit inner stride
  long long g1, g0, r;
};

// dst slot k (k = 0..terms-1) holds hi or lo of src: pattern bit k set -> lo. Slots lie side by side along the columns
// (concat_rows = 0: dst[.., r, k*C + c]) or stacked along the rows (concat_rows = 1: dst[.., k*R + r, c]).
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ src, View4 sv, __nv_bfloat16* __restrict__ dst, View4 dv, int G1, int G0, int R,
                  int C, int terms, int lo_mask, int concat_rows) {
  pdl_prologue();
  const int C4 = (C + 3) >> 2;
  const long long total = (long long)G1 * G0 * R * C4;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C4) * 4;
    long long t = e / C4;
    const int r = (int)(t % R); t /= R;
    const int g0 = (int)(t % G0);
    const int g1 = (int)(t / G0);
    const float* sp = src + g1 * sv.g1 + g0 * sv.g0 + (long long)r * sv.r + c;
    float x[4];
    const int n = C - c < 4 ? C - c : 4;
    if (n == 4 && ((reinterpret_cast<uintptr_t>(sp) & 15) == 0)) {
      const float4 v = *reinterpret_cast<const float4*>(sp);
      x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = i < n ? sp[i] : 0.f;
    }
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      hi[i] = __float2bfloat16_rn(x[i]);
      lo[i] = __float2bfloat16_rn(x[i] - __bfloat162float(hi[i]));     // exact difference (Sterbenz-like: same binade)
    }
    __nv_bfloat16* dbase = dst + g1 * dv.g1 + g0 * dv.g0;
    for (int k = 0; k < terms; ++k) {
      const __nv_bfloat16* v = (lo_mask >> k) & 1 ? lo : hi;
      __nv_bfloat16* dp = concat_rows ? dbase + ((long long)k * R + r) * dv.r + c : dbase + (long long)r * dv.r + (long long)k * C + c;
      if (n == 4 && ((reinterpret_cast<uintptr_t>(dp) & 7) == 0)) {
        uint2 u;
        __nv_bfloat162 a = __halves2bfloat162(v[0], v[1]), b = __halves2bfloat162(v[2], v[3]);
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(dp) = u;
      } else {
        for (int i = 0; i < n; ++i) dp[i] = v[i];
      }
    }
  }
}

template <typename TS, typename TD>
__device__ __forceinline__ TD conv(TS v);
template <> __device__ __forceinline__ float conv<float, float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 conv<float, __nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ float conv<__nv_bfloat16, float>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ __nv_bfloat16 conv<__nv_bfloat16, __nv_bfloat16>(__nv_bfloat16 v) { return v; }

template <typename TS, typename TD>
__global__ void __launch_bounds__(256)
convert_kernel(const TS* __restrict__ src, View4 sv, TD* __restrict__ dst, View4 dv, int G1, int G0, int R, int C) {
  pdl_prologue();
  const long long total = (long long)G1 * G0 * R * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    long long t = e / C;
    const int r = (int)(t % R); t /= R;
    const int g0 = (int)(t % G0);
    const int g1 = (int)(t / G0);
    dst[g1 * dv.g1 + g0 * dv.g0 + (long long)r * dv.r + c] = conv<TS, TD>(src[g1 * sv.g1 + g0 * sv.g0 + (long long)r * sv.r + c]);
  }
}

static int elem_grid(long long work) {
  long long b = ceil_div(work, 256);
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace ta

using namespace ta;

extern "C" int32_t ta_split_bf16(const float* src, int64_t s_g1, int64_t s_g0, int64_t s_r, void* dst_bf16, int64_t d_g1,
                                 int64_t d_g0, int64_t d_r, int32_t G1, int32_t G0, int32_t R, int32_t C, int32_t terms,
                                 int32_t lo_mask, int32_t concat_rows, void* stream) {
  TA_CHECK_ARG(src && dst_bf16, "ta_split_bf16: null pointer");
  TA_CHECK_ARG(G1 > 0 && G0 > 0 && R > 0 && C > 0, "ta_split_bf16: non-positive size");
  TA_CHECK_ARG(terms >= 1 && terms <= 3 && lo_mask >= 0 && lo_mask < (1 << terms), "ta_split_bf16: terms %d / lo_mask %d", (int)terms,
               (int)lo_mask);
  const long long work = (long long)G1 * G0 * R * ((C + 3) / 4);
  launch_kernel(split_bf16_kernel, dim3(elem_grid(work)), dim3(256), 0, (cudaStream_t)stream, src, View4{s_g1, s_g0, s_r},
                (__nv_bfloat16*)dst_bf16, View4{d_g1, d_g0, d_r}, (int)G1, (int)G0, (int)R, (int)C, (int)terms, (int)lo_mask,
                (int)(concat_rows != 0));
  TA_CHECK_LAUNCH();
  return 0;
}

extern "C" int32_t ta_convert(const void* src, int32_t src_is_bf16, int64_t s_g1, int64_t s_g0, int64_t s_r, void* dst,
                              int32_t dst_is_bf16, int64_t d_g1, int64_t d_g0, int64_t d_r, int32_t G1, int32_t G0, int32_t R,
                              int32_t C, void* stream) {
  TA_CHECK_ARG(src && dst, "ta_convert: null pointer");
  TA_CHECK_ARG(G1 > 0 && G0 > 0 && R > 0 && C > 0, "ta_convert: non-positive size");
  const long long work = (long long)G1 * G0 * R * C;
  const dim3 grid(elem_grid(work)), block(256);
  auto st = (cudaStream_t)stream;
  const View4 sv{s_g1, s_g0, s_r}, dv{d_g1, d_g0, d_r};
  if (!src_is_bf16 && dst_is_bf16)
    launch_kernel(convert_kernel<float, __nv_bfloat16>, grid, block, 0, st, (const float*)src, sv, (__nv_bfloat16*)dst, dv, (int)G1, (int)G0, (int)R, (int)C);
  else if (src_is_bf16 && !dst_is_bf16)
    launch_kernel(convert_kernel<__nv_bfloat16, float>, grid, block, 0, st, (const __nv_bfloat16*)src, sv, (float*)dst, dv, (int)G1, (int)G0, (int)R, (int)C);
  else if (!src_is_bf16)
    launch_kernel(convert_kernel<float, float>, grid, block, 0, st, (const float*)src, sv, (float*)dst, dv, (int)G1, (int)G0, (int)R, (int)C);
  else
    launch_kernel(convert_kernel<__nv_bfloat16, __nv_bfloat16>, grid, block, 0, st, (const __nv_bfloat16*)src, sv, (__nv_bfloat16*)dst, dv, (int)G1, (int)G0, (int)R, (int)C);
  TA_CHECK_LAUNCH();
  return 0;
}
