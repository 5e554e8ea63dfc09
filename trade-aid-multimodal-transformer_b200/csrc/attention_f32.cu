// K3p — fp32 causal attention forward for the "bf16x3" precision mode.
//
// The tensor-core attention kernels keep bf16 operands (q, k, v, P) and a bf16 output in every mode. With the GEMMs on split
// operands that rounding is ALL that is left of the logits error against the reference's fp32 forward
// (/root/reference/model.py:60-73, 136-159): 1.0e-3 normwise at c3 widths and 1.9e-3 at c4 - inside the north star's 2e-3, but
// with no margin - against 1.2e-5 with exact attention (measured on the CPU emulation, o / v / P / q,k contributing
// 8.4 / 5.2 / 4.9 / 1.3 e-4). This kernel is the exact one: plain fp32 FMAs, no tensor cores, used only by the precision mode
// (whose step is several times the throughput mode's anyway). Same contract as ta_attention_fwd: per source an independent
// causal softmax(q k^T * hs^-0.5) v, outputs summed; log-sum-exp per source saved for the backward kernel.
//
// One warp per query row: lanes split the keys for the scores (each lane one key row, q broadcast from shared memory), the
// probabilities of the row go through shared memory, then lanes split the head dimension for p V.
#include <cmath>

#include "../../include/tradeaid_b200.h"
#include "common.cuh"

namespace ta {

constexpr int AF32_MAX_SRC = 8;
constexpr int AF32_WARPS = 4;
constexpr int AF32_MAX_T = 1024;

struct AttnF32Params {
  const float* q;
  long long q_ld, q_hs, q_gs;
  const float* k[AF32_MAX_SRC];
  const float* v[AF32_MAX_SRC];
  long long kv_ld[AF32_MAX_SRC], kv_hs[AF32_MAX_SRC], kv_gs[AF32_MAX_SRC];
  float* o;
  long long o_ld, o_gs;
  float* o_src;
  long long o_src_stride;
  float* lse;
  int n_src, G, B, T, H, hs;
  float scale;
};

__global__ void __launch_bounds__(AF32_WARPS * 32) attn_fwd_f32_kernel(const __grid_constant__ AttnF32Params p) {
  pdl_prologue();
  __shared__ __align__(16) float s_q[AF32_WARPS][64];
  __shared__ float s_p[AF32_WARPS][AF32_MAX_T];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_block = AF32_WARPS;
  const int n_rb = (p.T + rows_per_block - 1) / rows_per_block;
  long long w = blockIdx.x;
  const int rb = (int)(w % n_rb); w /= n_rb;
  const int h = (int)(w % p.H); w /= p.H;
  const int b = (int)(w % p.B);
  const int g = (int)(w / p.B);
  const int r = rb * rows_per_block + warp;
  if (r >= p.T) return;
  const long long tok = (long long)b * p.T;
  const float* qr = p.q + g * p.q_gs + (tok + r) * p.q_ld + (long long)h * p.q_hs;
  for (int d = lane; d < p.hs; d += 32) s_q[warp][d] = qr[d];
  __syncwarp();
  float acc0 = 0.f, acc1 = 0.f;            // this lane's output dims lane, lane + 32, summed over the sources
  for (int s = 0; s < p.n_src; ++s) {
    const float* kb = p.k[s] + g * p.kv_gs[s] + tok * p.kv_ld[s] + (long long)h * p.kv_hs[s];
    const float* vb = p.v[s] + g * p.kv_gs[s] + tok * p.kv_ld[s] + (long long)h * p.kv_hs[s];
    float mx = -INFINITY;
    const bool vec = (p.hs & 3) == 0 && ((reinterpret_cast<uintptr_t>(kb) | (uintptr_t)(p.kv_ld[s] * 4)) & 15) == 0;
    for (int j = lane; j <= r; j += 32) {
      const float* kr = kb + (long long)j * p.kv_ld[s];
      float a = 0.f;
      if (vec) {              // 16-byte loads, four independent partial sums
        float a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int d = 0; d < p.hs; d += 4) {
          const float4 kk = *reinterpret_cast<const float4*>(kr + d);
          const float4 qq = *reinterpret_cast<const float4*>(&s_q[warp][d]);
          a = fmaf(qq.x, kk.x, a); a1 = fmaf(qq.y, kk.y, a1); a2 = fmaf(qq.z, kk.z, a2); a3 = fmaf(qq.w, kk.w, a3);
        }
        a = (a + a1) + (a2 + a3);
      } else {
        for (int d = 0; d < p.hs; ++d) a = fmaf(s_q[warp][d], kr[d], a);
      }
      a *= p.scale;
      s_p[warp][j] = a;
      mx = fmaxf(mx, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float l = 0.f;
    for (int j = lane; j <= r; j += 32) {
      const float e = expf(s_p[warp][j] - mx);
      s_p[warp][j] = e;
      l += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    const bool d0 = lane < p.hs, d1 = lane + 32 < p.hs;
    {
      // four key rows per iteration: eight independent loads in flight per lane
      float p0 = 0.f, p1 = 0.f, q0 = 0.f, q1 = 0.f, r0 = 0.f, r1 = 0.f, t0 = 0.f, t1 = 0.f;
      const long long ld = p.kv_ld[s];
      int j = 0;
      for (; j + 3 <= r; j += 4) {
        const float* vr = vb + (long long)j * ld;
        const float w0 = s_p[warp][j], w1 = s_p[warp][j + 1], w2 = s_p[warp][j + 2], w3 = s_p[warp][j + 3];
        if (d0) { p0 = fmaf(w0, vr[lane], p0); q0 = fmaf(w1, vr[ld + lane], q0); r0 = fmaf(w2, vr[2 * ld + lane], r0); t0 = fmaf(w3, vr[3 * ld + lane], t0); }
        if (d1) { p1 = fmaf(w0, vr[lane + 32], p1); q1 = fmaf(w1, vr[ld + lane + 32], q1); r1 = fmaf(w2, vr[2 * ld + lane + 32], r1); t1 = fmaf(w3, vr[3 * ld + lane + 32], t1); }
      }
      for (; j <= r; ++j) {
        const float pj = s_p[warp][j];
        const float* vr = vb + (long long)j * ld;
        if (d0) p0 = fmaf(pj, vr[lane], p0);
        if (d1) p1 = fmaf(pj, vr[lane + 32], p1);
      }
      o0 = (p0 + q0) + (r0 + t0);
      o1 = (p1 + q1) + (r1 + t1);
    }
    const float inv = 1.f / l;
    o0 *= inv; o1 *= inv;
    if (p.o_src) {
      float* os = p.o_src + (long long)s * p.o_src_stride + g * p.o_gs + (tok + r) * p.o_ld + (long long)h * p.hs;
      if (d0) os[lane] = o0;
      if (d1) os[lane + 32] = o1;
    }
    acc0 += o0; acc1 += o1;
    if (lane == 0) p.lse[((((long long)s * p.G + g) * p.B + b) * p.H + h) * p.T + r] = mx + logf(l);
    __syncwarp();
  }
  float* orow = p.o + g * p.o_gs + (tok + r) * p.o_ld + (long long)h * p.hs;
  if (lane < p.hs) orow[lane] = acc0;
  if (lane + 32 < p.hs) orow[lane + 32] = acc1;
}

}  // namespace ta

using namespace ta;

extern "C" int32_t ta_attention_fwd_f32(const ta_attn_tensor* q, const ta_attn_tensor* k, const ta_attn_tensor* v, int32_t n_src,
                                        float* o, int64_t ldo, int64_t o_gs, float* o_src, int64_t o_src_stride, float* lse,
                                        int32_t G, int32_t B, int32_t T, int32_t H, int32_t hs, void* stream) {
  TA_CHECK_ARG(q && k && v && o && lse, "ta_attention_fwd_f32: null pointer");
  TA_CHECK_ARG(n_src >= 1 && n_src <= AF32_MAX_SRC, "ta_attention_fwd_f32: n_src must be in [1, %d]", AF32_MAX_SRC);
  TA_CHECK_ARG(hs >= 1 && hs <= 64, "ta_attention_fwd_f32: head_size %d not supported (<= 64)", (int)hs);
  TA_CHECK_ARG(G > 0 && B > 0 && H > 0 && T > 0 && T <= AF32_MAX_T, "ta_attention_fwd_f32: T = %d not supported (<= %d)", (int)T, AF32_MAX_T);
  AttnF32Params p{};
  p.q = static_cast<const float*>(q->ptr); p.q_ld = q->ld; p.q_hs = q->head_stride; p.q_gs = q->gs;
  for (int s = 0; s < n_src; ++s) {
    TA_CHECK_ARG(k[s].ptr && v[s].ptr, "ta_attention_fwd_f32: null source pointer");
    TA_CHECK_ARG(k[s].ld == v[s].ld && k[s].head_stride == v[s].head_stride && k[s].gs == v[s].gs,
                 "ta_attention_fwd_f32: k and v of a source must share strides");
    p.k[s] = static_cast<const float*>(k[s].ptr); p.v[s] = static_cast<const float*>(v[s].ptr);
    p.kv_ld[s] = k[s].ld; p.kv_hs[s] = k[s].head_stride; p.kv_gs[s] = k[s].gs;
  }
  p.o = o; p.o_ld = ldo; p.o_gs = o_gs; p.o_src = o_src; p.o_src_stride = o_src_stride; p.lse = lse;
  p.n_src = n_src; p.G = G; p.B = B; p.T = T; p.H = H; p.hs = hs;
  p.scale = 1.f / sqrtf((float)hs);
  const long long blocks = (long long)G * B * H * ceil_div(T, AF32_WARPS);
  TA_CHECK_ARG(blocks < (1ll << 31), "ta_attention_fwd_f32: grid too large");
  launch_kernel(attn_fwd_f32_kernel, dim3((unsigned)blocks), dim3(AF32_WARPS * 32), 0, (cudaStream_t)stream, p);
  TA_CHECK_LAUNCH();
  return 0;
}
