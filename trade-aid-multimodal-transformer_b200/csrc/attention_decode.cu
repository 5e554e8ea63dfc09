// K3d — attention of ONE new query row per (group, batch, head) against a key/value cache: the decode step of
// `generate` (/root/reference/model.py:404-446) while the sequence still fits block_size. The reference re-runs the
// whole forward for every new token; in that growing phase every earlier position keeps its activations (causal
// masks, absolute positions), so only the new position is computed and its keys / values are appended to the cache.
// Same math as model.py:60-73 / 136-159 for the last row: per source an independent softmax(q k^T * hs^-0.5) v over
// positions [0, len), outputs of the sources summed.
//
// HBM-bound (every cached K and V row is read once per step): one warp per (group, batch, head); a lane owns one cached
// position per 32-position chunk - it scores its K row (2*hs contiguous bytes), the online softmax runs on warp-shuffle
// reductions, and it accumulates p * (its V row) privately; the lanes are summed once per source by recursive halving.
// All 16-byte loads of a chunk are independent (a first version that walked the V rows one after the other ran at
// 0.9 TB/s on c5 - one dependent load per step of the chain).
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "common.cuh"

namespace ta {

constexpr int AD_MAX_SRC = 16;
constexpr int AD_WARPS = 4;

struct AttnDecodeParams {
  const __nv_bfloat16* q;
  long long q_ld, q_hs, q_gs;
  const __nv_bfloat16* k[AD_MAX_SRC];
  const __nv_bfloat16* v[AD_MAX_SRC];
  long long kv_ld[AD_MAX_SRC], kv_hs[AD_MAX_SRC], kv_gs[AD_MAX_SRC], kv_bs[AD_MAX_SRC];
  __nv_bfloat16* o;
  long long o_ld, o_gs;
  int n_src, G, B, H, hs, len;
  float scale_log2;     // hs^-0.5 * log2(e)
};

// Sum over the 32 lanes of a per-lane array acc[HS] by recursive halving: at each step a lane keeps one half of its
// live elements and hands the other half to its partner, so the whole reduction costs HS - 1 (+ a few) shuffles instead
// of 5 * HS. Afterwards acc[0] (and acc[1] for HS = 64) hold complete sums: HS = 64 -> dims 2*lane, 2*lane + 1;
// HS = 32 -> dim lane; HS = 16 -> dim lane >> 1; HS = 8 -> dim lane >> 2 (duplicated over the lanes that share a dim).
template <int HS>
__device__ __forceinline__ void lane_reduce(float (&acc)[HS], int lane) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int mask = 16 >> s;
    const int nb = (HS >> s) > 1 ? (HS >> s) : 1;     // live elements before this step (compile time after unrolling)
    if (nb > (HS == 64 ? 2 : 1)) {
      const int half = nb / 2;
      const bool up = (lane & mask) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const float send = up ? acc[i] : acc[i + half];
        const float keep = up ? acc[i + half] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
      }
    } else {
      acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], mask);
      if (HS == 64) acc[1] += __shfl_xor_sync(0xffffffffu, acc[1], mask);
    }
  }
}

template <int HS>
__global__ void __launch_bounds__(AD_WARPS * 32) attn_decode_kernel(const __grid_constant__ AttnDecodeParams p) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * AD_WARPS + (threadIdx.x >> 5);
  const long long total = (long long)p.G * p.B * p.H;
  if (w >= total) return;
  const int h = (int)(w % p.H);
  const long long gb = w / p.H;
  const int b = (int)(gb % p.B), g = (int)(gb / p.B);

  // the query row of this head, pre-scaled, in every lane
  float q[HS];
  {
    const __nv_bfloat16* qp = p.q + g * p.q_gs + (long long)b * p.q_ld + (long long)h * p.q_hs;
#pragma unroll
    for (int i = 0; i < HS / 8; ++i) {
      const uint4 u = *reinterpret_cast<const uint4*>(qp + 8 * i);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        q[8 * i + 2 * j] = f.x * p.scale_log2;
        q[8 * i + 2 * j + 1] = f.y * p.scale_log2;
      }
    }
  }
  float out0 = 0.f, out1 = 0.f;
  for (int s = 0; s < p.n_src; ++s) {
    const long long base = g * p.kv_gs[s] + (long long)b * p.kv_bs[s] + (long long)h * p.kv_hs[s];
    const __nv_bfloat16* kb = p.k[s] + base;
    const __nv_bfloat16* vb = p.v[s] + base;
    const long long ld = p.kv_ld[s];
    // A lane owns one cached position per 32-position chunk in BOTH phases: it scores its K row and accumulates
    // p * (its V row) into a private acc[HS]; the lanes are summed once per source. Every load of a chunk (2 * HS / 8
    // 16-byte loads per lane) is independent and in flight together.
    float m = -INFINITY, l = 0.f;
    float acc[HS];
#pragma unroll
    for (int i = 0; i < HS; ++i) acc[i] = 0.f;
    for (int p0 = 0; p0 < p.len; p0 += 32) {
      const int pos = p0 + lane;
      const bool ok = pos < p.len;
      uint4 ku[HS / 8], vu[HS / 8];
      if (ok) {
        const uint4* kr = reinterpret_cast<const uint4*>(kb + (long long)pos * ld);
        const uint4* vr = reinterpret_cast<const uint4*>(vb + (long long)pos * ld);
#pragma unroll
        for (int i = 0; i < HS / 8; ++i) ku[i] = __ldg(kr + i);
#pragma unroll
        for (int i = 0; i < HS / 8; ++i) vu[i] = __ldg(vr + i);
      } else {
#pragma unroll
        for (int i = 0; i < HS / 8; ++i) ku[i] = vu[i] = make_uint4(0, 0, 0, 0);
      }
      float sc = 0.f;
#pragma unroll
      for (int i = 0; i < HS / 8; ++i) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&ku[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h2[j]);
          sc = fmaf(q[8 * i + 2 * j], f.x, sc);
          sc = fmaf(q[8 * i + 2 * j + 1], f.y, sc);
        }
      }
      if (!ok) sc = -INFINITY;
      float cm = sc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
      const float m_new = fmaxf(m, cm);                 // finite: position p0 is always valid
      const float corr = exp2f(m - m_new);              // exp2f(-inf) = 0 on the first chunk
      const float pr = exp2f(sc - m_new);               // 0 for masked lanes
      float ps = pr;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
      l = l * corr + ps;
      m = m_new;
#pragma unroll
      for (int i = 0; i < HS / 8; ++i) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&vu[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h2[j]);
          acc[8 * i + 2 * j] = fmaf(pr, f.x, acc[8 * i + 2 * j] * corr);
          acc[8 * i + 2 * j + 1] = fmaf(pr, f.y, acc[8 * i + 2 * j + 1] * corr);
        }
      }
    }
    lane_reduce<HS>(acc, lane);
    const float inv = 1.f / l;
    out0 += acc[0] * inv;
    if (HS == 64) out1 += acc[1] * inv;
  }
  __nv_bfloat16* op = p.o + g * p.o_gs + (long long)b * p.o_ld + (long long)h * HS;
  if (HS == 64) {
    *reinterpret_cast<__nv_bfloat162*>(op + 2 * lane) = __floats2bfloat162_rn(out0, out1);
  } else if (HS == 32) {
    op[lane] = __float2bfloat16_rn(out0);
  } else if (HS == 16) {
    if ((lane & 1) == 0) op[lane >> 1] = __float2bfloat16_rn(out0);
  } else {
    if ((lane & 3) == 0) op[lane >> 2] = __float2bfloat16_rn(out0);
  }
}

}  // namespace ta

extern "C" int32_t ta_attention_decode(const ta_attn_tensor* q, const ta_attn_tensor* k, const ta_attn_tensor* v,
                                       const int64_t* kv_batch_stride, int32_t n_src, void* o_bf16, int64_t ldo, int64_t o_gs,
                                       int32_t G, int32_t B, int32_t len, int32_t H, int32_t hs, void* stream) {
  using namespace ta;
  TA_CHECK_ARG(q && k && v && kv_batch_stride && o_bf16, "ta_attention_decode: null argument");
  TA_CHECK_ARG(n_src >= 1 && n_src <= AD_MAX_SRC, "ta_attention_decode: n_src %d not in [1, %d]", (int)n_src, AD_MAX_SRC);
  TA_CHECK_ARG(G > 0 && B > 0 && H > 0 && len > 0, "ta_attention_decode: non-positive shape");
  TA_CHECK_ARG(hs == 8 || hs == 16 || hs == 32 || hs == 64, "ta_attention_decode: head size %d not in {8, 16, 32, 64}", (int)hs);
  auto al16 = [](const void* ptr, int64_t a, int64_t b2, int64_t c) {
    return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && a % 8 == 0 && b2 % 8 == 0 && c % 8 == 0;
  };
  TA_CHECK_ARG(al16(q->ptr, q->ld, q->head_stride, q->gs), "ta_attention_decode: q must be 16-byte aligned (ptr, strides)");
  AttnDecodeParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q->ptr); p.q_ld = q->ld; p.q_hs = q->head_stride; p.q_gs = q->gs;
  for (int s = 0; s < n_src; ++s) {
    TA_CHECK_ARG(k[s].ptr && v[s].ptr, "ta_attention_decode: null cache pointer");
    TA_CHECK_ARG(al16(k[s].ptr, k[s].ld, k[s].head_stride, k[s].gs) && kv_batch_stride[s] % 8 == 0,
                 "ta_attention_decode: k cache of source %d must be 16-byte aligned (ptr, strides)", s);
    TA_CHECK_ARG(k[s].ld == v[s].ld && k[s].head_stride == v[s].head_stride && k[s].gs == v[s].gs,
                 "ta_attention_decode: k and v of a source must share strides");
    TA_CHECK_ARG((reinterpret_cast<uintptr_t>(v[s].ptr) & 15) == 0, "ta_attention_decode: v cache must be 16-byte aligned");
    p.k[s] = static_cast<const __nv_bfloat16*>(k[s].ptr);
    p.v[s] = static_cast<const __nv_bfloat16*>(v[s].ptr);
    p.kv_ld[s] = k[s].ld; p.kv_hs[s] = k[s].head_stride; p.kv_gs[s] = k[s].gs; p.kv_bs[s] = kv_batch_stride[s];
  }
  p.o = static_cast<__nv_bfloat16*>(o_bf16); p.o_ld = ldo; p.o_gs = o_gs;
  TA_CHECK_ARG((reinterpret_cast<uintptr_t>(o_bf16) & 3) == 0 && ldo % 2 == 0 && o_gs % 2 == 0, "ta_attention_decode: o misaligned");
  p.n_src = n_src; p.G = G; p.B = B; p.H = H; p.hs = hs; p.len = len;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)hs);
  const long long total = (long long)G * B * H;
  const unsigned grid = (unsigned)ceil_div(total, AD_WARPS);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (hs) {
    case 64: attn_decode_kernel<64><<<grid, AD_WARPS * 32, 0, st>>>(p); break;
    case 32: attn_decode_kernel<32><<<grid, AD_WARPS * 32, 0, st>>>(p); break;
    case 16: attn_decode_kernel<16><<<grid, AD_WARPS * 32, 0, st>>>(p); break;
    default: attn_decode_kernel<8><<<grid, AD_WARPS * 32, 0, st>>>(p); break;
  }
  TA_CHECK_LAUNCH();
  return 0;
}
