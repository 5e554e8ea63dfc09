// K3d — attention of ONE new query row per (group, batch, head) against a key/value cache: the decode step of
// `generate` (/root/reference/model.py:404-446) while the sequence still fits block_size. The reference re-runs the
// whole forward for every new token; in that growing phase every earlier position keeps its activations (causal
// masks, absolute positions), so only the new position is computed and its keys / values are appended to the cache.
// Same math as model.py:60-73 / 136-159 for the last row: per source an independent softmax(q k^T * hs^-0.5) v over
// positions [0, len), outputs of the sources summed.
//
// HBM-bound (every cached K and V row is read once per step): one warp per (group, batch, head); hs/8 neighbouring
// lanes share a cached row (16 bytes each), partial dot products are completed with warp shuffles, the online softmax
// runs on shuffle reductions, p * V is accumulated per lane for its 8 head dims and summed over the row groups once per
// source. All 16-byte loads of a chunk are independent and the next chunk's loads are issued as soon as registers free up.
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
  const int* len_dev;   // optional: positions already cached, on the device (CUDA-graph replay); len = *len_dev + len then
  float scale_log2;     // hs^-0.5 * log2(e)
};

// Work split inside a warp: LPR = HS / 8 neighbouring lanes share one cached row (each owns 16 bytes = 8 head dims of
// it), so one warp-wide 16-byte load touches 32 / LPR whole rows (HS = 64: four full 128-byte lines) instead of 32
// different lines - the first version of this kernel, one row per lane, was bound by L1 tag lookups (0.6 of the HBM
// rate). A lane takes AD_U rows per chunk: rows p0 + u * RPI + lane / LPR.
constexpr int AD_U = 8;

template <int HS>
__global__ void __launch_bounds__(AD_WARPS * 32) attn_decode_kernel(const __grid_constant__ AttnDecodeParams p) {
  pdl_prologue();
  const int len = p.len + (p.len_dev ? *p.len_dev : 0);      // cached positions to attend to
  constexpr int LPR = HS / 8;            // lanes per row
  constexpr int RPI = 32 / LPR;          // rows per warp-wide load instruction
  constexpr int CHUNK = RPI * AD_U;      // rows per chunk
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * AD_WARPS + (threadIdx.x >> 5);
  const long long total = (long long)p.G * p.B * p.H;
  if (w >= total) return;
  const int h = (int)(w % p.H);
  const long long gb = w / p.H;
  const int b = (int)(gb % p.B), g = (int)(gb / p.B);
  const int piece = lane % LPR, rg = lane / LPR;
  // HS is the head size rounded up to a power of two (lane layout); lanes whose 8 dims lie beyond the real head size
  // (hs = 24, 40, 48, 56) load nothing and contribute zeros
  const bool act = 8 * piece < p.hs;

  // this lane's 8 dims of the query row, pre-scaled
  float q[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) q[d] = 0.f;
  if (act) {
    const __nv_bfloat16* qp = p.q + g * p.q_gs + (long long)b * p.q_ld + (long long)h * p.q_hs + 8 * piece;
    const uint4 u = *reinterpret_cast<const uint4*>(qp);
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h2[j]);
      q[2 * j] = f.x * p.scale_log2;
      q[2 * j + 1] = f.y * p.scale_log2;
    }
  }
  float out[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) out[d] = 0.f;
  for (int s = 0; s < p.n_src; ++s) {
    const long long base = g * p.kv_gs[s] + (long long)b * p.kv_bs[s] + (long long)h * p.kv_hs[s] + 8 * piece;
    const __nv_bfloat16* kb = p.k[s] + base;
    const __nv_bfloat16* vb = p.v[s] + base;
    const long long ld = p.kv_ld[s];
    float m = -INFINITY, l = 0.f;
    float acc[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] = 0.f;
    uint4 ku[AD_U], vu[AD_U];
    auto load_rows = [&](uint4 (&dst)[AD_U], const __nv_bfloat16* base_ptr, int p0) {
#pragma unroll
      for (int u = 0; u < AD_U; ++u) {
        const int pos = p0 + u * RPI + rg;
        dst[u] = (pos < len && act) ? __ldg(reinterpret_cast<const uint4*>(base_ptr + (long long)pos * ld)) : make_uint4(0, 0, 0, 0);
      }
    };
    // software pipeline without extra registers: the K registers are dead once the scores are formed, the V registers
    // once the accumulators are updated - the next chunk's rows are requested right there
    load_rows(ku, kb, 0);
    load_rows(vu, vb, 0);
    for (int p0 = 0; p0 < len; p0 += CHUNK) {
      float sc[AD_U];
#pragma unroll
      for (int u = 0; u < AD_U; ++u) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&ku[u]);
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h2[j]);
          a = fmaf(q[2 * j], f.x, a);
          a = fmaf(q[2 * j + 1], f.y, a);
        }
        sc[u] = a;
      }
      if (p0 + CHUNK < len) load_rows(ku, kb, p0 + CHUNK);
      // complete the dot products over the LPR lanes of a row; afterwards every lane of the row holds the score
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < AD_U; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], o);
      }
      float cm = -INFINITY;
#pragma unroll
      for (int u = 0; u < AD_U; ++u) {
        if (p0 + u * RPI + rg >= len) sc[u] = -INFINITY;
        cm = fmaxf(cm, sc[u]);
      }
#pragma unroll
      for (int o = 16; o >= LPR; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
      const float m_new = fmaxf(m, cm);                 // finite: position p0 is always valid
      const float corr = exp2f(m - m_new);              // exp2f(-inf) = 0 on the first chunk
      float ps = 0.f;
#pragma unroll
      for (int u = 0; u < AD_U; ++u) {
        sc[u] = exp2f(sc[u] - m_new);                   // 0 for masked rows
        ps += sc[u];
      }
#pragma unroll
      for (int o = 16; o >= LPR; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
      l = l * corr + ps;
      m = m_new;
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] *= corr;
#pragma unroll
      for (int u = 0; u < AD_U; ++u) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&vu[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h2[j]);
          acc[2 * j] = fmaf(sc[u], f.x, acc[2 * j]);
          acc[2 * j + 1] = fmaf(sc[u], f.y, acc[2 * j + 1]);
        }
      }
      if (p0 + CHUNK < len) load_rows(vu, vb, p0 + CHUNK);
    }
    // sum the row groups (lanes with the same piece); every lane ends with the full sum of its 8 dims
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) {
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < 8; ++d) out[d] += acc[d] * inv;
  }
  if (rg == 0 && act) {
    uint4 u;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(out[2 * j], out[2 * j + 1]);
    *reinterpret_cast<uint4*>(p.o + g * p.o_gs + (long long)b * p.o_ld + (long long)h * p.hs + 8 * piece) = u;
  }
}

}  // namespace ta

extern "C" int32_t ta_attention_decode(const ta_attn_tensor* q, const ta_attn_tensor* k, const ta_attn_tensor* v,
                                       const int64_t* kv_batch_stride, int32_t n_src, void* o_bf16, int64_t ldo, int64_t o_gs,
                                       int32_t G, int32_t B, int32_t len, const int32_t* len_dev, int32_t H, int32_t hs,
                                       void* stream) {
  using namespace ta;
  TA_CHECK_ARG(q && k && v && kv_batch_stride && o_bf16, "ta_attention_decode: null argument");
  TA_CHECK_ARG(n_src >= 1 && n_src <= AD_MAX_SRC, "ta_attention_decode: n_src %d not in [1, %d]", (int)n_src, AD_MAX_SRC);
  TA_CHECK_ARG(G > 0 && B > 0 && H > 0 && (len > 0 || (len_dev && len >= 0)), "ta_attention_decode: non-positive shape");
  TA_CHECK_ARG(hs >= 8 && hs <= 64 && hs % 8 == 0, "ta_attention_decode: head size %d not supported (8..64, multiple of 8)", (int)hs);
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
  TA_CHECK_ARG((reinterpret_cast<uintptr_t>(o_bf16) & 15) == 0 && ldo % 8 == 0 && o_gs % 8 == 0, "ta_attention_decode: o must be 16-byte aligned (ptr, strides)");
  p.n_src = n_src; p.G = G; p.B = B; p.H = H; p.hs = hs; p.len = len; p.len_dev = len_dev;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)hs);
  const long long total = (long long)G * B * H;
  const unsigned grid = (unsigned)ceil_div(total, AD_WARPS);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // lane layout of the next power of two: 24 -> 32, 40 / 48 / 56 -> 64
  if (hs > 32) launch_kernel(attn_decode_kernel<64>, dim3(grid), dim3(AD_WARPS * 32), 0, st, p);
  else if (hs > 16) launch_kernel(attn_decode_kernel<32>, dim3(grid), dim3(AD_WARPS * 32), 0, st, p);
  else if (hs > 8) launch_kernel(attn_decode_kernel<16>, dim3(grid), dim3(AD_WARPS * 32), 0, st, p);
  else launch_kernel(attn_decode_kernel<8>, dim3(grid), dim3(AD_WARPS * 32), 0, st, p);
  TA_CHECK_LAUNCH();
  return 0;
}
