// K3 forward — causal flash-style attention on tcgen05/TMEM for both the self-attention heads
// (/root/reference/model.py:60-73) and the selective cross-attention heads (model.py:136-159, one independent causal
// softmax per key/value source modality, outputs summed).
//
// One CTA (128 threads) per (group g, batch b, head h, 128-row query tile). Per key/value block of KVB rows:
//   TMA  : Q tile once; K and V tiles double-buffered (prefetch of the next block/source overlaps the current one)
//   UMMA : S = Q K^T  (128 x KVB x hs, fp32 in TMEM)             -- issued by thread 0
//   SIMT : thread r owns row r of S: tcgen05.ld, scale, causal mask, online softmax (running max / sum, exp2),
//          P (bf16) written to 128B-swizzled smem as the A operand of the second UMMA
//   UMMA : O_blk = P V   (128 x hs x KVB; V is the MN-major B operand straight from its [kv, d] TMA tile)
//   SIMT : O = O * alpha + O_blk in registers; at the end of a source: O_total += O / l, lse stored
// Nothing of size T x T ever touches HBM. head_size <= 64 (padded with TMA zero-fill to a 128-byte row).
#include <cmath>
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace ta {

constexpr int ATT_MAX_SRC = 8;

struct AttnMaps {
  CUtensorMap q;
  CUtensorMap k[ATT_MAX_SRC];
  CUtensorMap v[ATT_MAX_SRC];
};

struct AttnFwdParams {
  int T, B, H, G, hs, n_src, n_qtiles;
  float scale_log2;
  __nv_bfloat16* o;       // [G][B*T][ldo], head h at column h*hs
  long long ldo, o_gs;
  __nv_bfloat16* o_src;   // optional per-source outputs [n_src][G][B*T][ldo] (needed by backward when n_src > 1)
  long long o_src_stride;
  float* lse;             // [n_src][G][B][H][T], natural log
};

template <int KVB, int DP>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const __grid_constant__ AttnMaps maps, const AttnFwdParams p) {
  constexpr int KV_TILE_BYTES = KVB * 128;
  constexpr int TMEM_COLS = (KVB + DP) <= 128 ? 128 : 256;
  // all shared memory is dynamic (no static part), so the tile base is the 1024-byte aligned start of the CTA window
  // and two CTAs fit one SM: [Q 16 KB][2 x (K, V)][P KVB/64 x 16 KB][barriers]
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + 16384;
  uint8_t* sP = sKV + 4 * KV_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (KVB / 64) * 16384);
  uint64_t& bar_q = bars[0];
  uint64_t* bar_kv = bars + 1;
  uint64_t& bar_s = bars[3];
  uint64_t& bar_o = bars[4];
  uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(bars + 5);
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const int tid = threadIdx.x, warp = tid >> 5;
  int bid = blockIdx.x;
  const int qt = bid % p.n_qtiles; bid /= p.n_qtiles;
  const int h = bid % p.H; bid /= p.H;
  const int b = bid % p.B;
  const int g = bid / p.B;
  const int q0 = qt * 128;
  const int last_row = (q0 + 127 < p.T - 1) ? q0 + 127 : p.T - 1;
  const int n_kvb = last_row / KVB + 1;
  const int total_steps = p.n_src * n_kvb;

  if (tid == 0) {
    mbar_init(&bar_q, 1);
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + KVB;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;

  auto issue_kv = [&](int step) {
    const int s = step / n_kvb, j = step - s * n_kvb, st = step & 1;
    uint8_t* dK = sKV + st * 2 * KV_TILE_BYTES;
    mbar_expect_tx(&bar_kv[st], 2 * KV_TILE_BYTES);
    tma_load_5d(dK, &maps.k[s], &bar_kv[st], 0, h, j * KVB, b, g);
    tma_load_5d(dK + KV_TILE_BYTES, &maps.v[s], &bar_kv[st], 0, h, j * KVB, b, g);
  };

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, KVB, 0, 0);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, DP, 0, 1);
  const int nk_qk = (p.hs + 15) / 16;
  auto issue_s = [&](int step) {   // S = Q K^T of a step whose K tile has landed
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sKV + (step & 1) * 2 * KV_TILE_BYTES);
    for (int kk = 0; kk < nk_qk; ++kk)
      umma_bf16(tmem_S, umma_desc_sw128(aQ + kk * 32, 16, 1024), umma_desc_sw128(aK + kk * 32, 16, 1024), idesc_s, kk != 0);
    umma_commit(&bar_s);
  };
  if (tid == 0) {
    mbar_expect_tx(&bar_q, 16384);
    tma_load_5d(sQ, &maps.q, &bar_q, 0, h, q0, b, g);
    issue_kv(0);
    if (total_steps > 1) issue_kv(1);
    mbar_wait(&bar_q, 0);
    mbar_wait(&bar_kv[0], 0);
    tc_fence_after();
    issue_s(0);
  }

  const int row = q0 + tid;  // query position of this thread
  float o_tot[DP], o_acc[DP];
#pragma unroll
  for (int i = 0; i < DP; ++i) o_tot[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  for (int step = 0; step < total_steps; ++step) {
    const int s = step / n_kvb, j = step - s * n_kvb, st = step & 1;
    const int kv0 = j * KVB;
    const uint32_t par = step & 1;
    if (j == 0) {
      m_run = -INFINITY;
      l_run = 0.f;
#pragma unroll
      for (int i = 0; i < DP; ++i) o_acc[i] = 0.f;
    }
    mbar_wait(&bar_s, par);
    tc_fence_after();
    // number of 32-column chunks that hold at least one unmasked column for some row of this warp
    const int warp_last_row = q0 + warp * 32 + 31;
    int n_chunks = (warp_last_row - kv0) / 32 + 1;
    if (n_chunks > KVB / 32) n_chunks = KVB / 32;
    if (n_chunks < 0) n_chunks = 0;
    // pass 1: row maximum
    float mx = m_run;
    for (int c = 0; c < n_chunks; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_S + lane_off + c * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = (kv0 + c * 32 + i <= row) ? __uint_as_float(v[i]) * p.scale_log2 : -INFINITY;
        mx = fmaxf(mx, x);
      }
    }
    const float alpha = exp2f(m_run - mx);  // m_run = -inf on the first block => 0
    float psum = 0.f;
    // pass 2: probabilities -> bf16 -> swizzled smem (A operand of P V)
    for (int c = 0; c < KVB / 32; ++c) {
      uint32_t packed[16];
      if (c < n_chunks) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_S + lane_off + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float x0 = (kv0 + c * 32 + i <= row) ? exp2f(__uint_as_float(v[i]) * p.scale_log2 - mx) : 0.f;
          const float x1 = (kv0 + c * 32 + i + 1 <= row) ? exp2f(__uint_as_float(v[i + 1]) * p.scale_log2 - mx) : 0.f;
          psum += x0 + x1;
          __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
          packed[i / 2] = *reinterpret_cast<uint32_t*>(&h2);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) packed[i] = 0u;
      }
      // 32 columns = 4 chunks of 16 B at logical chunk index (c%2)*4 + q inside half c/2
      uint8_t* rowp = sP + (c >> 1) * 16384 + tid * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int lc = (c & 1) * 4 + q;
        *reinterpret_cast<uint4*>(rowp + ((lc ^ (tid & 7)) << 4)) =
            make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
      }
    }
    l_run = l_run * alpha + psum;
    m_run = mx;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aP = smem_u32(sP), aV = smem_u32(sKV + st * 2 * KV_TILE_BYTES + KV_TILE_BYTES);
      // only the k-steps that can hold unmasked columns for some row of the tile
      int nk = (last_row - kv0) / 16 + 1;
      if (nk > KVB / 16) nk = KVB / 16;
      for (int kk = 0; kk < nk; ++kk)
        umma_bf16(tmem_O, umma_desc_sw128(aP + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                  umma_desc_sw128(aV + kk * 2048, 8192, 1024), idesc_o, kk != 0);
      umma_commit(&bar_o);
      // run ahead: S of the next step goes to the tensor pipe right behind P V (every thread has finished reading S)
      if (step + 1 < total_steps) {
        mbar_wait(&bar_kv[(step + 1) & 1], ((step + 1) >> 1) & 1);
        tc_fence_after();
        issue_s(step + 1);
      }
    }
    mbar_wait(&bar_o, par);
    tc_fence_after();
    if (tid == 0 && step + 2 < total_steps) issue_kv(step + 2);   // this step's K/V stage is free again
#pragma unroll
    for (int c = 0; c < DP / 16; ++c) {
      uint32_t v[16];
      tmem_ld_32x16(tmem_O + lane_off + c * 16, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) o_acc[c * 16 + i] = o_acc[c * 16 + i] * alpha + __uint_as_float(v[i]);
    }
    if (j == n_kvb - 1) {
      const float inv = 1.f / l_run;
      if (row < p.T) {
        p.lse[((((long long)s * p.G + g) * p.B + b) * p.H + h) * p.T + row] = m_run * 0.6931471805599453f + logf(l_run);
        if (p.o_src) {
          __nv_bfloat16* dst = p.o_src + (long long)s * p.o_src_stride + (long long)g * p.o_gs +
                               ((long long)b * p.T + row) * p.ldo + h * p.hs;
#pragma unroll
          for (int i = 0; i < DP; i += 8) {
            if (i < p.hs) {
              uint4 u;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(o_acc[i + 2 * e] * inv, o_acc[i + 2 * e + 1] * inv);
              *reinterpret_cast<uint4*>(dst + i) = u;
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < DP; ++i) o_tot[i] += o_acc[i] * inv;
    }
  }

  if (row < p.T) {
    __nv_bfloat16* dst = p.o + (long long)g * p.o_gs + ((long long)b * p.T + row) * p.ldo + h * p.hs;
#pragma unroll
    for (int i = 0; i < DP; i += 8) {
      if (i < p.hs) {
        uint4 u;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(o_tot[i + 2 * e], o_tot[i + 2 * e + 1]);
        *reinterpret_cast<uint4*>(dst + i) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn attn_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 5-D view {d < hs, head, t < T, batch, group} of a [G][B*T][ld] bf16 activation; box = {64, 1, box_rows, 1, 1}
int make_head_map(CUtensorMap* out, const ta_attn_tensor* t, int hs, int H, int T, int B, int G, int box_rows) {
  EncodeTiledFn enc = attn_encode_fn();
  TA_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  TA_CHECK_ARG(t->ptr && (reinterpret_cast<uintptr_t>(t->ptr) & 15) == 0, "attention tensor pointer must be 16-byte aligned");
  TA_CHECK_ARG(t->ld % 8 == 0 && t->head_stride % 8 == 0 && (G == 1 || t->gs % 8 == 0),
               "attention tensor strides must be multiples of 8 elements (ld %lld, head %lld, group %lld)",
               (long long)t->ld, (long long)t->head_stride, (long long)t->gs);
  cuuint64_t dims[5] = {(cuuint64_t)hs, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B, (cuuint64_t)G};
  cuuint64_t strides[4] = {(cuuint64_t)t->head_stride * 2, (cuuint64_t)t->ld * 2, (cuuint64_t)t->ld * 2 * T,
                           G > 1 ? (cuuint64_t)t->gs * 2 : 16};
  if (H == 1) strides[0] = 16;
  cuuint32_t box[5] = {64, 1, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(t->ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TA_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (attention) failed with CUresult %d", (int)r);
  return 0;
}

template <int KVB, int DP>
static int launch_attn_fwd(const AttnMaps& maps, const AttnFwdParams& p, cudaStream_t st) {
  constexpr int SMEM = 16384 + 4 * KVB * 128 + (KVB / 64) * 16384 + 64;
  static bool attr = false;
  if (!attr) {
    TA_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<KVB, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr = true;
  }
  const long long grid = (long long)p.n_qtiles * p.H * p.B * p.G;
  TA_CHECK_ARG(grid < (1ll << 31), "attention grid too large");
  attn_fwd_kernel<KVB, DP><<<(unsigned)grid, 128, SMEM, st>>>(maps, p);
  TA_CHECK_LAUNCH();
  return 0;
}

}  // namespace ta

using namespace ta;

extern "C" int32_t ta_attention_fwd(const ta_attn_tensor* q, const ta_attn_tensor* k, const ta_attn_tensor* v,
                                    int32_t n_src, void* o_bf16, int64_t ldo, int64_t o_gs, void* o_src_bf16,
                                    int64_t o_src_stride, float* lse, int32_t G, int32_t B, int32_t T, int32_t H,
                                    int32_t hs, float dropout_p, uint64_t seed, void* stream) {
  (void)seed;
  TA_CHECK_ARG(q && k && v && o_bf16 && lse, "ta_attention_fwd: null pointer");
  TA_CHECK_ARG(n_src >= 1 && n_src <= ATT_MAX_SRC, "ta_attention_fwd: n_src must be in [1, %d]", ATT_MAX_SRC);
  TA_CHECK_ARG(hs >= 8 && hs <= 64 && hs % 8 == 0, "ta_attention_fwd: head_size %d not supported (8..64, multiple of 8)", hs);
  TA_CHECK_ARG(G > 0 && B > 0 && T > 0 && H > 0, "ta_attention_fwd: non-positive size");
  TA_CHECK_ARG(ldo % 8 == 0 && o_gs % 8 == 0 && (reinterpret_cast<uintptr_t>(o_bf16) & 15) == 0, "ta_attention_fwd: output alignment");
  TA_CHECK_ARG(dropout_p == 0.f, "ta_attention_fwd: attention dropout not implemented in this build");
  const int KVB = T <= 64 ? 64 : 128;
  AttnMaps maps;
  if (make_head_map(&maps.q, q, hs, H, T, B, G, 128)) return 1;
  for (int s = 0; s < n_src; ++s) {
    if (make_head_map(&maps.k[s], &k[s], hs, H, T, B, G, KVB)) return 1;
    if (make_head_map(&maps.v[s], &v[s], hs, H, T, B, G, KVB)) return 1;
  }
  AttnFwdParams p{};
  p.T = T; p.B = B; p.H = H; p.G = G; p.hs = hs; p.n_src = n_src;
  p.n_qtiles = (T + 127) / 128;
  p.scale_log2 = (float)(1.4426950408889634 / sqrt((double)hs));
  p.o = (__nv_bfloat16*)o_bf16; p.ldo = ldo; p.o_gs = o_gs;
  p.o_src = (__nv_bfloat16*)o_src_bf16; p.o_src_stride = o_src_stride;
  p.lse = lse;
  auto st = (cudaStream_t)stream;
  const int DP = hs <= 16 ? 16 : (hs <= 32 ? 32 : 64);
  if (KVB == 64) {
    if (DP == 16) return launch_attn_fwd<64, 16>(maps, p, st);
    if (DP == 32) return launch_attn_fwd<64, 32>(maps, p, st);
    return launch_attn_fwd<64, 64>(maps, p, st);
  }
  if (DP == 16) return launch_attn_fwd<128, 16>(maps, p, st);
  if (DP == 32) return launch_attn_fwd<128, 32>(maps, p, st);
  return launch_attn_fwd<128, 64>(maps, p, st);
}
