// K3 backward — recompute attention backward on tcgen05/TMEM (autograd of /root/reference/model.py:60-73, 136-159).
//
// One CTA (128 threads) per (group, batch, head); the whole causal T x T triangle (T <= 256) is walked in 128 x 128
// tiles with every accumulator resident in TMEM (512 columns):
//   S, dP          2 x 128 columns   recomputed per tile pair:  S = Q_i K_j^T,  dP = dO_i V_j^T
//   dV_j, dK_j     2 x 64  columns   accumulated over the query tiles i >= j
//   dQ_0, dQ_1     2 x 64  columns   accumulated over key/value tiles j <= i AND over all key/value sources
// Thread r owns row r of the tile: P = exp2(S*c - lse), dS = P * (dP - D) * hs^-0.5 with D = rowsum(dO * O_src);
// P and dS go to 128B-swizzled smem once and are consumed by three UMMAs through two descriptor views of the same
// bytes: K-major (dQ += dS K) and MN-major (dV += P^T dO, dK += dS^T Q). K_j, V_j, Q_i, dO_i are read straight from
// their TMA tiles as MN-major B operands. No T x T tensor and no atomics; dq/dk/dv are written once, in bf16.
#include <cmath>
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace ta {

constexpr int ATTB_MAX_SRC = 8;
int make_head_map(CUtensorMap* out, const ta_attn_tensor* t, int hs, int H, int T, int B, int G, int box_rows);

struct AttnBwdMaps {
  CUtensorMap q, d_o, o;
  CUtensorMap k[ATTB_MAX_SRC];
  CUtensorMap v[ATTB_MAX_SRC];
};

struct GradOut {
  __nv_bfloat16* ptr;
  long long ld, head_stride, gs;
};

struct AttnBwdParams {
  int T, B, H, G, hs, n_src, n_qt;
  float scale, scale_log2;
  const __nv_bfloat16* o;      // [G][B*T][ldo]
  const __nv_bfloat16* o_src;  // [n_src] x same, or NULL when n_src == 1
  long long ldo, o_gs, o_src_stride;
  const float* lse;            // [n_src][G][B][H][T]
  GradOut dq;
  GradOut dk[ATTB_MAX_SRC];
  GradOut dv[ATTB_MAX_SRC];
};

template <int DP>
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const float (&v)[DP], int n) {
#pragma unroll
  for (int i = 0; i < DP; i += 8) {
    if (i < n) {
      uint4 u;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[i + 2 * e], v[i + 2 * e + 1]);
      *reinterpret_cast<uint4*>(dst + i) = u;
    }
  }
}

constexpr int ATTB_THREADS = 288;   // warps 0-7: softmax/gradient math (two column halves), warp 8: TMA + UMMA issuer

template <int DP>
__global__ void __launch_bounds__(ATTB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ AttnBwdMaps maps, const __grid_constant__ AttnBwdParams p) {
  constexpr int TILE = 16384;  // 128 rows x 128 B
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                // 2 tiles
  uint8_t* sdO = smem + 2 * TILE;    // 2 tiles
  uint8_t* sKV = smem + 4 * TILE;    // 2 stages x (K, V)
  uint8_t* sP = smem + 8 * TILE;     // 2 halves (kv 0-63, 64-127); also the landing zone of the O tiles for D
  uint8_t* sdS = smem + 10 * TILE;   // 2 halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 12 * TILE);
  uint64_t& bar_q = bars[0];         // Q and dO tiles landed
  uint64_t* bar_kv = bars + 1;       // [2] K/V stage landed
  uint64_t& bar_sp = bars[3];        // S and dP of a pair are in TMEM
  uint64_t& bar_g = bars[4];         // gradient UMMAs of a pair are complete (P/dS smem and dV/dK/dQ TMEM readable)
  uint64_t& bar_pr = bars[5];        // 256 arrivals: P/dS of a pair written to smem, S/dP TMEM consumed
  uint64_t& bar_o = bars[6];         // O tiles of a source landed (for D)
  uint64_t& bar_od = bars[7];        // 256 arrivals: D computed, O tiles consumed
  uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(bars + 8);
  float* sD = reinterpret_cast<float*>(bars + 10);   // [2][128] rowsum(dO * O_src) per query tile
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const int tid = threadIdx.x, warp = tid >> 5;
  int bid = blockIdx.x;
  const int h = bid % p.H; bid /= p.H;
  const int b = bid % p.B;
  const int g = bid / p.B;
  const int n_qt = p.n_qt;  // == number of kv tiles
  const int total_ks = p.n_src * n_qt;

  if (tid == 0) {
    mbar_init(&bar_q, 1);
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(&bar_sp, 1);
    mbar_init(&bar_g, 1);
    mbar_init(&bar_pr, 256);
    mbar_init(&bar_o, 1);
    mbar_init(&bar_od, 256);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 128, tm_dV = tmem_base + 256, tm_dK = tmem_base + 320;
  const uint32_t tm_dQ = tmem_base + 384;  // + 64 * i
  const int nk_qk = (p.hs + 15) / 16;

  if (warp == 8) {
    // ================================================================== TMA + UMMA issuer (one lane)
    if ((tid & 31) == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // S, dP
      constexpr uint32_t idesc_kv = umma_idesc_bf16(128, DP, 1, 1);   // dV, dK: A = P/dS seen MN-major, B = dO/Q MN-major
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, DP, 0, 1);    // dQ: A = dS K-major, B = K MN-major
      auto issue_kv = [&](int ks) {
        const int s = ks / n_qt, j = ks - s * n_qt, st = ks & 1;
        uint8_t* dK = sKV + st * 2 * TILE;
        mbar_expect_tx(&bar_kv[st], 2 * TILE);
        tma_load_5d(dK, &maps.k[s], &bar_kv[st], 0, h, j * 128, b, g);
        tma_load_5d(dK + TILE, &maps.v[s], &bar_kv[st], 0, h, j * 128, b, g);
      };
      auto issue_o = [&](int s) {   // O_src tiles of source s -> sP region (free: no gradient UMMA in flight)
        mbar_expect_tx(&bar_o, n_qt * TILE);
        for (int i = 0; i < n_qt; ++i) tma_load_5d(sP + i * TILE, &maps.o, &bar_o, 0, h, i * 128, b, p.o_src ? s : g);
      };
      auto issue_sdp = [&](int j, int i, int st) {
        const uint32_t aQ = smem_u32(sQ + i * TILE), aO = smem_u32(sdO + i * TILE);
        const uint32_t aK = smem_u32(sKV + st * 2 * TILE), aV = aK + TILE;
        for (int kk = 0; kk < nk_qk; ++kk)
          umma_bf16(tm_S, umma_desc_sw128(aQ + kk * 32, 16, 1024), umma_desc_sw128(aK + kk * 32, 16, 1024), idesc_s, kk != 0);
        for (int kk = 0; kk < nk_qk; ++kk)
          umma_bf16(tm_dP, umma_desc_sw128(aO + kk * 32, 16, 1024), umma_desc_sw128(aV + kk * 32, 16, 1024), idesc_s, kk != 0);
        umma_commit(&bar_sp);
      };
      mbar_expect_tx(&bar_q, n_qt * 2 * TILE);
      for (int i = 0; i < n_qt; ++i) {
        tma_load_5d(sQ + i * TILE, &maps.q, &bar_q, 0, h, i * 128, b, g);
        tma_load_5d(sdO + i * TILE, &maps.d_o, &bar_q, 0, h, i * 128, b, g);
      }
      issue_o(0);
      issue_kv(0);
      if (total_ks > 1) issue_kv(1);
      mbar_wait(&bar_q, 0);
      int pc = 0, g_seen = 0;
      for (int s = 0; s < p.n_src; ++s) {
        if (s > 0) issue_o(s);   // the last pair of the previous source was waited on (bar_g) below
        mbar_wait(&bar_od, s & 1);           // every math thread is done with the O tiles: sP may be overwritten
        for (int j = 0; j < n_qt; ++j) {
          const int ks = s * n_qt + j, st = ks & 1;
          mbar_wait(&bar_kv[st], (ks >> 1) & 1);
          for (int i = j; i < n_qt; ++i, ++pc) {
            // S/dP TMEM is free: the math threads arrived on bar_pr of the previous pair before we got here
            tc_fence_after();
            issue_sdp(j, i, st);
            mbar_wait(&bar_pr, pc & 1);        // P, dS of this pair are in smem
            tc_fence_after();
            const uint32_t aP = smem_u32(sP), aS = smem_u32(sdS), aQ = smem_u32(sQ + i * TILE), aO = smem_u32(sdO + i * TILE);
            const uint32_t aK = smem_u32(sKV + st * 2 * TILE);
            for (int kk = 0; kk < 8; ++kk)     // dV_j += P^T dO_i
              umma_bf16(tm_dV, umma_desc_sw128(aP + kk * 2048, TILE, 1024), umma_desc_sw128(aO + kk * 2048, TILE, 1024),
                        idesc_kv, (i > j) || kk != 0);
            for (int kk = 0; kk < 8; ++kk)     // dK_j += dS^T Q_i
              umma_bf16(tm_dK, umma_desc_sw128(aS + kk * 2048, TILE, 1024), umma_desc_sw128(aQ + kk * 2048, TILE, 1024),
                        idesc_kv, (i > j) || kk != 0);
            for (int kk = 0; kk < 8; ++kk)     // dQ_i += dS K_j
              umma_bf16(tm_dQ + 64 * i, umma_desc_sw128(aS + (kk >> 2) * TILE + (kk & 3) * 32, 16, 1024),
                        umma_desc_sw128(aK + kk * 2048, TILE, 1024), idesc_q, !(s == 0 && j == 0) || kk != 0);
            umma_commit(&bar_g);
            // P/dS smem and this pair's K/V must outlive the gradient UMMAs. (Issuing the next pair's S/dP before this
            // wait was measured slower: 1.41 ms vs 1.30 ms for the c4 self-attention layer.)
            mbar_wait(&bar_g, pc & 1);
            g_seen = pc + 1;
          }
          for (; g_seen < pc; ++g_seen) mbar_wait(&bar_g, g_seen & 1);   // phases must be consumed one by one
          if (ks + 2 < total_ks) issue_kv(ks + 2);
        }
      }
    }
  } else {
    // ================================================================== math warps
    const int r = tid & 127;          // row of the 128 x 128 pair tile (query row inside its tile / key row for dK, dV)
    const int half = tid >> 7;        // column half of the pair tile this thread handles
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    mbar_wait(&bar_q, 0);
    int pc = 0, g_seen = 0;
    for (int s = 0; s < p.n_src; ++s) {
      // D[row] = rowsum(dO * O_src): half 0 serves query tile 0, half 1 query tile 1, results shared through smem
      mbar_wait(&bar_o, s & 1);
      if (half < n_qt) {
        const int row = half * 128 + r;
        float acc = 0.f;
        if (row < p.T) {
          const uint8_t* orow = sP + half * TILE + r * 128;
          const uint8_t* drow = sdO + half * TILE + r * 128;
          for (int c = 0; c < p.hs / 8; ++c) {
            const uint4 a = *reinterpret_cast<const uint4*>(drow + ((c ^ (r & 7)) << 4));
            const uint4 o4 = *reinterpret_cast<const uint4*>(orow + ((c ^ (r & 7)) << 4));
            const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
            const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&o4);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              acc += __low2float(ah[e]) * __low2float(oh[e]) + __high2float(ah[e]) * __high2float(oh[e]);
          }
        }
        sD[half * 128 + r] = acc;
      }
      // named barrier over the 256 math threads only (the issuer warp never joins)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float D0 = sD[r], D1 = n_qt > 1 ? sD[128 + r] : 0.f;
      const float* lse_s = p.lse + ((((long long)s * p.G + g) * p.B + b) * p.H + h) * p.T;
      const float L0 = r < p.T ? lse_s[r] * 1.4426950408889634f : 0.f;
      const float L1 = (n_qt > 1 && 128 + r < p.T) ? lse_s[128 + r] * 1.4426950408889634f : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");   // sD is rewritten by the next source only after everyone has read it
      mbar_arrive(&bar_od);
      for (int j = 0; j < n_qt; ++j) {
        const int kv0 = j * 128;
        for (int i = j; i < n_qt; ++i, ++pc) {
          const int row = i * 128 + r;
          const bool row_ok = row < p.T;
          const float Dr = i == 0 ? D0 : D1, Lr = i == 0 ? L0 : L1;
          mbar_wait(&bar_sp, pc & 1);
          for (; g_seen < pc; ++g_seen) mbar_wait(&bar_g, g_seen & 1);   // complete by construction; keeps the phase in step
          tc_fence_after();
          uint32_t pp[32], dd[32];    // this thread's 64 columns of P and dS, packed bf16x2
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int col0 = kv0 + half * 64 + c * 32;
            // warp-uniform skip of chunks that are entirely above the diagonal for this warp's 32 rows
            if (col0 <= i * 128 + (warp & 3) * 32 + 31) {
              uint32_t sv[32], dv[32];
              tmem_ld_32x32(tm_S + lane_off + half * 64 + c * 32, sv);
              tmem_ld_32x32(tm_dP + lane_off + half * 64 + c * 32, dv);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
                if (row_ok && col0 + e <= row) {
                  p0 = exp2f(__uint_as_float(sv[e]) * p.scale_log2 - Lr);
                  d0 = p0 * (__uint_as_float(dv[e]) - Dr) * p.scale;
                }
                if (row_ok && col0 + e + 1 <= row) {
                  p1 = exp2f(__uint_as_float(sv[e + 1]) * p.scale_log2 - Lr);
                  d1 = p1 * (__uint_as_float(dv[e + 1]) - Dr) * p.scale;
                }
                __nv_bfloat162 ph = __floats2bfloat162_rn(p0, p1), dh = __floats2bfloat162_rn(d0, d1);
                pp[c * 16 + e / 2] = *reinterpret_cast<uint32_t*>(&ph);
                dd[c * 16 + e / 2] = *reinterpret_cast<uint32_t*>(&dh);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) pp[c * 16 + e] = dd[c * 16 + e] = 0u;
            }
          }
          // the gradient UMMAs of the previous pair (and, for the first pair of a source, the D pass) are done with sP/sdS:
          // the issuer waits on bar_g before it issues this pair's S/dP, so bar_sp above already implies it.
          const int off = half * TILE + r * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int lc = q ^ (r & 7);
            *reinterpret_cast<uint4*>(sP + off + (lc << 4)) = make_uint4(pp[4 * q], pp[4 * q + 1], pp[4 * q + 2], pp[4 * q + 3]);
            *reinterpret_cast<uint4*>(sdS + off + (lc << 4)) = make_uint4(dd[4 * q], dd[4 * q + 1], dd[4 * q + 2], dd[4 * q + 3]);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(&bar_pr);
        }
        // dV_j (half 0) and dK_j (half 1) of this source are complete after the last pair's gradient UMMAs
        for (; g_seen < pc; ++g_seen) mbar_wait(&bar_g, g_seen & 1);   // phases must be consumed one by one
        tc_fence_after();
        {
          const int kvrow = kv0 + r;
          const GradOut& go = half == 0 ? p.dv[s] : p.dk[s];
          const uint32_t src = (half == 0 ? tm_dV : tm_dK) + lane_off;
          float val[DP];
#pragma unroll
          for (int c = 0; c < DP / 16; ++c) {
            uint32_t v[16];
            tmem_ld_32x16(src + c * 16, v);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) val[c * 16 + e] = __uint_as_float(v[e]);
          }
          if (kvrow < p.T)
            store_row_bf16<DP>(go.ptr + (long long)g * go.gs + ((long long)b * p.T + kvrow) * go.ld + h * go.head_stride, val, p.hs);
          tc_fence_before();
        }
      }
    }
    // dQ tiles: half 0 stores tile 0, half 1 tile 1 (all gradient UMMAs have completed: bar_g of the last pair was waited)
    if (half < n_qt) {
      const int row = half * 128 + r;
      float val[DP];
#pragma unroll
      for (int c = 0; c < DP / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x16(tm_dQ + 64 * half + lane_off + c * 16, v);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) val[c * 16 + e] = __uint_as_float(v[e]);
      }
      if (row < p.T)
        store_row_bf16<DP>(p.dq.ptr + (long long)g * p.dq.gs + ((long long)b * p.T + row) * p.dq.ld + h * p.dq.head_stride, val, p.hs);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int DP>
static int launch_attn_bwd(const AttnBwdMaps& maps, const AttnBwdParams& p, cudaStream_t st) {
  constexpr int SMEM = 12 * 16384 + 80 + 1024;
  static bool attr = false;
  if (!attr) {
    TA_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr = true;
  }
  const long long grid = (long long)p.H * p.B * p.G;
  TA_CHECK_ARG(grid < (1ll << 31), "attention grid too large");
  attn_bwd_kernel<DP><<<(unsigned)grid, ATTB_THREADS, SMEM, st>>>(maps, p);
  TA_CHECK_LAUNCH();
  return 0;
}

static int grad_out(GradOut* o, const ta_attn_tensor* t, int G) {
  TA_CHECK_ARG(t && t->ptr && (reinterpret_cast<uintptr_t>(t->ptr) & 15) == 0, "attention gradient pointer must be 16-byte aligned");
  TA_CHECK_ARG(t->ld % 8 == 0 && t->head_stride % 8 == 0 && (G == 1 || t->gs % 8 == 0), "attention gradient strides must be multiples of 8");
  o->ptr = (__nv_bfloat16*)t->ptr; o->ld = t->ld; o->head_stride = t->head_stride; o->gs = t->gs;
  return 0;
}

}  // namespace ta

using namespace ta;

extern "C" int32_t ta_attention_bwd(const ta_attn_tensor* q, const ta_attn_tensor* k, const ta_attn_tensor* v,
                                    int32_t n_src, const void* o_bf16, const void* o_src_bf16, int64_t o_src_stride,
                                    const void* do_bf16, int64_t ldo, int64_t o_gs, const float* lse,
                                    const ta_attn_tensor* dq, const ta_attn_tensor* dk, const ta_attn_tensor* dv,
                                    int32_t G, int32_t B, int32_t T, int32_t H, int32_t hs, float dropout_p, uint64_t seed,
                                    void* stream) {
  (void)seed;
  TA_CHECK_ARG(q && k && v && o_bf16 && do_bf16 && lse && dq && dk && dv, "ta_attention_bwd: null pointer");
  TA_CHECK_ARG(n_src >= 1 && n_src <= ATTB_MAX_SRC, "ta_attention_bwd: n_src must be in [1, %d]", ATTB_MAX_SRC);
  TA_CHECK_ARG(n_src == 1 || o_src_bf16, "ta_attention_bwd: per-source outputs are required when n_src > 1");
  TA_CHECK_ARG(hs >= 8 && hs <= 64 && hs % 8 == 0, "ta_attention_bwd: head_size %d not supported", hs);
  TA_CHECK_ARG(T >= 1 && T <= 256, "ta_attention_bwd: T = %d not supported in this build (T <= 256)", T);
  TA_CHECK_ARG(dropout_p == 0.f, "ta_attention_bwd: attention dropout not implemented in this build");
  TA_CHECK_ARG(ldo % 8 == 0 && o_gs % 8 == 0, "ta_attention_bwd: output strides must be multiples of 8");
  AttnBwdMaps maps;
  if (make_head_map(&maps.q, q, hs, H, T, B, G, 128)) return 1;
  ta_attn_tensor dot{do_bf16, ldo, hs, o_gs};
  if (make_head_map(&maps.d_o, &dot, hs, H, T, B, G, 128)) return 1;
  // O tiles for D = rowsum(dO * O_src): self-attention reads o (group dim = modality); cross-attention (G == 1) reads
  // the per-source outputs, with the source index in the group dim of the tensor map
  TA_CHECK_ARG(n_src == 1 || G == 1, "ta_attention_bwd: n_src > 1 requires G == 1");
  if (n_src > 1) {
    ta_attn_tensor ot{o_src_bf16, ldo, hs, o_src_stride};
    if (make_head_map(&maps.o, &ot, hs, H, T, B, n_src, 128)) return 1;
  } else {
    ta_attn_tensor ot{o_bf16, ldo, hs, o_gs};
    if (make_head_map(&maps.o, &ot, hs, H, T, B, G, 128)) return 1;
  }
  AttnBwdParams p{};
  for (int s = 0; s < n_src; ++s) {
    if (make_head_map(&maps.k[s], &k[s], hs, H, T, B, G, 128)) return 1;
    if (make_head_map(&maps.v[s], &v[s], hs, H, T, B, G, 128)) return 1;
    if (grad_out(&p.dk[s], &dk[s], G) || grad_out(&p.dv[s], &dv[s], G)) return 1;
  }
  if (grad_out(&p.dq, dq, G)) return 1;
  p.T = T; p.B = B; p.H = H; p.G = G; p.hs = hs; p.n_src = n_src; p.n_qt = (T + 127) / 128;
  p.scale = (float)(1.0 / sqrt((double)hs));
  p.scale_log2 = (float)(1.4426950408889634 / sqrt((double)hs));
  p.o = (const __nv_bfloat16*)o_bf16;
  p.o_src = n_src > 1 ? (const __nv_bfloat16*)o_src_bf16 : nullptr;
  p.ldo = ldo; p.o_gs = o_gs; p.o_src_stride = o_src_stride;
  p.lse = lse;
  auto st = (cudaStream_t)stream;
  const int DP = hs <= 16 ? 16 : (hs <= 32 ? 32 : 64);
  if (DP == 16) return launch_attn_bwd<16>(maps, p, st);
  if (DP == 32) return launch_attn_bwd<32>(maps, p, st);
  return launch_attn_bwd<64>(maps, p, st);
}
