// K2 — grouped bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (UMMA 128 x BLOCK_N x 16,
// fp32 accumulators double-buffered in TMEM) -> tcgen05.ld epilogue with fused bias / tanh / ReLU / activation
// backward / residual add, fp32 and/or bf16 outputs. The epilogue moves its global traffic with TMA: the saved
// activation / residual tile is TMA-loaded into a swizzled per-warp staging buffer, combined in place with the
// accumulators (thread = row) and TMA-stored, so every global transaction is a full line (a row-per-thread store
// costs one LSU wavefront per 16 bytes and capped K=1024 GEMMs at ~0.6 of the mainloop rate). Operands that do not
// meet TMA's 16-byte pitch rules take the direct row-per-thread path. Persistent grid, warp-specialised:
//   warp 0    TMA producer (one lane)
//   warp 1    TMEM allocator + MMA issuer (one lane)
//   warps 2-5 epilogue (each owns the 32 TMEM lanes of its warp-id % 4 quarter)
// Replaces every nn.Linear forward/backward of /root/reference/model.py (see include/tradeaid_b200.h).
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "common.cuh"
#include "philox.cuh"
#include "ptx.cuh"

namespace ta {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // 64 bf16 = 128 bytes = one swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int SMEM_LIMIT = 227 * 1024 - 1024 - 512;  // opt-in maximum minus alignment slack and static barriers
constexpr int MAX_STAGES = 8;
constexpr int EPI_F32_BYTES = 32 * 32 * 4;           // one 32x32 fp32 chunk, 128-byte rows (SWIZZLE_128B)
constexpr int EPI_BF16_BYTES = 32 * 32 * 2;          // one 32x32 bf16 chunk, 64-byte rows (SWIZZLE_64B)

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // TMEM accumulator ring: all 512 columns (1 CTA per SM), up to 8 tiles deep for the narrow tiles so that the issuer
  // runs ahead of the epilogue round trip; two epilogue warps per lane quarter where the smem budget allows it
  static constexpr int NACC = 512 / BLOCK_N > 8 ? 8 : 512 / BLOCK_N;
  static constexpr int TMEM_COLS = 512;
  static constexpr int NW = BLOCK_N <= 128 ? 8 : 4;
  static constexpr int THREADS = (2 + NW) * 32;
};

struct GemmParams {
  int M, N, K, G0, G1;
  int a_mn, b_mn;
  int m_blocks, n_blocks, k_blocks;
  float alpha;
  int act, aux_mode;
  const float* bias;
  long long bias_gs0, bias_gs1;
  const __nv_bfloat16* aux;
  long long ld_aux, aux_gs0, aux_gs1;
  const float* res;
  long long ld_res, res_gs0, res_gs1;
  float* out_f32;
  long long ld_f32, f32_gs0, f32_gs1;
  __nv_bfloat16* out_bf16;
  long long ld_bf16, bf16_gs0, bf16_gs1;
  int vec_f32, vec_bf16, vec_res, vec_aux, vec_bias;
  int stages;          // depth of the A/B smem ring
  int tma_epi;         // 1: TMA-staged epilogue (maps below are valid)
  int nbf, nbb;        // staging buffers per epilogue warp: fp32 chunks / bf16 chunks
  int nst;             // rotation depth of the store-only buffers (2, or 4 for short-K problems bound by the store path)
  int bf16_direct;     // tma_epi with both outputs: bf16 goes out with row-per-thread stores
  int bf16_wide;       // bf16-only output staged two chunks at a time: [32 x 64] boxes, 128-byte rows (nbb / 2 buffers of 4 KB)
  int epi_warp_bytes;
  int m_major;         // tile order, see decode_tile
  uint32_t drop_thresh;     // 16-bit lane threshold of the epilogue dropout (0 = off), see philox.cuh
  float drop_scale;         // 1 / (1 - p)
  unsigned long long drop_seed;
  const unsigned long long* drop_step;
  int debug;           // TA_GEMM_DEBUG (dev only): 1 = skip global stores, 2 = skip the epilogue math and stores
};

struct EpiMaps {
  CUtensorMap f32, bf16, res, aux;
  CUtensorMap bf16w;   // the bf16 output again with [32 rows x 64 columns] boxes (SWIZZLE_128B), see GemmParams::bf16_wide
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// tile index -> (group, m0, n0). Default order: group-major (tiles of a group are neighbours: its operands stay hot in
// L2). m_major: row-block-major, for the block-diagonal q/k/v stage-2 problems (tens to hundreds of groups that are
// column slices of the SAME rows, K <= 256): concurrently running tiles then read / write adjacent segments of the same rows.
__device__ __forceinline__ void decode_tile(const GemmParams& p, int tile, int tiles_per_group, int block_n, int block_m, int& g,
                                            int& m0, int& n0) {
  if (p.m_major) {
    const int per_m = p.n_blocks * p.G0 * p.G1;
    const int mb = tile / per_m, rem = tile - mb * per_m;
    g = rem / p.n_blocks;
    m0 = mb * block_m;
    n0 = (rem - g * p.n_blocks) * block_n;
  } else {
    g = tile / tiles_per_group;
    const int t = tile - g * tiles_per_group;
    m0 = (t / p.n_blocks) * block_m;
    n0 = (t % p.n_blocks) * block_n;
  }
}

// Epilogue of one warp (warps 2-5; warp % 4 = TMEM lane quarter): shared by the 1-CTA and the 2-CTA kernels.
//   epi_smem    shared-window address of the staging buffers (4 warps x p.epi_warp_bytes)
//   tmem_empty0 address of tmem_empty_bar[0] of the CTA whose issuer warp waits on it: the own CTA's (m_off == 0, plain
//               shared address) or the leader's through the cluster window (PAIR peer, m_off != 0)
//   m_off       row offset of this CTA inside the tile (PAIR: rank * 128)
template <int BLOCK_N, bool PAIR, int NW, int NACC>
__device__ __forceinline__ void gemm_epilogue(const GemmParams& p, const EpiMaps& em, uint32_t epi_smem, uint64_t* tmem_full_bar,
                                              uint32_t tmem_empty0, uint64_t (*epi_bar)[3], uint32_t tmem_base, int warp, int lane,
                                              int first_tile, int tile_stride, int total_tiles, int tiles_per_group, int m_off) {
  // NW = 4: one warp per TMEM lane quarter, all columns. NW = 8: two warps per quarter, each half of the columns.
  const int quarter = warp & 3, ep = warp - 2;
  constexpr int NCH_W = (BLOCK_N / 32) / (NW / 4);     // chunks per warp
  const int c0 = (ep >> 2) * NCH_W;
  int local = 0;
  if (p.tma_epi) {
    // ---- TMA-staged epilogue. Per warp: nbf fp32 chunk buffers (SWIZZLE_128B) then nbb bf16 chunk buffers
    // (SWIZZLE_64B). Thread = accumulator row; 16-byte smem accesses hit swizzled chunk (j ^ row bits): conflict-free.
    const bool has_res = p.res != nullptr, has_aux = p.aux != nullptr, has_f32 = p.out_f32 != nullptr,
               has_bf = p.out_bf16 != nullptr;
    const bool has_load = has_res || has_aux;
    const int nl = has_res ? p.nbf : p.nbb;          // rotation depth of the loaded buffer family
    const int dbg = p.debug, act = p.act, nst = p.nst;
    const float alpha = p.alpha;
    const bool aux_tanh = p.aux_mode == TA_AUX_TANH_BWD;
    const uint32_t wbase = epi_smem + ep * p.epi_warp_bytes;
    const uint32_t fbuf0 = wbase, bbuf0 = wbase + p.nbf * EPI_F32_BYTES;
    const uint32_t lbar0 = smem_u32(&epi_bar[ep][0]);
    const uint32_t frow = lane * 128, fsw = lane & 7;            // fp32 row offset, swizzle key
    const uint32_t brow = lane * 64, bsw = (lane >> 1) & 3;      // bf16 row offset, swizzle key
    uint32_t lphase = 0;    // bit b: parity the next wait on loaded buffer b uses
    uint32_t sc = 0;        // running count of chunks stored (rotation of the store-only buffers)
    // wide bf16 staging: two neighbouring chunks share one [32 x 64] buffer with 128-byte rows and leave with ONE TMA store
    // (full 128-byte lines; a 32-column box writes 64-byte row pieces and the store path, not the math, paced the short-K
    // launches: profiles/r02_ncu stage-2 capture). Every warp owns an even number of chunks and N % 64 == 0 in this mode.
    const bool wide = p.bf16_wide != 0;
    const int nw = p.nbb >> 1;
    int wb = 0;             // wide buffer being filled
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++local) {
      int g, m0, n0;
      decode_tile(p, tile, tiles_per_group, BLOCK_N, PAIR ? 2 * BLOCK_M : BLOCK_M, g, m0, n0);
      m0 += m_off;
      const int g0 = g % p.G0, g1 = g / p.G0;
      const int acc = local % NACC;
      const uint32_t acc_phase = (local / NACC) & 1;
      const int row0 = m0 + quarter * 32;
      constexpr int NCH = BLOCK_N / 32;
      int nch = (p.N - n0 + 31) / 32;
      if (nch > NCH) nch = NCH;
      const int c1 = nch < c0 + NCH_W ? nch : c0 + NCH_W;   // this warp's chunks: [c0, c1)
      int next_load = c0;
      // residual / saved-activation chunk q of this tile -> loaded buffer (q - c0) % nl
      auto issue_load = [&](int q) {
        const int b = (q - c0) % nl;
        const uint32_t bar = lbar0 + b * 8;
        if (has_res) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(EPI_F32_BYTES) : "memory");
          tma_load_4d_u32(fbuf0 + b * EPI_F32_BYTES, &em.res, bar, n0 + q * 32, row0, g0, g1);
        } else {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(EPI_BF16_BYTES) : "memory");
          tma_load_4d_u32(bbuf0 + b * EPI_BF16_BYTES, &em.aux, bar, n0 + q * 32, row0, g0, g1);
        }
      };
      if (has_load && row0 < p.M && c0 < c1) {
        if (lane == 0) {
          bulk_wait_read<0>();     // every staging buffer of the previous tile has been read by its store
          for (; next_load < c1 && next_load < c0 + nl; ++next_load) issue_load(next_load);
        }
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      const float* bias = p.bias ? p.bias + g0 * p.bias_gs0 + g1 * p.bias_gs1 : nullptr;
      __nv_bfloat16* ob = p.bf16_direct ? p.out_bf16 + g0 * p.bf16_gs0 + g1 * p.bf16_gs1 + (long long)row * p.ld_bf16 : nullptr;
      const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
      if (row0 >= p.M || c0 >= c1) {   // this warp's rows lie beyond M or its columns beyond N: nothing to store
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR && m_off != 0) mbar_arrive_cluster(tmem_empty0 + 8 * acc); else mbar_arrive_a(tmem_empty0 + 8 * acc); }
        continue;
      }
      // The chunk loop stays ROLLED: unrolled x8 the kernel was ~0.5 MB of SASS and the epilogue warps spent most of
      // their time in instruction-fetch stalls (ncu: stall_no_inst on nearly every line).
      uint32_t v[32];
      tmem_ld_32x32(taddr_row + c0 * 32, v);
      int lb = 0;             // loaded buffer of this chunk ((c - c0) % nl)
#pragma unroll 1
      for (int c = c0; c < c1; ++c) {
        {
          const int col0 = n0 + c * 32;
          const bool full = col0 + 32 <= p.N;
          float4 b4[8];
          const bool vb = bias && full && p.vec_bias;
          if (vb) {
#pragma unroll
            for (int i = 0; i < 8; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(bias + col0) + i);
          }
          tmem_wait_ld();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * alpha;
          if (c + 1 < c1) {   // v is dead: the next chunk's TMEM read overlaps the rest of this chunk
            tmem_ld_32x32(taddr_row + (c + 1) * 32, v);
          } else {
            // all TMEM reads of this accumulator are done: hand it back to the MMA warp before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR && m_off != 0) mbar_arrive_cluster(tmem_empty0 + 8 * acc); else mbar_arrive_a(tmem_empty0 + 8 * acc); }
          }
          const uint32_t fbuf = fbuf0 + (has_res ? lb : (int)(sc & (uint32_t)(nst - 1))) * EPI_F32_BYTES;
          const uint32_t bbuf = bbuf0 + (has_aux ? lb : (int)(sc & (uint32_t)(nst - 1))) * EPI_BF16_BYTES;
          if (dbg != 2) {
            if (vb) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                f[4 * i] += b4[i].x; f[4 * i + 1] += b4[i].y; f[4 * i + 2] += b4[i].z; f[4 * i + 3] += b4[i].w;
              }
            } else if (bias) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (full || col0 + i < p.N) f[i] += __ldg(bias + col0 + i);
            }
            if (act == TA_ACT_TANH) {
              if (has_f32) {   // fp32 consumers get the accurate tanh
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = tanhf(f[i]);
              } else {         // bf16-only output: MUFU tanh (rel. error 2^-11) sits below the 2^-9 rounding of the
                               // store; tanhf costs ~60 issue cycles per warp and made the epilogue the bottleneck
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = tanh_approx(f[i]);
              }
            } else if (act == TA_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (p.drop_thresh) {   // inverted dropout on act(acc + bias), before the residual add (model.py:91, 116, 171)
              const unsigned long long sd = p.drop_seed + (p.drop_step ? *p.drop_step * 0x9E3779B97F4A7C15ull : 0ull);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t m = dropout_keep8(sd, 0u, (uint32_t)g, (uint32_t)row, (uint32_t)(col0 >> 3) + i, p.drop_thresh);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[8 * i + e] = (m >> e) & 1u ? f[8 * i + e] * p.drop_scale : 0.f;
              }
            }
            if (has_load) {
              mbar_wait(&epi_bar[ep][lb], (lphase >> lb) & 1);
              lphase ^= 1u << lb;
            }
            if (has_aux) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 u = dbg == 3 ? make_uint4(0, 0, 0, 0) : lds128(bbuf + brow + ((j ^ bsw) << 4));
                const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&u);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const float a = __bfloat162float(h[k]);
                  f[8 * j + k] = aux_tanh ? f[8 * j + k] * (1.f - a * a) : (a > 0.f ? f[8 * j + k] : 0.f);
                }
              }
            }
            if (has_res) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 u = dbg == 3 ? make_uint4(0, 0, 0, 0) : lds128(fbuf + frow + ((j ^ fsw) << 4));
                f[4 * j] += __uint_as_float(u.x); f[4 * j + 1] += __uint_as_float(u.y);
                f[4 * j + 2] += __uint_as_float(u.z); f[4 * j + 3] += __uint_as_float(u.w);
              }
            }
            if (has_f32 && dbg != 4) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sts128(fbuf + frow + ((j ^ fsw) << 4),
                       make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                                  __float_as_uint(f[4 * j + 3])));
            }
            if (has_bf) {
              uint4 u[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u[j]);
#pragma unroll
                for (int k = 0; k < 4; ++k) h2[k] = __floats2bfloat162_rn(f[8 * j + 2 * k], f[8 * j + 2 * k + 1]);
              }
              if (wide) {
                const uint32_t wbuf = bbuf0 + wb * (2 * EPI_BF16_BYTES) + lane * 128, hf = ((c - c0) & 1) << 2;
#pragma unroll
                for (int j = 0; j < 4; ++j) sts128(wbuf + (((hf + j) ^ (lane & 7)) << 4), u[j]);
              } else if (!p.bf16_direct) {
                if (dbg != 4)
#pragma unroll
                for (int j = 0; j < 4; ++j) sts128(bbuf + brow + ((j ^ bsw) << 4), u[j]);
              } else if (row_ok && dbg != 1) {
                if (full && p.vec_bf16) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(ob + col0 + 8 * j) = u[j];
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i)
                    if (col0 + i < p.N) ob[col0 + i] = __float2bfloat16_rn(f[i]);
                }
              }
            }
          } else if (has_load) {
            mbar_wait(&epi_bar[ep][lb], (lphase >> lb) & 1);
            lphase ^= 1u << lb;
          }
          if (wide) {
            if ((c - c0) & 1) {       // second chunk of the pair: one store of the [32 x 64] buffer
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (dbg == 0) tma_store_4d(&em.bf16w, bbuf0 + wb * (2 * EPI_BF16_BYTES), col0 - 32, row0, g0, g1);
                bulk_commit_group();
                // the buffer the next pair writes has been drained by its store (nw - 1 younger stores may be in flight)
                if (nw == 3) bulk_wait_read<2>(); else bulk_wait_read<1>();
              }
              if (++wb == nw) wb = 0;
              __syncwarp();
            }
            ++sc;
            continue;
          }
          if (dbg != 5) fence_proxy_async_smem();   // generic-proxy writes above -> visible to the TMA store
          __syncwarp();
          if (lane == 0) {
            if (dbg == 0 || dbg == 5) {
              if (has_f32) tma_store_4d(&em.f32, fbuf, col0, row0, g0, g1);
              if (has_bf && !p.bf16_direct) tma_store_4d(&em.bf16, bbuf, col0, row0, g0, g1);
            }
            bulk_commit_group();
            // the buffer the next chunk writes has been drained by its store (nst - 1 younger stores may be in flight)
            if (nst == 4) bulk_wait_read<3>(); else bulk_wait_read<1>();
            if (has_load)
              for (; next_load < c1 && next_load <= c + nl - 1; ++next_load) issue_load(next_load);
          }
          ++sc;
          if (++lb == nl) lb = 0;
          __syncwarp();
        }
      }
    }
    if (lane == 0) bulk_wait_all<0>();
  } else
  {
  for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++local) {
    int g, m0, n0;
    decode_tile(p, tile, tiles_per_group, BLOCK_N, PAIR ? 2 * BLOCK_M : BLOCK_M, g, m0, n0);
      m0 += m_off;
    const int g0 = g % p.G0, g1 = g / p.G0;
    const int acc = local % NACC;
    const uint32_t acc_phase = (local / NACC) & 1;
    mbar_wait(&tmem_full_bar[acc], acc_phase);
    tc_fence_after();
    const int row = m0 + quarter * 32 + lane;
    const bool row_ok = row < p.M;
    const float* bias = p.bias ? p.bias + g0 * p.bias_gs0 + g1 * p.bias_gs1 : nullptr;
    const __nv_bfloat16* aux = p.aux ? p.aux + g0 * p.aux_gs0 + g1 * p.aux_gs1 + (long long)row * p.ld_aux : nullptr;
    const float* res = p.res ? p.res + g0 * p.res_gs0 + g1 * p.res_gs1 + (long long)row * p.ld_res : nullptr;
    float* of = p.out_f32 ? p.out_f32 + g0 * p.f32_gs0 + g1 * p.f32_gs1 + (long long)row * p.ld_f32 : nullptr;
    __nv_bfloat16* ob = p.out_bf16 ? p.out_bf16 + g0 * p.bf16_gs0 + g1 * p.bf16_gs1 + (long long)row * p.ld_bf16 : nullptr;
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
    constexpr int NCH = BLOCK_N / 32;
    int nch = (p.N - n0 + 31) / 32;
    if (nch > NCH) nch = NCH;
    // rolled chunk loop (code size, see above); the TMEM read of chunk c+1 overlaps the processing of chunk c
    const int c1 = nch < c0 + NCH_W ? nch : c0 + NCH_W;
    uint32_t v[32];
    if (c0 < c1) tmem_ld_32x32(taddr_row + c0 * 32, v);
#pragma unroll 1
    for (int c = c0; c < c1; ++c) {
      {
        const int col0 = n0 + c * 32;
        const bool full = col0 + 32 <= p.N;
        // bias for these 32 columns: 8 independent 128-bit broadcast loads, issued before waiting on TMEM
        float4 b4[8];
        const bool vb = bias && full && p.vec_bias;
        if (vb) {
#pragma unroll
          for (int i = 0; i < 8; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(bias + col0) + i);
        }
        tmem_wait_ld();
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * p.alpha;
        if (c + 1 < c1) tmem_ld_32x32(taddr_row + (c + 1) * 32, v);
        if (row_ok) {
          if (vb) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              f[4 * i] += b4[i].x; f[4 * i + 1] += b4[i].y; f[4 * i + 2] += b4[i].z; f[4 * i + 3] += b4[i].w;
            }
          } else if (bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (full || col0 + i < p.N) f[i] += __ldg(bias + col0 + i);
          }
          if (p.act == TA_ACT_TANH) {
            if (of) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = tanhf(f[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = tanh_approx(f[i]);
            }
          } else if (p.act == TA_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          if (p.drop_thresh) {   // inverted dropout on act(acc + bias), before the residual add (model.py:91, 116, 171)
            const unsigned long long sd = p.drop_seed + (p.drop_step ? *p.drop_step * 0x9E3779B97F4A7C15ull : 0ull);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t m = dropout_keep8(sd, 0u, (uint32_t)g, (uint32_t)row, (uint32_t)(col0 >> 3) + i, p.drop_thresh);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[8 * i + e] = (m >> e) & 1u ? f[8 * i + e] * p.drop_scale : 0.f;
            }
          }
          if (aux) {
            if (full && p.vec_aux) {
              uint4 u4[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) u4[i] = *reinterpret_cast<const uint4*>(aux + col0 + 8 * i);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&u4[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float a = __bfloat162float(h[j]);
                  f[8 * i + j] = p.aux_mode == TA_AUX_TANH_BWD ? f[8 * i + j] * (1.f - a * a) : (a > 0.f ? f[8 * i + j] : 0.f);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) {
                  const float a = __bfloat162float(aux[col0 + i]);
                  f[i] = p.aux_mode == TA_AUX_TANH_BWD ? f[i] * (1.f - a * a) : (a > 0.f ? f[i] : 0.f);
                }
            }
          }
          if (res) {
            if (full && p.vec_res) {
              float4 r4[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) r4[i] = *reinterpret_cast<const float4*>(res + col0 + 4 * i);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                f[4 * i] += r4[i].x; f[4 * i + 1] += r4[i].y; f[4 * i + 2] += r4[i].z; f[4 * i + 3] += r4[i].w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) f[i] += res[col0 + i];
            }
          }
          if (of) {
            if (full && p.vec_f32) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(of + col0 + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) of[col0 + i] = f[i];
            }
          }
          if (ob) {
            if (full && p.vec_bf16) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 u;
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(f[i + 2 * j], f[i + 2 * j + 1]);
                *reinterpret_cast<uint4*>(ob + col0 + i) = u;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) ob[col0 + i] = __float2bfloat16_rn(f[i]);
            }
          }
        }
      }
    }
    // release this accumulator stage back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { if (PAIR && m_off != 0) mbar_arrive_cluster(tmem_empty0 + 8 * acc); else mbar_arrive_a(tmem_empty0 + 8 * acc); }
  }
  }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(GemmCfg<BLOCK_N>::THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ EpiMaps em, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  const int STAGES = p.stages;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  constexpr int NACC = Cfg::NACC;
  __shared__ __align__(8) uint64_t tmem_full_bar[NACC];
  __shared__ __align__(8) uint64_t tmem_empty_bar[NACC];
  __shared__ __align__(8) uint64_t epi_bar[Cfg::NW][3];
  __shared__ uint32_t tmem_base_smem;

  // the shuffle tells the compiler that `warp` is warp-uniform: producer and issuer then keep their TMA / UMMA operands
  // in uniform registers (no per-instruction R2UR waterfall)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], Cfg::NW);
    }
    for (int s = 0; s < 3 * Cfg::NW; ++s) mbar_init(&epi_bar[0][0] + s, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // PDL: everything above is on-chip set-up (barriers, TMEM, descriptor prefetch); from here on the kernel reads what
  // the previous kernel of the stream wrote
  pdl_launch_dependents();
  pdl_wait();

  const int tiles_per_group = p.m_blocks * p.n_blocks;
  const int total_tiles = tiles_per_group * p.G0 * p.G1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    {
      const uint32_t smem_a = smem_u32(smem), full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int g, m0, n0;
        decode_tile(p, tile, tiles_per_group, BLOCK_N, BLOCK_M, g, m0, n0);
        const int g0 = g % p.G0, g1 = g / p.G0;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait_a(empty0 + 8 * stage, phase ^ 1);
          const uint32_t fb = full0 + 8 * stage;
          mbar_expect_tx_ws(fb, Cfg::STAGE_BYTES);
          const uint32_t sA = smem_a + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + A_STAGE_BYTES;
          const int k0 = kb * BLOCK_K;
          if (!p.a_mn) {
            tma_load_4d_ws(sA, &tmA, fb, k0, m0, g0, g1);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_4d_ws(sA + c * (64 * BLOCK_K * 2), &tmA, fb, m0 + c * 64, k0, g0, g1);
          }
          if (!p.b_mn) {
            tma_load_4d_ws(sB, &tmB, fb, k0, n0, g0, g1);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_N / 64; ++c)
              tma_load_4d_ws(sB + c * (64 * BLOCK_K * 2), &tmB, fb, n0 + c * 64, k0, g0, g1);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    const uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, p.a_mn, p.b_mn);
    // K-major SW128: 8-row groups 1024 B apart (SBO), 16 k-elements = 32 B inside the swizzle row.
    // MN-major SW128: boxes of [64 k-rows][64 mn] (8 KB): next 64-mn chunk = LBO 8192, next 8 k-rows = SBO 1024,
    //                 16 k-rows = 2048 B.
    // The address field of a descriptor sits in its low bits: stepping an operand = adding (bytes >> 4).
    const uint64_t da0 = umma_desc_sw128(0, p.a_mn ? 64 * BLOCK_K * 2 : 16, 1024);
    const uint64_t db0 = umma_desc_sw128(0, p.b_mn ? 64 * BLOCK_K * 2 : 16, 1024);
    const uint32_t a_kstep = (p.a_mn ? 16 * 128 : 32) >> 4, b_kstep = (p.b_mn ? 16 * 128 : 32) >> 4;
    const uint32_t smem_a = smem_u32(smem), full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]), tempty0 = smem_u32(&tmem_empty_bar[0]);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local % NACC;
      const uint32_t acc_phase = (local / NACC) & 1;
      mbar_wait_a(tempty0 + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait_a(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sA = smem_a + stage * Cfg::STAGE_BYTES;
        const uint64_t da = da0 + ((sA & 0x3FFFF) >> 4), db = db0 + (((sA + A_STAGE_BYTES) & 0x3FFFF) >> 4);
        const int krem = p.K - kb * BLOCK_K;
        const int nk = krem >= BLOCK_K ? BLOCK_K / 16 : (krem + 15) / 16;
        for (int j = 0; j < nk; ++j) umma_bf16_ws(tmem_d, da + j * a_kstep, db + j * b_kstep, idesc, (kb | j) != 0);
        umma_commit_ws(empty0 + 8 * stage);
        if (kb == p.k_blocks - 1) umma_commit_ws(tfull0 + 8 * acc);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    gemm_epilogue<BLOCK_N, false, Cfg::NW, NACC>(p, em, smem_u32(smem) + STAGES * Cfg::STAGE_BYTES, tmem_full_bar, smem_u32(&tmem_empty_bar[0]),
                                  epi_bar, tmem_base, warp, lane, blockIdx.x, gridDim.x, total_tiles, tiles_per_group, 0);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ 2-CTA kernel
// CTA pair (cluster of 2, tcgen05 cta_group::2): one 256 x 256 output tile per pair. Each CTA owns 128 rows (its A tile,
// its TMEM accumulator, its epilogue) and stages only HALF of the B tile (128 of the 256 n-rows): the UMMA of the leader
// CTA reads both halves. L2 -> smem traffic per output tile drops from 48 KB to 32 KB per k-block and CTA, which is
// what bounds the 1-CTA kernel (~6300 B/clk L2 throughput chip-wide: 148 x 48 KB per 512 tensor-pipe cycles is 2x that).
//   leader (rank 0): producer, UMMA issuer, epilogue        peer (rank 1): producer, epilogue
//   full_bar[s]   (leader's): both CTAs' TMA loads complete_tx on it; the leader's producer expects 2 x STAGE bytes
//   empty_bar[s], tmem_full_bar[a] (both CTAs): signalled by the leader's multicast tcgen05.commit
//   tmem_empty_bar[a] (leader's): 16 arrivals, 8 epilogue warps of each CTA (the peer's arrive through the cluster window)
constexpr int G2_BLOCK_N = 256;
constexpr int G2_B_HALF = G2_BLOCK_N / 2;
constexpr int G2_STAGE_BYTES = A_STAGE_BYTES + G2_B_HALF * BLOCK_K * 2;   // 32 KB per CTA

constexpr int GEMM2_THREADS = 320;   // producer, issuer, 8 epilogue warps (two per TMEM lane quarter)

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM2_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ EpiMaps em, const GemmParams p) {
  const int STAGES = p.stages;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t epi_bar[8][3];
  __shared__ uint32_t tmem_base_smem;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 16);
    }
    for (int s = 0; s < 24; ++s) mbar_init(&epi_bar[0][0] + s, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(&tmem_base_smem, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();       // PDL: on-chip set-up done; wait for the previous kernel before touching global memory
  pdl_wait();

  const int m_pairs = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int tiles_per_group = m_pairs * p.n_blocks;
  const int total_tiles = tiles_per_group * p.G0 * p.G1;
  const int first_tile = blockIdx.x >> 1, tile_stride = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    const uint32_t smem_a = smem_u32(smem), empty0 = smem_u32(&empty_bar[0]);
    const uint32_t full0_leader = mapa_u32(smem_u32(&full_bar[0]), 0);   // the leader's full barriers
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
      int g, m0, n0;
      decode_tile(p, tile, tiles_per_group, G2_BLOCK_N, 2 * BLOCK_M, g, m0, n0);
      m0 += rank * BLOCK_M;
      n0 += rank * G2_B_HALF;
      const int g0 = g % p.G0, g1 = g / p.G0;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait_a(empty0 + 8 * stage, phase ^ 1);
        const uint32_t fb = full0_leader + 8 * stage;
        if (rank == 0) mbar_expect_tx_ws(smem_u32(&full_bar[stage]), 2 * G2_STAGE_BYTES);
        const uint32_t sA = smem_a + stage * G2_STAGE_BYTES;
        const uint32_t sB = sA + A_STAGE_BYTES;
        const int k0 = kb * BLOCK_K;
        if (!p.a_mn) {
          tma_load_4d_2cta_ws(sA, &tmA, fb, k0, m0, g0, g1);
        } else {
#pragma unroll
          for (int c = 0; c < BLOCK_M / 64; ++c)
            tma_load_4d_2cta_ws(sA + c * (64 * BLOCK_K * 2), &tmA, fb, m0 + c * 64, k0, g0, g1);
        }
        if (!p.b_mn) {
          tma_load_4d_2cta_ws(sB, &tmB, fb, k0, n0, g0, g1);
        } else {
#pragma unroll
          for (int c = 0; c < G2_B_HALF / 64; ++c)
            tma_load_4d_2cta_ws(sB + c * (64 * BLOCK_K * 2), &tmB, fb, n0 + c * 64, k0, g0, g1);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(2 * BLOCK_M, G2_BLOCK_N, p.a_mn, p.b_mn);
      const uint64_t da0 = umma_desc_sw128(0, p.a_mn ? 64 * BLOCK_K * 2 : 16, 1024);
      const uint64_t db0 = umma_desc_sw128(0, p.b_mn ? 64 * BLOCK_K * 2 : 16, 1024);
      const uint32_t a_kstep = (p.a_mn ? 16 * 128 : 32) >> 4, b_kstep = (p.b_mn ? 16 * 128 : 32) >> 4;
      const uint32_t smem_a = smem_u32(smem), full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]), tempty0 = smem_u32(&tmem_empty_bar[0]);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait_a(tempty0 + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * G2_BLOCK_N;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait_a(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sA = smem_a + stage * G2_STAGE_BYTES;
          const uint64_t da = da0 + ((sA & 0x3FFFF) >> 4), db = db0 + (((sA + A_STAGE_BYTES) & 0x3FFFF) >> 4);
          const int krem = p.K - kb * BLOCK_K;
          const int nk = krem >= BLOCK_K ? BLOCK_K / 16 : (krem + 15) / 16;
          for (int j = 0; j < nk; ++j) umma_bf16_2cta_ws(tmem_d, da + j * a_kstep, db + j * b_kstep, idesc, (kb | j) != 0);
          umma_commit_2cta_ws(empty0 + 8 * stage, 3);
          if (kb == p.k_blocks - 1) umma_commit_2cta_ws(tfull0 + 8 * acc, 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, their own 128 rows)
    // the leader's warps arrive on their own barrier (CTA scope), the peer's through the cluster window
    gemm_epilogue<G2_BLOCK_N, true, 8, 2>(p, em, smem_u32(smem) + STAGES * G2_STAGE_BYTES, tmem_full_bar,
                                    rank == 0 ? smem_u32(&tmem_empty_bar[0]) : mapa_u32(smem_u32(&tmem_empty_bar[0]), 0), epi_bar, tmem_base, warp, lane, first_tile,
                                    tile_stride, total_tiles, tiles_per_group, rank * BLOCK_M);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // no CTA leaves (or frees TMEM) while its peer may still touch its smem / barriers
  if (warp == 1) tmem_dealloc_2cta(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

// operand: logical (MN x K) bf16 matrix with two group dims; box_mn = rows of the MN dim per smem tile
static int make_operand_map(CUtensorMap* out, const ta_operand* op, int64_t MN, int64_t K, int G0, int G1, int box_mn) {
  EncodeTiledFn enc = get_encode_fn();
  TA_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  TA_CHECK_ARG((reinterpret_cast<uintptr_t>(op->ptr) & 15) == 0, "gemm operand pointer must be 16-byte aligned");
  TA_CHECK_ARG(op->ld % 8 == 0, "gemm operand ld (%lld) must be a multiple of 8 elements", (long long)op->ld);
  TA_CHECK_ARG(G0 == 1 || op->gs0 % 8 == 0, "gemm operand gs0 (%lld) must be a multiple of 8", (long long)op->gs0);
  TA_CHECK_ARG(G1 == 1 || op->gs1 % 8 == 0, "gemm operand gs1 (%lld) must be a multiple of 8", (long long)op->gs1);
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  cuuint32_t box[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const cuuint64_t safe = 16;
  if (op->major == 0) {  // stored [MN][K]
    TA_CHECK_ARG(op->ld >= K, "K-major operand ld < K");
    dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)MN;
    box[0] = BLOCK_K; box[1] = (cuuint32_t)box_mn;
  } else {               // stored [K][MN]
    TA_CHECK_ARG(op->ld >= MN, "MN-major operand ld < MN");
    dims[0] = (cuuint64_t)MN; dims[1] = (cuuint64_t)K;
    box[0] = 64; box[1] = BLOCK_K;
  }
  dims[2] = (cuuint64_t)G0; dims[3] = (cuuint64_t)G1;
  box[2] = 1; box[3] = 1;
  strides[0] = (cuuint64_t)op->ld * 2;
  strides[1] = G0 > 1 ? (cuuint64_t)op->gs0 * 2 : safe;
  strides[2] = G1 > 1 ? (cuuint64_t)op->gs1 * 2 : safe;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(op->ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TA_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (dims %llu %llu %llu %llu ld %lld)",
               (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
               (unsigned long long)dims[3], (long long)op->ld);
  return 0;
}

// epilogue tile map: logical [rows x cols] matrix of `es`-byte elements with two group dims, 32x32 boxes.
// Returns false when the operand does not satisfy TMA's alignment rules (caller takes the direct path).
static bool make_epi_map(CUtensorMap* out, const void* ptr, long long ld, long long gs0, long long gs1, int64_t rows,
                         int64_t cols, int G0, int G1, bool is_f32, bool wide_bf16 = false) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc || !ptr) return false;
  const int es = is_f32 ? 4 : 2;
  const int q = 16 / es;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ld % q || ld < cols) return false;
  if ((G0 > 1 && (gs0 % q || gs0 == 0)) || (G1 > 1 && (gs1 % q || gs1 == 0))) return false;
  cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)G0, (cuuint64_t)G1};
  cuuint64_t strides[3] = {(cuuint64_t)ld * es, G0 > 1 ? (cuuint64_t)gs0 * es : 16, G1 > 1 ? (cuuint64_t)gs1 * es : 16};
  cuuint32_t box[4] = {wide_bf16 ? 64u : 32u, 32, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   is_f32 || wide_bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Epilogue plan: which outputs go through TMA staging, how many staging buffers, and how many A/B ring stages are left.
static void plan_epilogue(GemmParams& p, EpiMaps& em, int stage_bytes, int ep_warps, int chunks_per_warp) {
  static const int force_legacy = env_int("TA_GEMM_LEGACY_EPI", 0);
  static const int force_nl = env_int("TA_GEMM_NLOAD", 0);
  memset(&em, 0, sizeof(em));
  const bool need_f = p.res || p.out_f32, need_b = p.aux || p.out_bf16;
  bool tma_ok = !force_legacy && !(p.aux && need_f);
  if (tma_ok && p.out_f32) tma_ok = make_epi_map(&em.f32, p.out_f32, p.ld_f32, p.f32_gs0, p.f32_gs1, p.M, p.N, p.G0, p.G1, true);
  if (tma_ok && p.res) tma_ok = make_epi_map(&em.res, p.res, p.ld_res, p.res_gs0, p.res_gs1, p.M, p.N, p.G0, p.G1, true);
  if (tma_ok && p.aux) tma_ok = make_epi_map(&em.aux, p.aux, p.ld_aux, p.aux_gs0, p.aux_gs1, p.M, p.N, p.G0, p.G1, false);
  p.bf16_direct = need_f && p.out_bf16 ? 1 : 0;
  if (tma_ok && p.out_bf16 && !p.bf16_direct)
    tma_ok = make_epi_map(&em.bf16, p.out_bf16, p.ld_bf16, p.bf16_gs0, p.bf16_gs1, p.M, p.N, p.G0, p.G1, false);
  p.tma_epi = tma_ok ? 1 : 0;
  p.nbf = p.nbb = 0;
  p.nst = 2;
  p.bf16_wide = 0;
  // TA_GEMM_WIDE (dev): 0 = 32-column bf16 store boxes everywhere; 1 = wide boxes where the epilogue paces the launch: short K
  // (<= 512: q/k/v stage 2 forward 0.34 -> 0.30 ms) and activation epilogues (feed-forward 1 forward -4 %, the tanh GEMMs -5 %);
  // not the tanh' / ReLU' epilogues, whose staging buffers double as the landing zone of the saved activation's TMA loads;
  // 2 = wherever they apply (the plain long-K launches then lose an A/B ring stage to the larger buffers and gain nothing)
  static const int wide_mode = env_int("TA_GEMM_WIDE", 1);
  const bool wide_wanted = wide_mode == 2 || (wide_mode == 1 && (p.k_blocks <= 8 || p.act != TA_ACT_NONE));
  if (tma_ok && wide_wanted && p.out_bf16 && !need_f && !p.aux && p.N % 64 == 0 && chunks_per_warp % 2 == 0 &&
      make_epi_map(&em.bf16w, p.out_bf16, p.ld_bf16, p.bf16_gs0, p.bf16_gs1, p.M, p.N, p.G0, p.G1, false, true))
    p.bf16_wide = 1;
  if (tma_ok) {
    // short K, bf16 store only (q/k/v stage 2 forward): the tile's math is a few hundred cycles, the kernel runs at the
    // rate its TMA stores drain - four 2 KB chunk buffers per warp keep three stores in flight instead of one
    static const int force_nst = env_int("TA_GEMM_NST", 0);
    if (!need_f && need_b && !p.aux && ((force_nst == 0 && p.k_blocks <= 4) || force_nst == 4)) p.nst = 4;
    if (need_f) p.nbf = p.res ? 3 : 2;
    else if (need_b) p.nbb = p.aux ? 3 : p.nst;
    if (p.bf16_wide) p.nbb = p.nst == 4 ? 6 : 4;      // three / two 4 KB buffers per warp
    if (force_nl == 2 || force_nl == 3) { if (p.res) p.nbf = force_nl; if (p.aux) p.nbb = force_nl; }
    auto epi_bytes = [&]() { return ep_warps * (p.nbf * EPI_F32_BYTES + p.nbb * EPI_BF16_BYTES); };
    // long K: the epilogue has slack, the A/B ring does not - two buffers (no residual prefetch distance of 2) are enough
    if (!force_nl && p.nbf == 3 && p.k_blocks >= 32) p.nbf = 2;
    // keep the A/B ring at >= 4 stages when dropping the third (prefetch) buffer buys it
    if (!force_nl && p.nbf == 3 && (SMEM_LIMIT - epi_bytes()) / stage_bytes < 4 &&
        (SMEM_LIMIT - ep_warps * 2 * EPI_F32_BYTES) / stage_bytes >= 4 && p.k_blocks >= 16)
      p.nbf = 2;
    p.epi_warp_bytes = p.nbf * EPI_F32_BYTES + p.nbb * EPI_BF16_BYTES;
  } else {
    p.bf16_direct = 0;
    p.epi_warp_bytes = 0;
  }
  int stages = (SMEM_LIMIT - ep_warps * p.epi_warp_bytes) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.stages = stages;
}

template <int BLOCK_N>
static int launch_gemm(const ta_operand* A, const ta_operand* B, GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap tmA, tmB;
  if (make_operand_map(&tmA, A, p.M, p.K, p.G0, p.G1, BLOCK_M)) return 1;
  if (make_operand_map(&tmB, B, p.N, p.K, p.G0, p.G1, BLOCK_N)) return 1;
  p.n_blocks = (int)ceil_div(p.N, BLOCK_N);
  EpiMaps em;
  plan_epilogue(p, em, Cfg::STAGE_BYTES, Cfg::NW, (BLOCK_N / 32) / (Cfg::NW / 4));
  const int stages = p.stages;
  const int smem_bytes = stages * Cfg::STAGE_BYTES + Cfg::NW * p.epi_warp_bytes + 1024;
  static bool attr_set[TA_MAX_DEVICES] = {};
  TA_CHECK_CUDA(ensure_dynamic_smem(gemm_bf16_kernel<BLOCK_N>, SMEM_LIMIT + 1024, attr_set));
  const long long total = (long long)p.m_blocks * p.n_blocks * p.G0 * p.G1;
  const int grid = (int)(total < num_sms() ? total : num_sms());
  launch_kernel(gemm_bf16_kernel<BLOCK_N>, dim3(grid), dim3(Cfg::THREADS), smem_bytes, stream, tmA, tmB, em, p);
  TA_CHECK_LAUNCH();
  return 0;
}

// 2-CTA (CTA pair) launch: 256 x 256 tiles, clusters of 2 along x
static int launch_gemm2(const ta_operand* A, const ta_operand* B, GemmParams& p, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  if (make_operand_map(&tmA, A, p.M, p.K, p.G0, p.G1, BLOCK_M)) return 1;
  if (make_operand_map(&tmB, B, p.N, p.K, p.G0, p.G1, G2_B_HALF)) return 1;
  p.n_blocks = (int)ceil_div(p.N, G2_BLOCK_N);
  EpiMaps em;
  plan_epilogue(p, em, G2_STAGE_BYTES, 8, (G2_BLOCK_N / 32) / 2);
  const int smem_bytes = p.stages * G2_STAGE_BYTES + 8 * p.epi_warp_bytes + 1024;
  static bool attr_set[TA_MAX_DEVICES] = {};
  TA_CHECK_CUDA(ensure_dynamic_smem(gemm2_bf16_kernel, SMEM_LIMIT + 1024, attr_set));
  const long long m_pairs = ceil_div(p.M, 2 * BLOCK_M);
  const long long total = m_pairs * p.n_blocks * p.G0 * p.G1;
  const long long max_clusters = num_sms() / 2;
  const int clusters = (int)(total < max_clusters ? total : max_clusters);
  launch_kernel(gemm2_bf16_kernel, dim3(2 * clusters), dim3(GEMM2_THREADS), smem_bytes, stream, tmA, tmB, em, p);
  TA_CHECK_LAUNCH();
  return 0;
}

}  // namespace ta

extern "C" int32_t ta_gemm_bf16(const ta_operand* A, const ta_operand* B, int64_t M, int64_t N, int64_t K, int32_t G0,
                                int32_t G1, const ta_epilogue* ep, void* stream) {
  using namespace ta;
  TA_CHECK_ARG(A && B && ep, "ta_gemm_bf16: null argument");
  TA_CHECK_ARG(M > 0 && N > 0 && K > 0 && G0 > 0 && G1 > 0, "ta_gemm_bf16: non-positive shape");
  TA_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "ta_gemm_bf16: shape too large");
  TA_CHECK_ARG(ep->out_f32 || ep->out_bf16, "ta_gemm_bf16: no output");
  TA_CHECK_ARG(ep->dropout_p >= 0.f && ep->dropout_p < 1.f, "ta_gemm_bf16: dropout_p must be in [0, 1)");
  TA_CHECK_ARG(!ep->aux || ep->aux_mode != TA_AUX_NONE, "ta_gemm_bf16: aux given without aux_mode");
  GemmParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.G0 = G0; p.G1 = G1;
  p.a_mn = A->major ? 1 : 0; p.b_mn = B->major ? 1 : 0;
  p.m_blocks = (int)ceil_div(M, BLOCK_M);
  p.k_blocks = (int)ceil_div(K, BLOCK_K);
  p.alpha = ep->alpha; p.act = ep->act; p.aux_mode = ep->aux ? ep->aux_mode : 0;
  p.bias = ep->bias; p.bias_gs0 = ep->bias_gs0; p.bias_gs1 = ep->bias_gs1;
  p.aux = static_cast<const __nv_bfloat16*>(ep->aux); p.ld_aux = ep->ld_aux; p.aux_gs0 = ep->aux_gs0; p.aux_gs1 = ep->aux_gs1;
  p.res = ep->residual; p.ld_res = ep->ld_res; p.res_gs0 = ep->res_gs0; p.res_gs1 = ep->res_gs1;
  p.out_f32 = ep->out_f32; p.ld_f32 = ep->ld_f32; p.f32_gs0 = ep->f32_gs0; p.f32_gs1 = ep->f32_gs1;
  p.out_bf16 = static_cast<__nv_bfloat16*>(ep->out_bf16); p.ld_bf16 = ep->ld_bf16; p.bf16_gs0 = ep->bf16_gs0; p.bf16_gs1 = ep->bf16_gs1;
  auto al = [](const void* ptr, long long ld, long long s0, long long s1, int elems, int bytes) {
    return ptr && (reinterpret_cast<uintptr_t>(ptr) % bytes == 0) && ld % elems == 0 && s0 % elems == 0 && s1 % elems == 0;
  };
  p.vec_f32 = al(p.out_f32, p.ld_f32, p.f32_gs0, p.f32_gs1, 4, 16);
  p.vec_res = al(p.res, p.ld_res, p.res_gs0, p.res_gs1, 4, 16);
  p.vec_bf16 = al(p.out_bf16, p.ld_bf16, p.bf16_gs0, p.bf16_gs1, 8, 16);
  p.vec_aux = al(p.aux, p.ld_aux, p.aux_gs0, p.aux_gs1, 8, 16);
  p.vec_bias = al(p.bias, 4, p.bias_gs0, p.bias_gs1, 4, 16);
  { static const int dbg = env_int("TA_GEMM_DEBUG", 0); p.debug = dbg; }
  p.drop_thresh = ep->dropout_p > 0.f ? dropout_thresh16(ep->dropout_p) : 0u;
  p.drop_scale = ep->dropout_p > 0.f ? 1.f / (1.f - ep->dropout_p) : 1.f;
  p.drop_seed = ep->dropout_seed;
  p.drop_step = dropout_step_counter();
  {
    static const int force_mm = env_int("TA_GEMM_M_MAJOR", -1);
    p.m_major = force_mm >= 0 ? force_mm : ((long long)G0 * G1 >= 16 && p.k_blocks <= 4 && p.m_blocks > 1);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // tile width: the widest UMMA N that still yields about a wave of tiles
  const long long groups = (long long)G0 * G1;
  auto tiles = [&](int bn) { return groups * p.m_blocks * ceil_div(N, bn); };
  static const int force_bn = env_int("TA_GEMM_BN", 0);
  int bn = 256;
  if (force_bn == 64 || force_bn == 128 || force_bn == 256) bn = force_bn;
  else if (N <= 64) bn = 64;
  else if (N <= 128) bn = 128;
  else if (tiles(256) < num_sms() && tiles(128) > tiles(256)) bn = 128;
  if (bn == 128 && N > 64 && tiles(128) < num_sms() / 2 && tiles(64) > tiles(128)) bn = 64;
  // CTA pairs (256 x 256 tiles, B split across the pair) for the large problems: a third less L2 -> smem traffic
  static const int use_2cta = env_int("TA_GEMM_2CTA", 1);
  static const int min_k_2cta = env_int("TA_GEMM_2CTA_MINK", 128);
  const long long pair_tiles = groups * ceil_div(M, 2 * BLOCK_M) * ceil_div(N, 256);
  if (use_2cta && M >= 256 && N >= 256 && K >= min_k_2cta && pair_tiles >= num_sms() / 4) {
    if (bn == 256) return launch_gemm2(A, B, p, st);
    // 128-wide tiles were picked only to fill a wave (fewer than one 256-wide tile per SM): if the CTA pairs need no
    // more time - waves of pair tiles x 2 <= waves of 128 x 128 tiles - take them, they move a third less L2 -> smem
    // traffic per flop (the 1-CTA kernel ran the small dW problems, M, N ~ 1024, K = 16384, at 0.62 tensor-pipe activity)
    if (!force_bn && bn == 128 && 2 * ceil_div(pair_tiles, num_sms() / 2) <= ceil_div(tiles(128), num_sms()))
      return launch_gemm2(A, B, p, st);
  }
  if (bn == 256) return launch_gemm<256>(A, B, p, st);
  if (bn == 128) return launch_gemm<128>(A, B, p, st);
  return launch_gemm<64>(A, B, p, st);
}
