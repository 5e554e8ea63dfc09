// K3 forward, second generation — causal flash-style attention for the self-attention heads (/root/reference/model.py:60-73)
// and the selective cross-attention heads (model.py:136-159: one independent causal softmax per source, outputs summed).
//
// Why a second kernel: the first one (attention_fwd.cu) works on ONE 128 x 128 score tile at a time with 16 math warps,
// thread = (row, column quarter). The four warps of an SM sub-partition exchange row maxima through shared memory and a named
// barrier for every tile, so all of them load from TMEM at the same time, wait at the barrier at the same time and queue at the
// exponential unit at the same time: tensor pipe 15 %, XU 30 %, issue slots 53 % busy, 342 us for a c4 self-attention launch
// whose HBM time is 170 us (profiles/r01_prof_attn_fwd_v5_ncu_full.txt; ~430 warp instructions per tile and warp).
// Here:
//   * one thread owns one query ROW (128 score columns in registers): no exchange, no named barrier, the row sum and the
//     log-sum-exp are thread-local;
//   * TWO independent tile streams per CTA (4 math warps + loader + issuer + storer warp each, own barriers, own shared
//     memory, own half of tensor memory). A stream works on whole (work item, query tile) units; units are dealt to the
//     streams in the pattern 0 1 1 0 so both get the same number of tiles for any T. While one stream waits for its
//     S = Q K^T or reads tensor memory, the other one keeps the exponential unit busy - the phases interleave by themselves;
//   * P never touches shared memory: the un-normalised probabilities are packed to bf16 and written back to tensor memory
//     (tcgen05.st, 64 columns of their own), and O += P V takes its A operand from tensor memory (tcgen05.mma, "TS" form).
//     No st.shared, no proxy fence, no swizzle arithmetic, and 64 KB of shared memory freed for the second stream;
//   * S = Q K^T of the NEXT tile is issued as soon as every row has read the current S out of tensor memory (s_empty), so it
//     runs under the exponentials of the current tile (the first version of this file aliased P over S, which chained
//     S(n+1) behind P V(n): ncu showed the math warps waiting on mbarriers half of the time, 305 us per c4 launch);
//   * the sum over the cross-attention sources lives in 64 registers of the row's thread (tensor memory is full: per stream
//     S [0, 128), P [128, 192), O [192, 256)), and every math warp stores its own 32 output rows by TMA from its part of a
//     staging tile - no storer warp and no barrier between warps, so the CTA is 12 warps and a math thread may hold 240 registers.
// Per stream smem: Q 2 slots, K/V 2 stages, one staging tile. Nothing of size T x T touches HBM. head_size <= 64.
#include <cmath>
#include <cstdlib>
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "attention_fwd.cuh"
#include "common.cuh"
#include "philox.cuh"
#include "ptx.cuh"

namespace ta {

constexpr int A2_THREADS = 384;       // warps 0-3 / 4-7: math of stream 0 / 1; 8, 9 (10, 11): loader, issuer of stream 0 (1)
constexpr int A2_KV_STAGES = 2;
constexpr int A2_STREAM_TILES = 2 + 2 * A2_KV_STAGES + 1;                 // Q x2, (K, V) x stages, staging
constexpr int A2_STREAM_BYTES = A2_STREAM_TILES * AF_TILE;                // 114 688
constexpr int A2_BARS = 4 + 2 * A2_KV_STAGES + 6;                         // per stream
constexpr int A2_SMEM = 2 * A2_STREAM_BYTES + 2 * A2_BARS * 8 + 16;

// Walk of one stream over its units (work item, query tile i) and inside a unit over (source s, key/value block j <= i).
// Every role of the stream walks the same sequence with its own copy.
struct A2Iter {
  int work, i, s, j;
  int u;           // units of the CTA so far (both streams): decides the owner
  int n, k, qc;    // running counters of THIS stream: tiles, softmax segments, units
  int stream, n_qt, n_src, stride, total;
  __device__ __forceinline__ bool mine() const { return (((u & 3) == 1) || ((u & 3) == 2)) == (stream == 1); }
  __device__ __forceinline__ void next_unit() {
    ++u;
    if (i + 1 < n_qt) { ++i; return; }
    i = 0; work += stride;
  }
  __device__ __forceinline__ void init(int first, int stream_, int n_qt_, int n_src_, int stride_, int total_) {
    work = first; i = s = j = 0; u = 0; n = k = qc = 0;
    stream = stream_; n_qt = n_qt_; n_src = n_src_; stride = stride_; total = total_;
    while (work < total && !mine()) next_unit();
  }
  __device__ __forceinline__ bool done() const { return work >= total; }
  __device__ __forceinline__ bool first_in_seg() const { return j == 0; }
  __device__ __forceinline__ bool last_in_seg() const { return j == i; }
  __device__ __forceinline__ bool new_unit() const { return j == 0 && s == 0; }
  __device__ __forceinline__ bool last_in_unit() const { return j == i && s == n_src - 1; }
  __device__ __forceinline__ void next() {
    ++n;
    if (j < i) { ++j; return; }
    ++k; j = 0;
    if (s + 1 < n_src) { ++s; return; }
    s = 0; ++qc;
    do { next_unit(); } while (work < total && !mine());
  }
};

template <int DP, bool MULTI, bool DROPOUT>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnFwdParams p) {
  constexpr int NS = A2_KV_STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  // the shuffle tells the compiler that `warp` is warp-uniform (TMA / UMMA operands then stay in uniform registers)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int stream = warp < 8 ? (warp >> 2) : ((warp - 8) >> 1);
  const int role = warp < 8 ? 0 : 1 + ((warp - 8) & 1);      // 0 math, 1 loader, 2 issuer

  uint8_t* sbase = smem + stream * A2_STREAM_BYTES;
  const uint32_t aQ = smem_u32(sbase);                        // 2 slots
  const uint32_t aKV = aQ + 2 * AF_TILE;                      // NS x (K, V)
  const uint32_t aStage = aKV + NS * 2 * AF_TILE;             // one output tile staged for the TMA store
  uint64_t* bars0 = reinterpret_cast<uint64_t*>(smem + 2 * A2_STREAM_BYTES);
  const uint32_t bb = smem_u32(bars0) + stream * A2_BARS * 8;
  const uint32_t b_q_full = bb, b_q_empty = bb + 16, b_kv_full = bb + 32, b_kv_empty = bb + 32 + 8 * NS;
  const uint32_t b_s_full = bb + 32 + 16 * NS, b_s_empty = b_s_full + 8, b_p_full = b_s_full + 16, b_pv_done = b_s_full + 24,
                 b_o_full = b_s_full + 32, b_o_empty = b_s_full + 40;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars0 + 2 * A2_BARS);

  if (tid == 0) {
    for (int st = 0; st < 2; ++st) {
      uint64_t* b = bars0 + st * A2_BARS;
      for (int x = 0; x < 2; ++x) { mbar_init(&b[x], 1); mbar_init(&b[2 + x], 1); }
      for (int x = 0; x < NS; ++x) { mbar_init(&b[4 + x], 1); mbar_init(&b[4 + NS + x], 1); }
      uint64_t* c = b + 4 + 2 * NS;
      mbar_init(&c[0], 1);      // s_full   : S = Q K^T of the tile is in tensor memory
      mbar_init(&c[1], 128);    // s_empty  : every row has read it (the next tile's S may be issued)
      mbar_init(&c[2], 128);    // p_full   : every row has written P (and rescaled O if its maximum moved)
      mbar_init(&c[3], 1);      // pv_done  : O += P V of the tile complete (P may be overwritten; the rescale path)
      mbar_init(&c[4], 1);      // o_full   : last P V of the segment complete
      mbar_init(&c[5], 128);    // o_empty  : the segment epilogue has read O
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_base_smem + stream * 256;     // S [0, 128), P [128, 192), O [192, 256)
  const uint32_t tmem_p = tmem_s + 128, tmem_o = tmem_s + 192;
  pdl_launch_dependents();       // PDL: on-chip set-up done; wait for the previous kernel before touching global memory
  pdl_wait();

  A2Iter it;
  it.init(blockIdx.x, stream, p.n_qtiles, p.n_src, gridDim.x, p.total_work);

  if (role != 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");      // the CTA owns 384 x 168 = 64 512 registers: 128 x 24 + 256 x 240 (an inc beyond the pool blocks forever)
    if (role == 1) {
      // ================================================================== loader (whole warp, one elected lane issues)
      for (; !it.done(); it.next()) {
        int w = it.work;
        const int h = w % p.H; w /= p.H;
        const int b = w % p.B;
        const int g = w / p.B;
        if (it.new_unit()) {
          const int qs = it.qc & 1;
          mbar_wait_a(b_q_empty + 8 * qs, ((it.qc >> 1) & 1) ^ 1);
          mbar_expect_tx_ws(b_q_full + 8 * qs, AF_TILE);
          tma_load_5d_ws(aQ + qs * AF_TILE, &maps.q, b_q_full + 8 * qs, 0, h, it.i * 128, b, g);
        }
        const int st = it.n % NS;
        mbar_wait_a(b_kv_empty + 8 * st, ((it.n / NS) & 1) ^ 1);
        mbar_expect_tx_ws(b_kv_full + 8 * st, 2 * AF_TILE);
        const uint32_t dK = aKV + st * 2 * AF_TILE;
        tma_load_5d_ws(dK, &maps.k[it.s], b_kv_full + 8 * st, 0, h, it.j * 128, b, g);
        tma_load_5d_ws(dK + AF_TILE, &maps.v[it.s], b_kv_full + 8 * st, 0, h, it.j * 128, b, g);
      }
    } else if (role == 2) {
      // ================================================================== UMMA issuer (whole warp, one elected lane issues)
      // S(n+1) goes out as soon as every row has read S(n) (s_empty) - it runs under the exponentials of tile n; P V(n) goes
      // out when every row has written P(n) (p_full). P has columns of its own, so the two do not constrain each other.
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DP, 0, 1);
      const int nk_qk = (p.hs + 15) / 16;
      const uint64_t d_k = umma_desc_sw128(0, 16, 1024);        // K-major operand (Q, K), 128-byte rows
      const uint64_t d_v = umma_desc_sw128(0, 8192, 1024);      // V: MN-major B operand
      auto issue_s = [&](const A2Iter& a) {
        const int qs = a.qc & 1, st = a.n % NS;
        if (a.new_unit()) mbar_wait_a(b_q_full + 8 * qs, (a.qc >> 1) & 1);
        mbar_wait_a(b_kv_full + 8 * st, (a.n / NS) & 1);
        if (a.n > 0) mbar_wait_a(b_s_empty, (a.n - 1) & 1);
        tc_fence_after();
        const uint64_t dq = d_k + (((aQ + qs * AF_TILE) & 0x3FFFF) >> 4), dkk = d_k + (((aKV + st * 2 * AF_TILE) & 0x3FFFF) >> 4);
        for (int kk = 0; kk < nk_qk; ++kk) umma_bf16_ws(tmem_s, dq + kk * 2, dkk + kk * 2, idesc_s, kk != 0);
        umma_commit_ws(b_s_full);
      };
      auto issue_pv = [&](const A2Iter& a) {
        const int qs = a.qc & 1, st = a.n % NS;
        if (a.first_in_seg()) mbar_wait_a(b_o_empty, (a.k & 1) ^ 1);        // the previous segment's O has been read out
        mbar_wait_a(b_p_full, a.n & 1);
        tc_fence_after();
        const uint64_t dv = d_v + (((aKV + st * 2 * AF_TILE + AF_TILE) & 0x3FFFF) >> 4);
        // only the k-steps (16 key rows) that can hold unmasked columns for some row of the tile
        const int last_row = (a.i * 128 + 127 < p.T - 1) ? a.i * 128 + 127 : p.T - 1;
        int nk = (last_row - a.j * 128) / 16 + 1;
        if (nk > 8) nk = 8;
        const uint32_t accf = a.first_in_seg() ? 0u : 1u;
        for (int kk = 0; kk < nk; ++kk) umma_bf16_ts_ws(tmem_o, tmem_p + kk * 8, dv + kk * 128, idesc_o, accf | (kk != 0));
        umma_commit_ws(b_pv_done);
        umma_commit_ws(b_kv_empty + 8 * st);
        if (a.last_in_seg()) umma_commit_ws(b_o_full);
        if (a.last_in_unit()) umma_commit_ws(b_q_empty + 8 * qs);
      };
      A2Iter is = it;      // S runs one tile ahead of P V
      if (!is.done()) { issue_s(is); is.next(); }
      for (; !it.done(); it.next()) {
        if (!is.done()) { issue_s(is); is.next(); }
        issue_pv(it);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    // ================================================================== math warps: thread = query row
    const int wq = warp & 3, lane = tid & 31;
    const int r = wq * 32 + lane;                                  // row of the tile = TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t aStageW = aStage + wq * 4096;                    // this warp's 32 rows of the staging tile
    float m_run = -INFINITY, l_run = 0.f;
    float sum[MULTI ? DP : 1];                                     // running sum over the sources of this row's output
    // deferred segment epilogue: the last P V of a segment completes while the NEXT tile's scores are loaded and exponentiated,
    // so its result is picked up after that tile's P has been written - the math warps never sit waiting for the tensor pipe
    // (waiting for o_full right after the segment's last tile cost 19 % of their time)
    bool pend = false;
    int pe_i = 0, pe_s = 0, pe_k = 0, pe_h = 0, pe_b = 0, pe_g = 0;
    float pe_m = 0.f, pe_l = 0.f;

    // this thread's row of the staging tile (128B-swizzled like the TMA store box): 16 output columns = two 16-byte chunks
    auto stage16 = [&](int c, const float (&x)[16]) {
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint4 u;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int q = 0; q < 4; ++q) h2[q] = __floats2bfloat162_rn(x[8 * hlf + 2 * q], x[8 * hlf + 2 * q + 1]);
        sts128(aStage + r * 128 + (((((c >> 3) + hlf)) ^ (r & 7)) << 4), u);
      }
    };
    // Every warp stores its own 32 rows (the store maps have 32-row boxes): lane 0 issues the TMA store and knows when the
    // previous one has read the warp's part of the staging tile - no barrier between the warps of a stream (with 128-row
    // stores two named barriers per segment cost 9 % of the cross-attention kernel, mostly waiting for the warp with the
    // most unmasked columns)
    auto staging_acquire = [&]() {
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
    };
    auto staging_store = [&](const CUtensorMap* m0, const CUtensorMap* m1) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_5d(m0, aStageW, 0, pe_h, pe_i * 128 + wq * 32, pe_b, pe_g);
        if (m1) tma_store_5d(m1, aStageW, 0, pe_h, pe_i * 128 + wq * 32, pe_b, pe_g);
        bulk_commit_group();
      }
    };
    // O / l, += over the sources (registers), bf16 tile staged for the TMA store, log-sum-exp
    auto epilogue = [&]() {
      const float inv = 1.f / pe_l;
      const bool ev_a = MULTI ? p.o_src != nullptr : true;    // this segment's own normalised output is stored
      mbar_wait_a(b_o_full, pe_k & 1);
      tc_fence_after();
      if (ev_a) staging_acquire();
#pragma unroll
      for (int c = 0; c < DP; c += 16) {
        uint32_t ov[16];
        float val[16];
        tmem_ld_32x16(tmem_o + lane_off + c, ov);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) val[e] = __uint_as_float(ov[e]) * inv;
        if (ev_a) stage16(c, val);
        if constexpr (MULTI) {
#pragma unroll
          for (int e = 0; e < 16; ++e) sum[c + e] = pe_s > 0 ? sum[c + e] + val[e] : val[e];
        }
      }
      tc_fence_before();
      mbar_arrive_a(b_o_empty);
      if constexpr (MULTI) {
        if (ev_a) staging_store(&maps.o_src[pe_s], nullptr);
        if (pe_s == p.n_src - 1) {     // the sum over the sources
          staging_acquire();
#pragma unroll
          for (int c = 0; c < DP; c += 16) {
            float val[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) val[e] = sum[c + e];
            stage16(c, val);
          }
          staging_store(&maps.o, nullptr);
        }
      } else {
        staging_store(&maps.o, p.o_src ? &maps.o_src[0] : nullptr);
      }
      const int prow = pe_i * 128 + r;
      if (prow < p.T)
        p.lse[((((long long)pe_s * p.G + pe_g) * p.B + pe_b) * p.H + pe_h) * p.T + prow] = pe_m * 0.6931471805599453f + logf(pe_l);
      pend = false;
    };

    for (; !it.done(); it.next()) {
      const bool diag = it.j == it.i;
      if (it.first_in_seg()) { m_run = -INFINITY; l_run = 0.f; }
      // 32-column chunks this warp needs: all four below the diagonal tile, else those that reach its last row (warp-uniform)
      const int nch = diag ? wq + 1 : 4;
      mbar_wait_a(b_s_full, it.n & 1);
      tc_fence_after();
      uint32_t sv[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < nch) tmem_ld_32x32(tmem_s + lane_off + c * 32, sv[c]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive_a(b_s_empty);                                      // the next tile's S may be issued
      if (diag) {
        // the chunk the diagonal crosses (warp-uniform index wq): column e is unmasked for this row iff e <= lane; a masked
        // score becomes -inf, which the maximum ignores and ex2 turns into an exact zero - one code path for every chunk
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c == wq) {
#pragma unroll
            for (int e = 0; e < 32; ++e) sv[c][e] = e <= lane ? sv[c][e] : 0xff800000u;
          }
        }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nch) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              mx4[q] = fmaxf(mx4[q], fmaxf(__uint_as_float(sv[c][e + 2 * q]), __uint_as_float(sv[c][e + 2 * q + 1])));
          }
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;   // column 0 is never masked: finite
      float m_new = fmaxf(m_run, mx);
      if (m_run != -INFINITY && m_new - m_run <= 8.f) m_new = m_run;   // lazy rescale: P <= 2^8, exact after / l
      const float alpha = ex2_approx(m_run - m_new);                    // first block: exp2(-inf) = 0
      float ps4[4] = {0.f, 0.f, 0.f, 0.f};
      if (it.n > 0) {                   // P V of the previous tile has read P (and O may be rescaled below)
        mbar_wait_a(b_pv_done, (it.n - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
        if (c < nch) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float x0 = ex2_approx(fmaf(__uint_as_float(sv[c][e]), p.scale_log2, -m_new));
            const float x1 = ex2_approx(fmaf(__uint_as_float(sv[c][e + 1]), p.scale_log2, -m_new));
            ps4[(e >> 1) & 3] += x0 + x1;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h2);
          }
          if constexpr (DROPOUT) {
            // inverted dropout on the probabilities (the softmax denominator is taken BEFORE it, model.py:68-69): the keep
            // mask is a function of (seed, plane, query row, key column), regenerated by the backward kernel
            const unsigned long long dsd = p.drop_seed + (p.drop_step ? *p.drop_step * 0x9E3779B97F4A7C15ull : 0ull);
            const uint32_t plane = (uint32_t)it.s * (uint32_t)p.total_work + (uint32_t)it.work;   // ((s*G + g)*B + b)*H + h
            const __nv_bfloat162 sc2 = __floats2bfloat162_rn(p.drop_scale, p.drop_scale);
#pragma unroll 1
            for (int i8 = 0; i8 < 4; ++i8) {
              const uint32_t mk = dropout_keep8(dsd, 1u, plane, (uint32_t)(it.i * 128 + r), (uint32_t)((it.j * 128 + c * 32) >> 3) + i8,
                                                p.drop_thresh);
#pragma unroll
              for (int q = 0; q < 4; ++q) {          // static register indices: select the four words of group i8
                uint32_t w4 = i8 == 0 ? pk[q] : i8 == 1 ? pk[4 + q] : i8 == 2 ? pk[8 + q] : pk[12 + q];
                __nv_bfloat162 v2 = *reinterpret_cast<__nv_bfloat162*>(&w4);
                v2 = __hmul2(v2, sc2);
                uint32_t u = *reinterpret_cast<uint32_t*>(&v2);
                u &= ((mk >> (2 * q)) & 1u ? 0x0000FFFFu : 0u) | ((mk >> (2 * q + 1)) & 1u ? 0xFFFF0000u : 0u);
                if (i8 == 0) pk[q] = u; else if (i8 == 1) pk[4 + q] = u; else if (i8 == 2) pk[8 + q] = u; else pk[12 + q] = u;
              }
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = 0u;
        }
        // P as the A operand of P V: row = lane, key columns 2w / 2w+1 in the halves of 32-bit column w
        tmem_st_cols<16>(tmem_p + lane_off + c * 16, pk);
      }
      l_run = l_run * alpha + ((ps4[0] + ps4[1]) + (ps4[2] + ps4[3]));
      if (!it.first_in_seg() && __any_sync(0xffffffffu, alpha != 1.f)) {
        // the running maximum of some row moved by more than 2^8: rescale O (P V of the previous tile is complete: pv_done)
#pragma unroll
        for (int c = 0; c < DP; c += 16) {
          uint32_t ov[16];
          tmem_ld_32x16(tmem_o + lane_off + c, ov);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
          tmem_st_cols<16>(tmem_o + lane_off + c, ov);
        }
      }
      m_run = m_new;
      if (pend) epilogue();            // the previous segment (its O must be read before this tile's P V may overwrite it)
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive_a(b_p_full);
      if (it.last_in_seg()) {
        pend = true;
        int w = it.work;
        pe_h = w % p.H; w /= p.H;
        pe_b = w % p.B;
        pe_g = w / p.B;
        pe_i = it.i; pe_s = it.s; pe_k = it.k;
        pe_m = m_run; pe_l = l_run;
      }
    }
    if (pend) epilogue();
    if (lane == 0) bulk_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(*tmem_base_smem, 512);
}

template <int DP, bool MULTI, bool DROPOUT>
static int launch_fwd2_d(const AttnMaps& maps, const AttnFwdParams& p, cudaStream_t st) {
  static bool attr[TA_MAX_DEVICES] = {};
  TA_CHECK_CUDA(ensure_dynamic_smem(attn_fwd2_kernel<DP, MULTI, DROPOUT>, A2_SMEM, attr));
  const int grid = p.total_work < num_sms() ? p.total_work : num_sms();
  launch_kernel(attn_fwd2_kernel<DP, MULTI, DROPOUT>, dim3(grid), dim3(A2_THREADS), A2_SMEM, st, maps, p);
  TA_CHECK_LAUNCH();
  return 0;
}

// dropout is a template parameter: its Philox code inside the unrolled tile loop made the hot path miss the instruction
// cache (no_inst 11-22 % of the stall samples) although it never ran
template <int DP, bool MULTI>
static int launch_fwd2(const AttnMaps& maps, const AttnFwdParams& p, cudaStream_t st) {
  return p.drop_thresh ? launch_fwd2_d<DP, MULTI, true>(maps, p, st) : launch_fwd2_d<DP, MULTI, false>(maps, p, st);
}

int launch_attn_fwd2(const AttnMaps& maps, const AttnFwdParams& p, cudaStream_t st) {
  if (p.n_src > 1) {
    if (p.hs <= 16) return launch_fwd2<16, true>(maps, p, st);
    if (p.hs <= 32) return launch_fwd2<32, true>(maps, p, st);
    return launch_fwd2<64, true>(maps, p, st);
  }
  if (p.hs <= 16) return launch_fwd2<16, false>(maps, p, st);
  if (p.hs <= 32) return launch_fwd2<32, false>(maps, p, st);
  return launch_fwd2<64, false>(maps, p, st);
}

}  // namespace ta
