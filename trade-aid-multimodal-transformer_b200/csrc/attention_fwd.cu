// K3 forward — causal flash-style attention on tcgen05/TMEM for both the self-attention heads
// (/root/reference/model.py:60-73) and the selective cross-attention heads (model.py:136-159, one independent causal
// softmax per key/value source modality, outputs summed).
//
// Persistent, warp-specialised: one CTA per SM walks (group, batch, head) work items; inside an item the schedule is
// q tile i (128 rows) -> source s -> key/value block j <= i (128 rows): one (i, s) pass is a softmax "segment".
//   warp 16  loader : TMA Q tiles (2 slots) and K/V blocks (3-stage ring); runs ahead across segments and work items
//   warp 17  issuer : S = Q K^T (128x128x hs) into one of two TMEM S buffers, one pair AHEAD of the softmax;
//                     O += P V (128 x hs x 128, V as the MN-major B operand) into one of two TMEM O buffers
//   warp 18  storer : TMA stores of the output tiles the math warps stage in smem
//   warps 0-15 math : thread (row r, column quarter) owns 32 S columns: tcgen05.ld, causal mask, online softmax with exp2
//                     (row max exchanged between the four quarters through smem, lazy rescale of O in TMEM only when the
//                     max moves by more than 2^8), P (bf16) into one of two 128B-swizzled smem buffers; the segment
//                     epilogue (O / l, += over sources, lse) is deferred behind the next pair's exponentials so the
//                     last P V of a segment never stalls the math warps.
// The first version ran these phases one after another in a 128-thread CTA (tensor pipe 8% active, warps mostly
// waiting on mbarriers: profiles/r01_ncu_attn_v2_*.csv). Nothing of size T x T ever touches HBM. head_size <= 64.
#include <cmath>
#include <cstdlib>
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "attention_fwd.cuh"
#include "common.cuh"
#include "philox.cuh"
#include "ptx.cuh"

namespace ta {

constexpr int AF_MATH_WARPS = 16;
constexpr int AF_THREADS = (AF_MATH_WARPS + 4) * 32;   // warps 0-15 math, 16 loader, 17 issuer, 18-19 idle (whole warpgroup for setmaxnreg)
constexpr int AF_KV_STAGES = 3;

// The (work, i, s, j) schedule; every warp role walks the same sequence with its own copy.
struct AfIter {
  int work, i, s, j;
  int n, k, qc;   // running counters over the CTA's lifetime: pairs, segments, q tiles
  __device__ __forceinline__ void init(int first) { work = first; i = s = j = 0; n = k = qc = 0; }
  __device__ __forceinline__ bool first_in_seg() const { return j == 0; }
  __device__ __forceinline__ bool last_in_seg() const { return j == i; }
  __device__ __forceinline__ bool new_q() const { return j == 0 && s == 0; }
  __device__ __forceinline__ void next(int n_qt, int n_src, int stride) {
    ++n;
    if (j < i) { ++j; return; }
    ++k; j = 0;
    if (s + 1 < n_src) { ++s; return; }
    s = 0; ++qc;
    if (i + 1 < n_qt) { ++i; return; }
    i = 0; work += stride;
  }
};

template <int DP, bool MULTI>
__global__ void __launch_bounds__(AF_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnFwdParams p) {
  constexpr int EC = DP / 4 < 8 ? 8 : DP / 4;   // output columns per (active) math thread
  constexpr int NS = AF_KV_STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                              // 2 slots
  uint8_t* sKV = smem + 2 * AF_TILE;               // NS x (K, V)
  uint8_t* sP = sKV + NS * 2 * AF_TILE;            // 2 buffers x 2 column halves
  uint8_t* sStage = sP + 4 * AF_TILE;              // one 128 x 128 B output tile staged for the TMA store
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + AF_TILE);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* kv_full = bars + 4;       // [NS]
  uint64_t* kv_empty = bars + 4 + NS; // [NS]
  uint64_t* s_full = bars + 4 + 2 * NS;   // [2]
  uint64_t* s_empty = s_full + 2;     // [2] 256 arrivals
  uint64_t* p_full = s_full + 4;      // [2] 256 arrivals
  uint64_t* p_empty = s_full + 6;     // [2] P V of the pair complete (P buffer free, O updated)
  uint64_t* o_full = s_full + 8;      // [2]
  uint64_t* o_empty = s_full + 10;    // [2] 256 arrivals
  uint64_t* stg_full = s_full + 12;   // 512 arrivals: the staging tile holds an output tile
  uint64_t* stg_free = s_full + 13;   // the TMA store has read it
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(s_full + 14);
  float* sMax = reinterpret_cast<float*>(s_full + 15);   // [2 parity][4 quarters][128]
  float* sL = sMax + 1024;                                // [2 parity][4 quarters][128]
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  // the shuffle tells the compiler that `warp` is warp-uniform (TMA / UMMA operands then stay in uniform registers)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    for (int x = 0; x < 2; ++x) {
      mbar_init(&q_full[x], 1); mbar_init(&q_empty[x], 1);
      mbar_init(&s_full[x], 1); mbar_init(&s_empty[x], AF_MATH_WARPS * 32);
      mbar_init(&p_full[x], AF_MATH_WARPS * 32); mbar_init(&p_empty[x], 1);
      mbar_init(&o_full[x], 1); mbar_init(&o_empty[x], AF_MATH_WARPS * 32);
    }
    for (int x = 0; x < NS; ++x) { mbar_init(&kv_full[x], 1); mbar_init(&kv_empty[x], 1); }
    mbar_init(stg_full, AF_MATH_WARPS * 32); mbar_init(stg_free, 1);
    fence_mbar_init();
  }
  if (warp == AF_MATH_WARPS + 1) {
    tmem_alloc(tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_launch_dependents();       // PDL: on-chip set-up done; wait for the previous kernel before touching global memory
  pdl_wait();
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,256+DP) O1 [320,320+DP) sum over sources [384,384+DP)
  AfIter it;
  it.init(blockIdx.x);
  const int n_qt = p.n_qtiles, n_src = p.n_src, stride = gridDim.x, total = p.total_work;

  if (warp >= AF_MATH_WARPS) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == AF_MATH_WARPS) {
    // ================================================================== loader (whole warp, one elected lane issues)
    {
      const uint32_t aQ = smem_u32(sQ), aKV = smem_u32(sKV);
      const uint32_t bq_full = smem_u32(q_full), bq_empty = smem_u32(q_empty), bkv_full = smem_u32(kv_full), bkv_empty = smem_u32(kv_empty);
      for (; it.work < total; it.next(n_qt, n_src, stride)) {
        int w = it.work;
        const int h = w % p.H; w /= p.H;
        const int b = w % p.B;
        const int g = w / p.B;
        if (it.new_q()) {
          const int qs = it.qc & 1;
          mbar_wait_a(bq_empty + 8 * qs, ((it.qc >> 1) & 1) ^ 1);
          mbar_expect_tx_ws(bq_full + 8 * qs, AF_TILE);
          tma_load_5d_ws(aQ + qs * AF_TILE, &maps.q, bq_full + 8 * qs, 0, h, it.i * 128, b, g);
        }
        const int st = it.n % NS, u = it.n / NS;
        mbar_wait_a(bkv_empty + 8 * st, (u & 1) ^ 1);
        mbar_expect_tx_ws(bkv_full + 8 * st, 2 * AF_TILE);
        const uint32_t dK = aKV + st * 2 * AF_TILE;
        tma_load_5d_ws(dK, &maps.k[it.s], bkv_full + 8 * st, 0, h, it.j * 128, b, g);
        tma_load_5d_ws(dK + AF_TILE, &maps.v[it.s], bkv_full + 8 * st, 0, h, it.j * 128, b, g);
      }
    }
  } else if (warp == AF_MATH_WARPS + 1) {
    // ================================================================== UMMA issuer (whole warp, one elected lane issues)
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DP, 0, 1);
      const int nk_qk = (p.hs + 15) / 16;
      const uint32_t aQ0 = smem_u32(sQ), aKV = smem_u32(sKV), aP0 = smem_u32(sP);
      const uint32_t bq_full = smem_u32(q_full), bq_empty = smem_u32(q_empty), bkv_full = smem_u32(kv_full), bkv_empty = smem_u32(kv_empty);
      const uint32_t bs_full = smem_u32(s_full), bs_empty = smem_u32(s_empty), bp_full = smem_u32(p_full), bp_empty = smem_u32(p_empty);
      const uint32_t bo_full = smem_u32(o_full), bo_empty = smem_u32(o_empty);
      // descriptors step by adding (bytes >> 4) to the low word
      const uint64_t d_k = umma_desc_sw128(0, 16, 1024);        // K-major operand (Q, K, P), 128-byte rows
      const uint64_t d_v = umma_desc_sw128(0, 8192, 1024);      // V: MN-major B operand
      auto issue_s = [&](const AfIter& a) {
        const int qs = a.qc & 1, st = a.n % NS, sb = a.n & 1;
        if (a.new_q()) mbar_wait_a(bq_full + 8 * qs, (a.qc >> 1) & 1);
        mbar_wait_a(bkv_full + 8 * st, (a.n / NS) & 1);
        mbar_wait_a(bs_empty + 8 * sb, ((a.n >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t dq = d_k + (((aQ0 + qs * AF_TILE) & 0x3FFFF) >> 4), dkk = d_k + (((aKV + st * 2 * AF_TILE) & 0x3FFFF) >> 4);
        for (int kk = 0; kk < nk_qk; ++kk) umma_bf16_ws(tmem_base + sb * 128, dq + kk * 2, dkk + kk * 2, idesc_s, kk != 0);
        umma_commit_ws(bs_full + 8 * sb);
      };
      auto issue_pv = [&](const AfIter& a) {
        const int st = a.n % NS, pb = a.n & 1, ob = a.k & 1;
        if (a.first_in_seg()) mbar_wait_a(bo_empty + 8 * ob, ((a.k >> 1) & 1) ^ 1);
        mbar_wait_a(bp_full + 8 * pb, (a.n >> 1) & 1);
        tc_fence_after();
        const uint64_t dp = d_k + (((aP0 + pb * 2 * AF_TILE) & 0x3FFFF) >> 4);
        const uint64_t dv = d_v + (((aKV + st * 2 * AF_TILE + AF_TILE) & 0x3FFFF) >> 4);
        // only the k-steps that can hold unmasked columns for some row of the tile
        const int last_row = (a.i * 128 + 127 < p.T - 1) ? a.i * 128 + 127 : p.T - 1;
        int nk = (last_row - a.j * 128) / 16 + 1;
        if (nk > 8) nk = 8;
        const uint32_t accf = a.first_in_seg() ? 0u : 1u;
        for (int kk = 0; kk < nk; ++kk)
          umma_bf16_ws(tmem_base + 256 + ob * 64, dp + (kk >> 2) * (AF_TILE >> 4) + (kk & 3) * 2, dv + kk * 128, idesc_o,
                       accf | (kk != 0));
        umma_commit_ws(bp_empty + 8 * pb);
        umma_commit_ws(bkv_empty + 8 * st);
        if (a.last_in_seg()) umma_commit_ws(bo_full + 8 * ob);
        if ((a.j == a.i && a.s == n_src - 1)) umma_commit_ws(bq_empty + 8 * (a.qc & 1));
      };
      AfIter is = it;   // S runs one pair ahead of P V
      if (is.work < total) { issue_s(is); is.next(n_qt, n_src, stride); }
      for (; it.work < total; it.next(n_qt, n_src, stride)) {
        if (is.work < total) { issue_s(is); is.next(n_qt, n_src, stride); }
        issue_pv(it);
      }
    }
  } else if (warp == AF_MATH_WARPS + 2) {
    // ================================================================== storer (one lane): TMA stores of the output tiles
    if ((tid & 31) == 0) {
      const uint32_t aStage = smem_u32(sStage), bfull = smem_u32(stg_full), bfree = smem_u32(stg_free);
      int ev = 0;
      auto store_event = [&](const CUtensorMap* m0, const CUtensorMap* m1, int h, int i, int b, int g) {
        if (p.debug & 32) return;
        mbar_wait_a(bfull, ev & 1);
        tma_store_5d(m0, aStage, 0, h, i * 128, b, g);
        if (m1) tma_store_5d(m1, aStage, 0, h, i * 128, b, g);
        bulk_commit_group();
        bulk_wait_read<0>();
        mbar_arrive_a(bfree);
        ++ev;
      };
      for (; it.work < total; it.next(n_qt, n_src, stride)) {
        if (!it.last_in_seg()) continue;
        int w = it.work;
        const int h = w % p.H; w /= p.H;
        const int b = w % p.B;
        const int g = w / p.B;
        if constexpr (MULTI) {
          if (p.o_src) store_event(&maps.o_src[it.s], nullptr, h, it.i, b, g);     // this source's normalised output
          if (it.s == p.n_src - 1) store_event(&maps.o, nullptr, h, it.i, b, g);   // sum over the sources
        } else {
          store_event(&maps.o, p.o_src ? &maps.o_src[0] : nullptr, h, it.i, b, g);
        }
      }
      bulk_wait_all<0>();
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ================================================================== math warps
    // thread = (row r, column quarter qd): 32 S columns and EC output columns; the four warps that share a TMEM lane
    // quarter (warp % 4) exchange row maxima / row sums through smem and a 128-thread named barrier
    const int r = tid & 127, qd = tid >> 7, wq = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const int bar_id = 1 + wq;
    const bool ep_active = qd * EC < DP;        // DP = 16: only two quarters own output columns
    float m_run = -INFINITY, l_run = 0.f;
    // deferred segment epilogue
    bool pend = false;
    int pe_work = 0, pe_i = 0, pe_s = 0, pe_k = 0;
    bool pe_last_src = false;
    float pe_m = 0.f, pe_l = 0.f;
    int xpar = 0;     // parity of the max-exchange buffer
    int sev = 0;      // staging events so far (parity of stg_free)
    const uint32_t aStage = smem_u32(sStage), bstg_full = smem_u32(stg_full), bstg_free = smem_u32(stg_free);
    // Output path: 8 fp32 columns at a time (small temporaries: any spill goes to L2 here, the L1 is all shared
    // memory) TMEM -> normalise -> bf16 -> this thread's row of the staging tile (128B-swizzled like the TMA store
    // box); the storer warp issues the store. Rows beyond T and columns beyond hs are clipped by the tensor map.
    auto stage8 = [&](int c, const float (&x)[8]) {
      uint4 u;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) h2[q] = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
      sts128(aStage + r * 128 + ((((qd * EC + c) >> 3) ^ (r & 7)) << 4), u);
    };
    auto epilogue = [&]() {
      float* sl = sL + (pe_k & 1) * 512;
      sl[qd * 128 + r] = pe_l;
      named_bar_sync(bar_id, 128);
      const float l = (sl[r] + sl[128 + r]) + (sl[256 + r] + sl[384 + r]);
      const float inv = 1.f / l;
      const int ob = pe_k & 1;
      const bool skip = p.debug & 32;
      // event A: this segment's normalised output (the only output for a single source; o_src[s] otherwise)
      const bool ev_a = (MULTI ? p.o_src != nullptr : true) && !skip;
      mbar_wait(&o_full[ob], (pe_k >> 1) & 1);
      tc_fence_after();
      if (ev_a) mbar_wait_a(bstg_free, (sev & 1) ^ 1);    // the previous tile has been read out of the staging buffer
      if (ep_active) {
        const uint32_t to = tmem_base + 256 + ob * 64 + lane_off + qd * EC;
        const uint32_t tt = tmem_base + 384 + lane_off + qd * EC;   // running sum over the sources (MULTI)
#pragma unroll
        for (int c = 0; c < EC; c += 8) {
          uint32_t ov[8];
          float val[8];
          tmem_ld_32x8(to + c, ov);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) val[e] = __uint_as_float(ov[e]) * inv;
          if (ev_a) stage8(c, val);
          if constexpr (MULTI) {
            if (pe_s > 0) {
              tmem_ld_32x8(tt + c, ov);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 8; ++e) val[e] += __uint_as_float(ov[e]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) ov[e] = __float_as_uint(val[e]);
            tmem_st_cols<8>(tt + c, ov);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&o_empty[ob]);
      if (ev_a) {
        fence_proxy_async_smem();
        mbar_arrive_a(bstg_full);
        ++sev;
      }
      if constexpr (MULTI) {
        if (pe_last_src && !skip) {     // event B: the sum over the sources
          mbar_wait_a(bstg_free, (sev & 1) ^ 1);
          if (ep_active) {
            tmem_wait_st();
            const uint32_t tt = tmem_base + 384 + lane_off + qd * EC;
#pragma unroll
            for (int c = 0; c < EC; c += 8) {
              uint32_t ov[8];
              float val[8];
              tmem_ld_32x8(tt + c, ov);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 8; ++e) val[e] = __uint_as_float(ov[e]);
              stage8(c, val);
            }
          }
          fence_proxy_async_smem();
          mbar_arrive_a(bstg_full);
          ++sev;
        }
      }
      const int row = pe_i * 128 + r;
      if (qd == 0 && row < p.T) {
        int w = pe_work;
        const int h = w % p.H; w /= p.H;
        const int b = w % p.B;
        const int g = w / p.B;
        p.lse[((((long long)pe_s * p.G + g) * p.B + b) * p.H + h) * p.T + row] = pe_m * 0.6931471805599453f + logf(l);
      }
      pend = false;
    };

    for (; it.work < total; it.next(n_qt, n_src, stride)) {
      const int sb = it.n & 1, pb = sb;
      const int row = it.i * 128 + r, kv0 = it.j * 128;
      const bool diag = it.j == it.i;
      if (it.first_in_seg()) { m_run = -INFINITY; l_run = 0.f; }
      mbar_wait(&s_full[sb], (it.n >> 1) & 1);
      tc_fence_after();
      // this thread's 32 columns: needed unless they lie above the diagonal for every row of the warp (warp-uniform)
      const int col_base = kv0 + qd * 32;
      const bool need = !diag || col_base <= it.i * 128 + wq * 32 + 31;
      // columns of this chunk that are unmasked for this row: all 32 below the diagonal tile, else a prefix
      const int nvalid = !diag ? 32 : (row - col_base + 1 < 0 ? 0 : (row - col_base + 1 > 32 ? 32 : row - col_base + 1));
      uint32_t sv[32];
      if (need) {
        tmem_ld_32x32(tmem_base + sb * 128 + lane_off + qd * 32, sv);
        tmem_wait_ld();
      }
      tc_fence_before();
      mbar_arrive(&s_empty[sb]);             // the issuer may overwrite this S buffer (pair n + 2)
      float mx = -INFINITY;
      if (need) {
        if (nvalid == 32) {
#pragma unroll
          for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(sv[e]));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) mx = e < nvalid ? fmaxf(mx, __uint_as_float(sv[e])) : mx;
        }
      }
      float* sm = sMax + xpar * 512;
      xpar ^= 1;
      sm[qd * 128 + r] = mx;
      named_bar_sync(bar_id, 128);
      mx = fmaxf(fmaxf(sm[r], sm[128 + r]), fmaxf(sm[256 + r], sm[384 + r])) * p.scale_log2;  // column 0 is never masked
      float m_new = fmaxf(m_run, mx);
      if (m_run != -INFINITY && m_new - m_run <= 8.f) m_new = m_run;   // lazy rescale: P <= 2^8, exact after / l
      const float alpha = ex2_approx(m_run - m_new);                    // first block: exp2(-inf) = 0
      float psum = 0.f;
      uint32_t pk[16];
      if (need) {
        if (nvalid == 32) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float x0 = ex2_approx(__uint_as_float(sv[e]) * p.scale_log2 - m_new);
            const float x1 = ex2_approx(__uint_as_float(sv[e + 1]) * p.scale_log2 - m_new);
            psum += x0 + x1;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
            pk[e / 2] = *reinterpret_cast<uint32_t*>(&h2);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float x0 = ex2_approx(__uint_as_float(sv[e]) * p.scale_log2 - m_new);
            float x1 = ex2_approx(__uint_as_float(sv[e + 1]) * p.scale_log2 - m_new);
            x0 = e < nvalid ? x0 : 0.f;
            x1 = e + 1 < nvalid ? x1 : 0.f;
            psum += x0 + x1;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
            pk[e / 2] = *reinterpret_cast<uint32_t*>(&h2);
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = 0u;
      }
      l_run = l_run * alpha + psum;      // the softmax denominator is taken BEFORE dropout (model.py:68-69)
      if (p.drop_thresh && need) {
        // inverted dropout on the probabilities: the keep mask is a function of (seed, source / group / batch / head
        // plane, query row, key column), regenerated by the backward kernel
        const unsigned long long sd = p.drop_seed + (p.drop_step ? *p.drop_step * 0x9E3779B97F4A7C15ull : 0ull);
        int w = it.work;
        const uint32_t plane = (uint32_t)it.s * (uint32_t)total + (uint32_t)w;   // ((s*G + g)*B + b)*H + h
        const __nv_bfloat162 sc2 = __floats2bfloat162_rn(p.drop_scale, p.drop_scale);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t m = dropout_keep8(sd, 1u, plane, (uint32_t)row, (uint32_t)(col_base >> 3) + i, p.drop_thresh);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 v2 = *reinterpret_cast<__nv_bfloat162*>(&pk[4 * i + e]);
            v2 = __hmul2(v2, sc2);
            uint32_t u = *reinterpret_cast<uint32_t*>(&v2);
            u &= ((m >> (2 * e)) & 1u ? 0x0000FFFFu : 0u) | ((m >> (2 * e + 1)) & 1u ? 0xFFFF0000u : 0u);
            pk[4 * i + e] = u;
          }
        }
      }
      if (pend) epilogue();                  // previous segment: its last P V has had the exponentials above to finish
      // P -> smem (A operand of P V); the P V that last read this buffer (pair n - 2) must be complete
      mbar_wait(&p_empty[pb], ((it.n >> 1) & 1) ^ 1);
      uint8_t* rowp = sP + pb * 2 * AF_TILE + (qd >> 1) * AF_TILE + r * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(rowp + ((((qd & 1) * 4 + q) ^ (r & 7)) << 4)) =
            make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      fence_proxy_async_smem();
      if (!it.first_in_seg() && __any_sync(0xffffffffu, alpha != 1.f)) {
        // the running max of some row moved: rescale this thread's O columns in TMEM once P V of pair n - 1 is done
        mbar_wait(&p_empty[pb ^ 1], ((it.n - 1) >> 1) & 1);
        tc_fence_after();
        if (ep_active) {
          uint32_t ov[EC];
          const uint32_t to = tmem_base + 256 + (it.k & 1) * 64 + lane_off + qd * EC;
          tmem_ld_cols<EC>(to, ov);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < EC; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
          tmem_st_cols<EC>(to, ov);
          tmem_wait_st();
        }
      }
      m_run = m_new;
      tc_fence_before();
      mbar_arrive(&p_full[pb]);
      if (it.last_in_seg()) {
        pend = true;
        pe_work = it.work; pe_i = it.i; pe_s = it.s; pe_k = it.k; pe_last_src = it.s == p.n_src - 1;
        pe_m = m_run; pe_l = l_run;
      }
    }
    if (pend) epilogue();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AF_MATH_WARPS + 1) tmem_dealloc(tmem_base, 512);
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

template <int DP, bool MULTI>
static int launch_attn_fwd(const AttnMaps& maps, const AttnFwdParams& p, cudaStream_t st) {
  constexpr int SMEM = (2 + 2 * AF_KV_STAGES + 4 + 1) * AF_TILE + 256 + 8192 + 64;
  static bool attr[TA_MAX_DEVICES] = {};
  TA_CHECK_CUDA(ensure_dynamic_smem(attn_fwd_kernel<DP, MULTI>, SMEM, attr));
  const int grid = p.total_work < num_sms() ? p.total_work : num_sms();
  launch_kernel(attn_fwd_kernel<DP, MULTI>, dim3(grid), dim3(AF_THREADS), SMEM, st, maps, p);
  TA_CHECK_LAUNCH();
  return 0;
}

}  // namespace ta

using namespace ta;

extern "C" int32_t ta_attention_fwd(const ta_attn_tensor* q, const ta_attn_tensor* k, const ta_attn_tensor* v,
                                    int32_t n_src, void* o_bf16, int64_t ldo, int64_t o_gs, void* o_src_bf16,
                                    int64_t o_src_stride, float* lse, int32_t G, int32_t B, int32_t T, int32_t H,
                                    int32_t hs, float dropout_p, uint64_t seed, void* stream) {
  TA_CHECK_ARG(q && k && v && o_bf16 && lse, "ta_attention_fwd: null pointer");
  TA_CHECK_ARG(n_src >= 1 && n_src <= ATT_MAX_SRC, "ta_attention_fwd: n_src must be in [1, %d]", ATT_MAX_SRC);
  TA_CHECK_ARG(hs >= 8 && hs <= 64 && hs % 8 == 0, "ta_attention_fwd: head_size %d not supported (8..64, multiple of 8)", hs);
  TA_CHECK_ARG(G > 0 && B > 0 && T > 0 && H > 0, "ta_attention_fwd: non-positive size");
  TA_CHECK_ARG(ldo % 8 == 0 && o_gs % 8 == 0 && (reinterpret_cast<uintptr_t>(o_bf16) & 15) == 0, "ta_attention_fwd: output alignment");
  TA_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "ta_attention_fwd: dropout_p must be in [0, 1)");
  // TA_ATTN_FWD=1 keeps the first-generation kernel of this file (one tile at a time on 16 math warps, 128-row output
  // stores); default: the two-stream kernel of attention_fwd2.cu, whose math warps store 32 rows each
  static const int version = [] { const char* e = getenv("TA_ATTN_FWD"); return e ? atoi(e) : 2; }();
  const int store_rows = version == 1 ? 128 : 32;
  AttnMaps maps;
  if (make_head_map(&maps.q, q, hs, H, T, B, G, 128)) return 1;
  for (int s = 0; s < n_src; ++s) {
    if (make_head_map(&maps.k[s], &k[s], hs, H, T, B, G, 128)) return 1;
    if (make_head_map(&maps.v[s], &v[s], hs, H, T, B, G, 128)) return 1;
  }
  {
    ta_attn_tensor ot{o_bf16, ldo, hs, o_gs};
    if (make_head_map(&maps.o, &ot, hs, H, T, B, G, store_rows)) return 1;
    if (o_src_bf16)
      for (int s = 0; s < n_src; ++s) {
        ta_attn_tensor st{static_cast<const __nv_bfloat16*>(o_src_bf16) + (long long)s * o_src_stride, ldo, hs, o_gs};
        if (make_head_map(&maps.o_src[s], &st, hs, H, T, B, G, store_rows)) return 1;
      }
  }
  AttnFwdParams p{};
  p.T = T; p.B = B; p.H = H; p.G = G; p.hs = hs; p.n_src = n_src;
  p.n_qtiles = (T + 127) / 128;
  p.scale_log2 = (float)(1.4426950408889634 / sqrt((double)hs));
  { const char* e = getenv("TA_ATTN_DEBUG"); p.debug = e ? atoi(e) : 0; }
  p.drop_thresh = dropout_p > 0.f ? dropout_thresh16(dropout_p) : 0u;
  p.drop_scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  p.drop_seed = seed;
  p.drop_step = dropout_step_counter();
  p.o = (__nv_bfloat16*)o_bf16; p.ldo = ldo; p.o_gs = o_gs;
  p.o_src = (__nv_bfloat16*)o_src_bf16; p.o_src_stride = o_src_stride;
  p.lse = lse;
  const long long total = (long long)G * B * H;
  TA_CHECK_ARG(total < (1ll << 31), "attention grid too large");
  p.total_work = (int)total;
  auto st = (cudaStream_t)stream;
  if (version != 1) return launch_attn_fwd2(maps, p, st);
  if (n_src > 1) {
    if (hs <= 16) return launch_attn_fwd<16, true>(maps, p, st);
    if (hs <= 32) return launch_attn_fwd<32, true>(maps, p, st);
    return launch_attn_fwd<64, true>(maps, p, st);
  }
  if (hs <= 16) return launch_attn_fwd<16, false>(maps, p, st);
  if (hs <= 32) return launch_attn_fwd<32, false>(maps, p, st);
  return launch_attn_fwd<64, false>(maps, p, st);
}
