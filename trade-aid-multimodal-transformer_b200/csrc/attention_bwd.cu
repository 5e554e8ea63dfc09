// K3 backward — recompute attention backward on tcgen05/TMEM (autograd of /root/reference/model.py:60-73, 136-159).
//
// Persistent, warp-specialised: one CTA per SM walks (group, batch, head) work items; inside an item the whole causal
// T x T triangle (T <= 256) is walked in 128 x 128 tile pairs, source s -> key/value tile j -> query tile i >= j, with
// every accumulator resident in TMEM (512 columns):
//   S, dP          2 x 128 columns   recomputed per pair:  S = Q_i K_j^T,  dP = dO_i V_j^T
//   dV_j, dK_j     2 x 64  columns   accumulated over the query tiles i >= j
//   dQ_0, dQ_1     2 x 64  columns   accumulated over key/value tiles j <= i AND over all key/value sources
//   warp 16  loader : TMA Q/dO tiles of the head, K/V tiles through a 2-stage ring (runs ahead)
//   warp 17  issuer : S and dP of pair n+1 are issued as soon as the math warps have pulled S / dP of pair n out of
//                     TMEM; the three gradient UMMAs of pair n run while the math warps work on pair n+1
//   warps 0-15 math : thread (row r, column quarter) owns 32 columns: P = exp2(S*c - lse), dS = P * (dP - D) with
//                     D = rowsum(dO * O_src) computed once per head straight from global memory; P and dS go to
//                     128B-swizzled smem once and are consumed by three UMMAs through two descriptor views of the same
//                     bytes: K-major (dQ += dS K) and MN-major (dV += P^T dO, dK += dS^T Q). The dK/dV/dQ read-outs are
//                     deferred behind the next pair's math so the gradient UMMAs never stall the math warps.
// K_j, V_j, Q_i, dO_i are read straight from their TMA tiles as MN-major B operands. hs^-0.5 is applied to dQ / dK at
// read-out. No T x T tensor and no atomics; dq/dk/dv are written once, in bf16.
// The previous version ran S/dP -> math -> gradient UMMAs strictly one after another in a 9-warp CTA (tensor pipe 8%
// active, 43% of the samples waiting on mbarriers: profiles/r01_ncu_attn_v2.txt).
#include <cmath>
#include <cstdlib>
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "common.cuh"
#include "philox.cuh"
#include "ptx.cuh"

namespace ta {

constexpr int ATTB_MAX_SRC = 8;
int make_head_map(CUtensorMap* out, const ta_attn_tensor* t, int hs, int H, int T, int B, int G, int box_rows);

struct AttnBwdMaps {
  CUtensorMap q, d_o;
  CUtensorMap k[ATTB_MAX_SRC];
  CUtensorMap v[ATTB_MAX_SRC];
  CUtensorMap dq;                  // TMA store targets (same head geometry as q / k / v)
  CUtensorMap dk[ATTB_MAX_SRC];
  CUtensorMap dv[ATTB_MAX_SRC];
};

struct GradOut {
  __nv_bfloat16* ptr;
  long long ld, head_stride, gs;
};

struct AttnBwdParams {
  int T, B, H, G, hs, n_src, n_qt, total_work;
  long long* trace;   // TA_ATTN_TRACE=<device address> (dev only): clock64 stamps of CTA 0, 4 roles x 1024 slots
  uint32_t drop_thresh;            // dropout on the probabilities: 16-bit lane threshold, 0 = off (philox.cuh)
  float drop_scale;                // 1 / (1 - p)
  unsigned long long drop_seed;
  const unsigned long long* drop_step;
  int debug;   // TA_ATTN_DEBUG (dev only): 1 = no D/lse prologue loads, 2 = no read-out stores, 4 = no P/dS math
  float scale, scale_log2;
  const __nv_bfloat16* d_o;    // [G][B*T][ldo]
  const __nv_bfloat16* o;      // [G][B*T][ldo]
  const __nv_bfloat16* o_src;  // [n_src] x same, or NULL when n_src == 1
  long long ldo, o_gs, o_src_stride;
  const float* lse;            // [n_src][G][B][H][T]
  GradOut dq;
  GradOut dk[ATTB_MAX_SRC];
  GradOut dv[ATTB_MAX_SRC];
};

// dev-only timeline: role 0 loader, 1 S/dP issuer, 2 math (warp 0), 3 gradient issuer; slot = running event index
#ifndef TA_ATTN_TRACE_BUILD
#define AB_TRACE(role, cnt, tag) do { (void)(cnt); } while (0)      // compiled out: -DTA_ATTN_TRACE_BUILD brings the timeline back
#else
#define AB_TRACE(role, cnt, tag)                                                                     \
  do {                                                                                               \
    if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (cnt) < 1024) {                     \
      p.trace[(role) * 1024 + (cnt)] = (clock64() << 8) | (tag);                                     \
      ++(cnt);                                                                                       \
    }                                                                                                \
  } while (0)
#endif

constexpr int AB_MATH_WARPS = 16;
constexpr int AB_MATH_THREADS = AB_MATH_WARPS * 32;
// warps 0-15 math, 16 loader, 17 S/dP issuer, 18 gradient issuer, 19 storer (TMA stores of dK/dV/dQ)
constexpr int AB_THREADS = (AB_MATH_WARPS + 4) * 32;
constexpr int AB_TILE = 16384;                          // 128 rows x 128 B

// The (work, s, j, i) pair schedule; every warp role walks the same sequence with its own copy.
struct AbIter {
  int work, s, j, i;
  int n, ks, hc;   // running counters over the CTA's lifetime: pairs, (s, j) key/value tiles, heads
  __device__ __forceinline__ void init(int first) { work = first; s = j = i = 0; n = ks = hc = 0; }
  __device__ __forceinline__ bool first_of_head() const { return s == 0 && j == 0 && i == 0; }
  __device__ __forceinline__ bool first_i() const { return i == j; }
  __device__ __forceinline__ void next(int n_qt, int n_src, int stride) {
    ++n;
    if (i + 1 < n_qt) { ++i; return; }
    ++ks;
    if (j + 1 < n_qt) { ++j; i = j; return; }
    j = i = 0;
    if (s + 1 < n_src) { ++s; return; }
    s = 0; ++hc;
    work += stride;
  }
};

// barrier table (index * 8 bytes from the table base)
enum : uint32_t {
  AB_Q_FULL = 0, AB_Q_EMPTY = 2, AB_KV_FULL = 4, AB_KV_EMPTY = 6, AB_S_FULL = 8, AB_DP_FULL = 10, AB_S_EMPTY = 12,
  AB_DP_EMPTY = 14, AB_PDS_FULL = 16, AB_G_DONE = 17, AB_DKV_EMPTY = 18, AB_DQ_EMPTY = 19,
  AB_SKV_FULL = 20, AB_SQ_FULL = 21, AB_NBARS = 22
};

// DROPOUT is a template parameter: the mask code sat in the middle of the hot loop (two Philox draws and sixteen selects per
// 16-column piece) and was executed-around on every pair of the common p = 0 case
template <int DP, bool DROPOUT>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ AttnBwdMaps maps, const __grid_constant__ AttnBwdParams p) {
  constexpr int TILE = AB_TILE;
  constexpr int OC = DP / 2;   // gradient columns per math thread at read-out
  extern __shared__ __align__(1024) uint8_t smem[];
  // [Q 2 slots][dO 2 slots][2 stages x (K, V)][P: 2 column halves][dS: 2 column halves][barriers][D][lse]
  // Dead operand tiles double as the staging tiles of the gradient read-outs (TMA stores): dK_j / dV_j go into the
  // K_j / V_j tiles of their own ring stage, dQ into the Q/dO slot of the head's last query tile; the STORER warp
  // releases those tiles to the loader once the TMA store has drained them.
  const uint32_t sQ = smem_u32(smem), sdO = sQ + 2 * TILE, sKV = sQ + 4 * TILE, sP = sQ + 8 * TILE, sdS = sQ + 10 * TILE;
  const uint32_t bb = sQ + 12 * TILE;                        // barrier i at bb + 8 * i
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(smem + 12 * TILE + AB_NBARS * 8);
  const uint32_t sD = bb + AB_NBARS * 8 + 16;                // fp32 [n_src][256] rowsum(dO * O_src)
  const uint32_t sLse = sD + ATTB_MAX_SRC * 256 * 4;         // fp32 [n_src][256] lse * log2(e)
  if ((sQ & 1023u) != 0) __trap();

  // the shuffle tells the compiler that `warp` is warp-uniform: the issuer branches then compute their descriptors in
  // uniform registers instead of moving every tcgen05.mma operand across with R2UR
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    for (uint32_t x = 0; x < AB_NBARS; ++x) {
      const bool wide = x == AB_S_EMPTY || x == AB_S_EMPTY + 1 || x == AB_DP_EMPTY || x == AB_DP_EMPTY + 1 ||
                        x == AB_PDS_FULL || x == AB_DKV_EMPTY || x == AB_DQ_EMPTY || x == AB_SKV_FULL || x == AB_SQ_FULL;
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bb + 8 * x), "r"(wide ? AB_MATH_THREADS : 1));
    }
    fence_mbar_init();
  }
  if (warp == AB_MATH_WARPS + 1) {
    tmem_alloc(tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_launch_dependents();       // PDL: on-chip set-up done; wait for the previous kernel before touching global memory
  pdl_wait();
  // TMEM columns: S halves [0,128) dP halves [128,256) dV [256,320) dK [320,384) dQ_0 [384,448) dQ_1 [448,512)
  const int n_qt = p.n_qt, n_src = p.n_src, stride = gridDim.x, total = p.total_work;
  AbIter it;
  it.init(blockIdx.x);
  // Q/dO smem slot of query tile ti: two tiles per head use slot = tile, one tile per head alternates between heads
  auto slot = [&](const AbIter& a, int ti) { return n_qt == 2 ? ti : (a.hc & 1); };
  auto slot_use = [&](const AbIter& a) { return n_qt == 2 ? a.hc : (a.hc >> 1); };
  auto bar = [&](uint32_t i) { return reinterpret_cast<uint64_t*>(smem + 12 * TILE) + i; };

  if (warp >= AB_MATH_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == AB_MATH_WARPS) {
      // ================================================================== loader (one lane)
      if ((tid & 31) == 0) {
        int tc = 0;
        auto load_kv = [&](const AbIter& a, int h, int b, int g) {
          const int st = a.ks & 1;
          mbar_wait_a(bb + 8 * (AB_KV_EMPTY + st), ((a.ks >> 1) & 1) ^ 1);
          AB_TRACE(0, tc, 1);
          mbar_expect_tx(bar(AB_KV_FULL + st), 2 * TILE);
          uint8_t* dK = smem + 4 * TILE + st * 2 * TILE;
          tma_load_5d(dK, &maps.k[a.s], bar(AB_KV_FULL + st), 0, h, a.j * 128, b, g);
          tma_load_5d(dK + TILE, &maps.v[a.s], bar(AB_KV_FULL + st), 0, h, a.j * 128, b, g);
        };
        for (; it.work < total; it.next(n_qt, n_src, stride)) {
          if (!it.first_i()) continue;       // one K/V tile per (s, j); Q/dO once per head
          int w = it.work;
          const int h = w % p.H; w /= p.H;
          const int b = w % p.B;
          const int g = w / p.B;
          if (it.first_of_head()) {
            // tile 0 is released two pairs before the end of the previous head, so its reload (and the first pair of
            // this head) overlaps the previous head's tail; tile 1 lands during the first pair
            for (int i = 0; i < n_qt; ++i) {
              const int sl = slot(it, i);
              mbar_wait_a(bb + 8 * (AB_Q_EMPTY + sl), (slot_use(it) & 1) ^ 1);
              AB_TRACE(0, tc, 2 + sl);
              mbar_expect_tx(bar(AB_Q_FULL + sl), 2 * TILE);
              tma_load_5d(smem + sl * TILE, &maps.q, bar(AB_Q_FULL + sl), 0, h, i * 128, b, g);
              tma_load_5d(smem + 2 * TILE + sl * TILE, &maps.d_o, bar(AB_Q_FULL + sl), 0, h, i * 128, b, g);
              if (i == 0) load_kv(it, h, b, g);   // the first pair needs K/V tile 0 next: before waiting for slot 1
            }
          } else {
            load_kv(it, h, b, g);
          }
        }
      }
    } else if (warp == AB_MATH_WARPS + 3) {
      // ================================================================== storer (one lane): TMA stores of the read-outs
      if ((tid & 31) == 0) {
        int ekv = 0, eq = 0;
        for (; it.work < total; it.next(n_qt, n_src, stride)) {
          if (it.i != n_qt - 1) continue;
          int w = it.work;
          const int h = w % p.H; w /= p.H;
          const int b = w % p.B;
          const int g = w / p.B;
          // dK_j, dV_j of the key/value tile whose last pair this is: staged in the K/V tiles of its own stage
          const int st = it.ks & 1;
          mbar_wait_a(bb + 8 * AB_SKV_FULL, ekv & 1);
          if (!(p.debug & 2)) {
            tma_store_5d(&maps.dk[it.s], sKV + st * 2 * TILE, 0, h, it.j * 128, b, g);
            tma_store_5d(&maps.dv[it.s], sKV + st * 2 * TILE + TILE, 0, h, it.j * 128, b, g);
          }
          bulk_commit_group();
          bulk_wait_read<0>();
          mbar_arrive_a(bb + 8 * (AB_KV_EMPTY + st));      // the loader may refill this stage
          ++ekv;
          if (it.j == n_qt - 1 && it.s == n_src - 1) {   // last pair of the head: dQ tiles, staged in the slot of tile n_qt - 1
            const int sl = slot(it, n_qt - 1);
            mbar_wait_a(bb + 8 * AB_SQ_FULL, eq & 1);
            if (!(p.debug & 2)) {
              tma_store_5d(&maps.dq, sQ + sl * TILE, 0, h, 0, b, g);
              if (n_qt > 1) tma_store_5d(&maps.dq, sdO + sl * TILE, 0, h, 128, b, g);
            }
            bulk_commit_group();
            bulk_wait_read<0>();
            mbar_arrive_a(bb + 8 * (AB_Q_EMPTY + sl));     // the loader may refill this slot
            ++eq;
          }
        }
        bulk_wait_all<0>();
      }
    } else {
      // ================================================================== UMMA issuers (whole warp, one elected lane issues)
      // warp 17: S = Q K^T and dP = dO V^T of every pair, as early as the math warps free the TMEM tiles
      // warp 18: the three gradient UMMAs of every pair, as soon as P / dS are in smem
      // Two warps because issuing costs ~50 cycles per tcgen05.mma here: one warp serialised 40 of them per pair
      // (2000 cycles, longer than the tensor pipe needs to execute them).
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);    // S, dP column halves
      constexpr uint32_t idesc_kv = umma_idesc_bf16(128, DP, 1, 1);   // dV, dK: A = P/dS seen MN-major, B = dO/Q MN-major
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, DP, 0, 1);    // dQ: A = dS K-major, B = K MN-major
      // descriptors advance by adding (bytes >> 4) to the low word: the address field sits in bits [0, 14)
      const uint64_t d_k = umma_desc_sw128(0, 16, 1024);       // K-major operand, 128-byte rows
      const uint64_t d_mn = umma_desc_sw128(0, TILE, 1024);    // MN-major view of a 128 x 128 B tile
      auto dk = [&](uint32_t addr) { return d_k + ((addr & 0x3FFFF) >> 4); };
      auto dmn = [&](uint32_t addr) { return d_mn + ((addr & 0x3FFFF) >> 4); };
      int tc = 0;
      if (warp == AB_MATH_WARPS + 1) {
        const int nk_qk = (p.hs + 15) / 16;
        for (; it.work < total; it.next(n_qt, n_src, stride)) {
          const AbIter& a = it;
          const int sl = slot(a, a.i), st = a.ks & 1;
          // first pair of the head that touches query tile i: (s = 0, j = 0, i)
          if (a.s == 0 && a.j == 0) mbar_wait_a(bb + 8 * (AB_Q_FULL + sl), slot_use(a) & 1);
          if (a.first_i()) mbar_wait_a(bb + 8 * (AB_KV_FULL + st), (a.ks >> 1) & 1);
          AB_TRACE(1, tc, 1);   // operands of S/dP in smem
          const uint64_t aQ = dk(sQ + sl * TILE), aO = dk(sdO + sl * TILE), aK = dk(sKV + st * 2 * TILE), aV = dk(sKV + st * 2 * TILE + TILE);
          // 64 key/value columns (half hh) at a time; each half has its own TMEM tile and barriers, so the math warps
          // release a tile after a single 16-column read
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            mbar_wait_a(bb + 8 * (AB_S_EMPTY + hh), (a.n & 1) ^ 1);
            tc_fence_after();
            for (int kk = 0; kk < nk_qk; ++kk)
              umma_bf16_ws(tmem_base + hh * 64, aQ + kk * 2, aK + hh * 512 + kk * 2, idesc_s, kk != 0);
            umma_commit_ws(bb + 8 * (AB_S_FULL + hh));
            mbar_wait_a(bb + 8 * (AB_DP_EMPTY + hh), (a.n & 1) ^ 1);
            tc_fence_after();
            for (int kk = 0; kk < nk_qk; ++kk)
              umma_bf16_ws(tmem_base + 128 + hh * 64, aO + kk * 2, aV + hh * 512 + kk * 2, idesc_s, kk != 0);
            umma_commit_ws(bb + 8 * (AB_DP_FULL + hh));
            AB_TRACE(1, tc, 2 + hh);   // S/dP half issued
          }
        }
      } else {
        for (; it.work < total; it.next(n_qt, n_src, stride)) {
          const AbIter& a = it;
          const int st = a.ks & 1, sl = slot(a, a.i);
          mbar_wait_a(bb + 8 * AB_PDS_FULL, a.n & 1);
          if (a.first_i()) mbar_wait_a(bb + 8 * AB_DKV_EMPTY, (a.ks & 1) ^ 1);        // dV/dK of the previous tile were read out
          if (a.first_of_head()) mbar_wait_a(bb + 8 * AB_DQ_EMPTY, (a.hc & 1) ^ 1);   // dQ of the previous head was read out
          tc_fence_after();
          AB_TRACE(3, tc, 4);   // pds_full seen
          const uint64_t mQ = dmn(sQ + sl * TILE), mO = dmn(sdO + sl * TILE), mK = dmn(sKV + st * 2 * TILE);
          const uint64_t mP = dmn(sP), mS = dmn(sdS), kS = dk(sdS);
          const uint32_t acc_kv = a.first_i() ? 0u : 1u, acc_q = (a.s == 0 && a.j == 0) ? 0u : 1u;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)     // dV_j += P^T dO_i   (k-step = 16 query rows = 2048 B in both tiles)
            umma_bf16_ws(tmem_base + 256, mP + kk * 128, mO + kk * 128, idesc_kv, acc_kv | (kk != 0));
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)     // dK_j += dS^T Q_i
            umma_bf16_ws(tmem_base + 320, mS + kk * 128, mQ + kk * 128, idesc_kv, acc_kv | (kk != 0));
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)     // dQ_i += dS K_j   (A: 32 B per k-step inside a column half, halves TILE apart)
            umma_bf16_ws(tmem_base + 384 + 64 * a.i, kS + (kk >> 2) * (TILE >> 4) + (kk & 3) * 2, mK + kk * 128, idesc_q,
                         acc_q | (kk != 0));
          umma_commit_ws(bb + 8 * AB_G_DONE);
          AB_TRACE(3, tc, 5);   // grads issued
          // last pair that reads query tile i < n_qt - 1: its slot goes back to the loader. (The K/V stage and the slot
          // of the last query tile are released by the storer: they stage the dK/dV and dQ read-outs first.)
          if (a.s == n_src - 1 && a.j == a.i && a.i != n_qt - 1) umma_commit_ws(bb + 8 * (AB_Q_EMPTY + sl));
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ================================================================== math warps
    const int r = tid & 127, qd = tid >> 7, wq = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    // deferred read-outs, packed: bit 0 dK/dV pending, bit 1 dQ pending, bit 2/3 their g_done parities,
    // bit 4 ring stage of the pending key/value tile, bit 5 Q/dO slot that stages dQ
    uint32_t pend = 0;
    // TMEM -> scale -> bf16 -> staging tile row r (128B-swizzled, the layout of the TMA store box), RC columns at a time
    constexpr int RC = OC < 16 ? OC : 16;
    auto read_stage = [&](uint32_t taddr, uint32_t stile, int col0, float mul, bool do_ld) {
#pragma unroll
      for (int c = 0; c < OC; c += RC) {
        uint32_t v[RC];
        if (do_ld) {
          tmem_ld_cols<RC>(taddr + col0 + c, v);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < RC; e += 8) {
            uint4 u;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              h2[q] = __floats2bfloat162_rn(__uint_as_float(v[e + 2 * q]) * mul, __uint_as_float(v[e + 2 * q + 1]) * mul);
            sts128(stile + r * 128 + ((((col0 + c + e) >> 3) ^ (r & 7)) << 4), u);
          }
        }
      }
    };
    // dK_j (quarters 2, 3) -> K tile, dV_j (quarters 0, 1) -> V tile of the stage the tile was loaded into (dead: every
    // UMMA that reads it is complete); the storer warp issues the TMA stores and hands the stage back to the loader
    auto readout_kv = [&]() {
      mbar_wait_a(bb + 8 * AB_G_DONE, (pend >> 2) & 1);
      tc_fence_after();
      const bool is_k = qd >= 2;
      read_stage(tmem_base + (is_k ? 320 : 256) + lane_off, sKV + ((pend >> 4) & 1) * 2 * TILE + (is_k ? 0 : TILE), (qd & 1) * OC,
                 is_k ? p.scale : 1.f, true);
      tc_fence_before();
      mbar_arrive_a(bb + 8 * AB_DKV_EMPTY);
      fence_proxy_async_smem();
      mbar_arrive_a(bb + 8 * AB_SKV_FULL);
      pend &= ~0x15u;
    };
    // dQ_0 (quarters 0, 1) -> Q tile, dQ_1 (quarters 2, 3) -> dO tile of the slot of the head's last query tile
    auto readout_q = [&]() {
      mbar_wait_a(bb + 8 * AB_G_DONE, (pend >> 3) & 1);
      tc_fence_after();
      const int ti = qd >> 1;
      read_stage(tmem_base + 384 + 64 * ti + lane_off, (ti ? sdO : sQ) + ((pend >> 5) & 1) * TILE, (qd & 1) * OC, p.scale, ti < n_qt);
      tc_fence_before();
      mbar_arrive_a(bb + 8 * AB_DQ_EMPTY);
      fence_proxy_async_smem();
      mbar_arrive_a(bb + 8 * AB_SQ_FULL);
      pend &= ~0x2au;
    };
    // per-thread piece of the D / lse prologue: row tid >> 1, 16-byte vectors 2v + (tid & 1) of the head's hs columns
    const int drow = tid >> 1, dpart = tid & 1, nvec = p.hs >> 3;
    int g2 = 0, b2 = 0, h2 = 0;
    int tc = 0;
    const bool ttr = warp == 0;
    int g = 0, b = 0, h = 0;

    for (; it.work < total; it.next(n_qt, n_src, stride)) {
      if (it.first_of_head()) {
        if (it.hc == 0) {
          int w = it.work;
          h = w % p.H; w /= p.H;
          b = w % p.B;
          g = w / p.B;
        } else {
          g = g2; b = b2; h = h2;      // decoded one head ahead (below)
        }
        // D[s][row] = rowsum(dO * O_s) and lse * log2(e) for the whole head, straight from global memory (two threads
        // per row). The NEXT head's inputs are prefetched into L2 here, a whole head ahead of their use; this head's
        // dO and first-source loads are issued before the pending read-outs so that their latency overlaps them.
        // (Eight lanes per row - whole 128-byte lines per load instruction - measured slower: 0.61 vs 0.57 ms.)
        if (ttr) AB_TRACE(2, tc, 7);            // prologue start
        named_bar_sync(5, AB_MATH_THREADS);     // sD / sLse of the previous head are no longer read
        if (ttr) AB_TRACE(2, tc, 8);            // first barrier passed
        const bool rok = drow < p.T && !(p.debug & 1);
        const long long roff = ((long long)b * p.T + drow) * p.ldo + h * p.hs;
        if (it.work + stride < total) {
          int w = it.work + stride;
          h2 = w % p.H; w /= p.H;
          b2 = w % p.B;
          g2 = w / p.B;
          if (rok && dpart == 0) {
            const long long roff2 = ((long long)b2 * p.T + drow) * p.ldo + h2 * p.hs;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.d_o + (long long)g2 * p.o_gs + roff2));
            for (int s = 0; s < n_src; ++s) {
              const __nv_bfloat16* ob = p.o_src ? p.o_src + (long long)s * p.o_src_stride : p.o + (long long)g2 * p.o_gs;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(ob + roff2));
              if ((drow & 31) == 0)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.lse + ((((long long)s * p.G + g2) * p.B + b2) * p.H + h2) * p.T + drow));
            }
          }
        }
        if (ttr) AB_TRACE(2, tc, 14);           // prefetches issued
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        uint4 dvv[4], ov[4];
        float lv;
        auto load_src = [&](int s) {
          const __nv_bfloat16* ob = p.o_src ? p.o_src + (long long)s * p.o_src_stride : p.o + (long long)g * p.o_gs;
#pragma unroll
          for (int v = 0; v < 4; ++v)
            ov[v] = (rok && 2 * v + dpart < nvec) ? *reinterpret_cast<const uint4*>(ob + roff + (2 * v + dpart) * 8) : z4;
          lv = (rok && dpart == 0) ? p.lse[((((long long)s * p.G + g) * p.B + b) * p.H + h) * p.T + drow] : 0.f;
        };
#pragma unroll
        for (int v = 0; v < 4; ++v)
          dvv[v] = (rok && 2 * v + dpart < nvec) ? *reinterpret_cast<const uint4*>(p.d_o + (long long)g * p.o_gs + roff + (2 * v + dpart) * 8) : z4;
        load_src(0);
        if (ttr) AB_TRACE(2, tc, 11);           // loads issued
        if (pend & 1) readout_kv();
        if (ttr) AB_TRACE(2, tc, 13);           // kv read-out done
        if (pend & 2) readout_q();
        if (ttr) AB_TRACE(2, tc, 9);            // read-outs done
        for (int s = 0; s < n_src; ++s) {
          if (s > 0) load_src(s);
          float acc = 0.f;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&dvv[v]);
            const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov[v]);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              acc += __low2float(ah[e]) * __low2float(oh[e]) + __high2float(ah[e]) * __high2float(oh[e]);
          }
          acc += __shfl_xor_sync(0xffffffffu, acc, 1);
          if (dpart == 0) {
            sts_f32(sD + (s * 256 + drow) * 4, acc);
            sts_f32(sLse + (s * 256 + drow) * 4, lv * 1.4426950408889634f);
          }
        }
        if (ttr) AB_TRACE(2, tc, 10);           // D stored
        named_bar_sync(5, AB_MATH_THREADS);
        if (ttr) AB_TRACE(2, tc, 1);            // head prologue done
      }
      const int row = it.i * 128 + r;
      const bool diag = it.i == it.j;
      const float Lr = lds_f32(sLse + (it.s * 256 + row) * 4);
      const float Dr = lds_f32(sD + (it.s * 256 + row) * 4);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col_base = it.j * 128 + hh * 64 + qd * 16;
        // these 16 columns are needed unless they lie above the diagonal for every row of the warp (warp-uniform)
        const bool need = (!diag || col_base <= it.i * 128 + wq * 32 + 31) && !(p.debug & 4);
        int nvalid = !diag ? 16 : (row - col_base + 1 < 0 ? 0 : (row - col_base + 1 > 16 ? 16 : row - col_base + 1));
        if (row >= p.T) nvalid = 0;
        uint32_t pp[8], dd[8];     // 16 columns of P and dS, packed bf16x2
        mbar_wait_a(bb + 8 * (AB_S_FULL + hh), it.n & 1);
        mbar_wait_a(bb + 8 * (AB_DP_FULL + hh), it.n & 1);   // issued right behind S: no register is held across a wait
        tc_fence_after();
        if (ttr) AB_TRACE(2, tc, 2 + hh);   // S/dP half available
        if (need) {
          uint32_t sv[16], dv[16];
          tmem_ld_32x16(tmem_base + hh * 64 + lane_off + qd * 16, sv);
          tmem_ld_32x16(tmem_base + 128 + hh * 64 + lane_off + qd * 16, dv);
          tmem_wait_ld();
          tc_fence_before();
          mbar_arrive_a(bb + 8 * (AB_S_EMPTY + hh));
          mbar_arrive_a(bb + 8 * (AB_DP_EMPTY + hh));
          // dropout (model.py:69, 151): O = (P o M / (1-p)) V, so dV takes the dropped P~ = P o M / (1-p) and
          // dS = P o (dP o M / (1-p) - D); the mask is regenerated from (seed, plane, row, column) as in the forward
          uint32_t keep = 0xFFFFu;
          if constexpr (DROPOUT) {
            const unsigned long long sd = p.drop_seed + (p.drop_step ? *p.drop_step * 0x9E3779B97F4A7C15ull : 0ull);
            const uint32_t plane = (uint32_t)it.s * (uint32_t)total + (uint32_t)it.work;
            keep = dropout_keep8(sd, 1u, plane, (uint32_t)row, (uint32_t)(col_base >> 3), p.drop_thresh) |
                   (dropout_keep8(sd, 1u, plane, (uint32_t)row, (uint32_t)(col_base >> 3) + 1, p.drop_thresh) << 8);
          }
          const float dsc = p.drop_scale;
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float p0 = ex2_approx(__uint_as_float(sv[e]) * p.scale_log2 - Lr);
            float p1 = ex2_approx(__uint_as_float(sv[e + 1]) * p.scale_log2 - Lr);
            p0 = e < nvalid ? p0 : 0.f;
            p1 = e + 1 < nvalid ? p1 : 0.f;
            float m0 = 1.f, m1 = 1.f;
            if constexpr (DROPOUT) {
              m0 = (keep >> e) & 1u ? dsc : 0.f;
              m1 = (keep >> (e + 1)) & 1u ? dsc : 0.f;
            }
            const float d0 = p0 * (__uint_as_float(dv[e]) * m0 - Dr), d1 = p1 * (__uint_as_float(dv[e + 1]) * m1 - Dr);
            __nv_bfloat162 pv = __floats2bfloat162_rn(p0 * m0, p1 * m1), dh = __floats2bfloat162_rn(d0, d1);
            pp[e / 2] = *reinterpret_cast<uint32_t*>(&pv);
            dd[e / 2] = *reinterpret_cast<uint32_t*>(&dh);
          }
        } else {
          tc_fence_before();
          mbar_arrive_a(bb + 8 * (AB_S_EMPTY + hh));
          mbar_arrive_a(bb + 8 * (AB_DP_EMPTY + hh));
#pragma unroll
          for (int e = 0; e < 8; ++e) pp[e] = dd[e] = 0u;
        }
        if (hh == 0) {
          if (ttr) AB_TRACE(2, tc, 4);      // half 0 math done
          // P/dS smem is free once the gradient UMMAs of the previous pair are complete (they had the math above to run)
          mbar_wait_a(bb + 8 * AB_G_DONE, (it.n & 1) ^ 1);
          if (ttr) AB_TRACE(2, tc, 5);      // g_done(n-1) seen
        }
        const uint32_t off = hh * TILE + r * 128;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t lc = ((qd * 2 + q) ^ (r & 7)) << 4;
          sts128(sP + off + lc, make_uint4(pp[4 * q], pp[4 * q + 1], pp[4 * q + 2], pp[4 * q + 3]));
          sts128(sdS + off + lc, make_uint4(dd[4 * q], dd[4 * q + 1], dd[4 * q + 2], dd[4 * q + 3]));
        }
        if (hh == 0) {
          // deferred read-out of the previous key/value tile (its last gradient UMMAs are the ones waited for above);
          // placed here because no P/dS registers are live between the two column halves
          if (pend & 1) readout_kv();
        }
      }
      fence_proxy_async_smem();
      mbar_arrive_a(bb + 8 * AB_PDS_FULL);
      if (ttr) AB_TRACE(2, tc, 6);          // pds_full arrive
      if (it.i == n_qt - 1)   // last pair of this key/value tile: dK/dV read-out after the next pair's first half
        pend = (pend & ~0x14u) | 1u | ((it.n & 1u) << 2) | ((uint32_t)(it.ks & 1) << 4);
      if (it.i == n_qt - 1 && it.j == n_qt - 1 && it.s == n_src - 1)   // last pair of the head
        pend = (pend & ~0x28u) | 2u | ((it.n & 1u) << 3) | ((uint32_t)slot(it, n_qt - 1) << 5);
    }
    if (pend & 1) readout_kv();
    if (pend & 2) readout_q();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AB_MATH_WARPS + 1) tmem_dealloc(tmem_base, 512);
}

template <int DP, bool DROPOUT>
static int launch_attn_bwd_d(const AttnBwdMaps& maps, const AttnBwdParams& p, cudaStream_t st) {
  constexpr int SMEM = 12 * AB_TILE + AB_NBARS * 8 + 16 + 2 * ATTB_MAX_SRC * 256 * 4 + 64;
  static bool attr[TA_MAX_DEVICES] = {};
  TA_CHECK_CUDA(ensure_dynamic_smem(attn_bwd_kernel<DP, DROPOUT>, SMEM, attr));
  const int grid = p.total_work < num_sms() ? p.total_work : num_sms();
  launch_kernel(attn_bwd_kernel<DP, DROPOUT>, dim3(grid), dim3(AB_THREADS), SMEM, st, maps, p);
  TA_CHECK_LAUNCH();
  return 0;
}

template <int DP>
static int launch_attn_bwd(const AttnBwdMaps& maps, const AttnBwdParams& p, cudaStream_t st) {
  return p.drop_thresh ? launch_attn_bwd_d<DP, true>(maps, p, st) : launch_attn_bwd_d<DP, false>(maps, p, st);
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
  TA_CHECK_ARG(q && k && v && o_bf16 && do_bf16 && lse && dq && dk && dv, "ta_attention_bwd: null pointer");
  TA_CHECK_ARG(n_src >= 1 && n_src <= ATTB_MAX_SRC, "ta_attention_bwd: n_src must be in [1, %d]", ATTB_MAX_SRC);
  TA_CHECK_ARG(n_src == 1 || o_src_bf16, "ta_attention_bwd: per-source outputs are required when n_src > 1");
  TA_CHECK_ARG(hs >= 8 && hs <= 64 && hs % 8 == 0, "ta_attention_bwd: head_size %d not supported", hs);
  TA_CHECK_ARG(T >= 1 && T <= 256, "ta_attention_bwd: T = %d not supported in this build (T <= 256)", T);
  TA_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "ta_attention_bwd: dropout_p must be in [0, 1)");
  TA_CHECK_ARG(ldo % 8 == 0 && o_gs % 8 == 0, "ta_attention_bwd: output strides must be multiples of 8");
  AttnBwdMaps maps;
  if (make_head_map(&maps.q, q, hs, H, T, B, G, 128)) return 1;
  ta_attn_tensor dot{do_bf16, ldo, hs, o_gs};
  if (make_head_map(&maps.d_o, &dot, hs, H, T, B, G, 128)) return 1;
  TA_CHECK_ARG(n_src == 1 || G == 1, "ta_attention_bwd: n_src > 1 requires G == 1");
  AttnBwdParams p{};
  for (int s = 0; s < n_src; ++s) {
    if (make_head_map(&maps.k[s], &k[s], hs, H, T, B, G, 128)) return 1;
    if (make_head_map(&maps.v[s], &v[s], hs, H, T, B, G, 128)) return 1;
    if (grad_out(&p.dk[s], &dk[s], G) || grad_out(&p.dv[s], &dv[s], G)) return 1;
    if (make_head_map(&maps.dk[s], &dk[s], hs, H, T, B, G, 128)) return 1;
    if (make_head_map(&maps.dv[s], &dv[s], hs, H, T, B, G, 128)) return 1;
  }
  if (grad_out(&p.dq, dq, G)) return 1;
  if (make_head_map(&maps.dq, dq, hs, H, T, B, G, 128)) return 1;
  p.T = T; p.B = B; p.H = H; p.G = G; p.hs = hs; p.n_src = n_src; p.n_qt = (T + 127) / 128;
  p.scale = (float)(1.0 / sqrt((double)hs));
  p.scale_log2 = (float)(1.4426950408889634 / sqrt((double)hs));
  const long long total = (long long)G * B * H;
  TA_CHECK_ARG(total < (1ll << 31), "attention grid too large");
  p.total_work = (int)total;
  { const char* e = getenv("TA_ATTN_DEBUG"); p.debug = e ? atoi(e) : 0; }
  p.drop_thresh = dropout_p > 0.f ? dropout_thresh16(dropout_p) : 0u;
  p.drop_scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  p.drop_seed = seed;
  p.drop_step = dropout_step_counter();
  { const char* e = getenv("TA_ATTN_TRACE"); p.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr; }
  p.d_o = (const __nv_bfloat16*)do_bf16;
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
