// K3 backward — recompute attention backward on tcgen05/TMEM (autograd of /root/reference/model.py:60-73, 136-159).
//
// Persistent, warp-specialised: one CTA per SM walks (group, batch, head) work items; inside an item the whole causal
// T x T triangle (T <= 256) is walked in 128 x 128 tile pairs, source s -> key/value tile j -> query tile i >= j, with
// every accumulator resident in TMEM (512 columns):
//   S, dP          2 x 128 columns   recomputed per pair:  S = Q_i K_j^T,  dP = dO_i V_j^T
//   dV_j, dK_j     2 x 64  columns   accumulated over the query tiles i >= j
//   dQ_0, dQ_1     2 x 64  columns   accumulated over key/value tiles j <= i AND over all key/value sources
//   warp 16  loader : TMA Q/dO tiles of the head, K/V tiles through a 2-stage ring (runs ahead)
//   warp 17  issuer : S and dP of pair n+1 are issued as soon as the math warps have pulled S / dP of pair n out of TMEM
//   warp 18  issuer : the three gradient UMMAs of pair n run while the math warps work on pair n+1; its commits hand
//                     the K/V stage and the Q/dO slot back to the loader the moment their last reader has completed
//   warp 19  storer : TMA stores of the dK/dV/dQ read-outs from two staging tiles of their own
//   warps 0-15 math : thread (row r, column quarter) owns 32 columns: P = exp2(S*c - lse), dS = P * (dP - D); lse and
//                     D = rowsum(dO * O_src) come from global tables (D: attn_bwd_dtab_kernel below, one HBM pass over dO
//                     and the per-source outputs right before this kernel), loaded one pair ahead. P and dS go to
//                     128B-swizzled smem once and are consumed by three UMMAs through two descriptor views of the same
//                     bytes: K-major (dQ += dS K) and MN-major (dV += P^T dO, dK += dS^T Q). The dK/dV/dQ read-outs are
//                     deferred behind the next pair's math so the gradient UMMAs never stall the math warps.
// K_j, V_j, Q_i, dO_i are read straight from their TMA tiles as MN-major B operands. hs^-0.5 is applied to dQ / dK at
// read-out. No T x T tensor and no atomics; dq/dk/dv are written once, in bf16.
// History (CTA timelines: profiles/r02_attn_bwd_trace_*.txt). v1 ran S/dP -> math -> gradient UMMAs strictly one after
// another in a 9-warp CTA (tensor pipe 8% active). v2 overlapped them but (a) computed D in a per-head prologue of the math
// warps - 9 k of a 21 k-cycle head at the c4 self shape, 38 k of 116 k with 7 sources, every UMMA queue drained meanwhile -
// and (b) staged the read-outs in the dead K/V and Q/dO tiles, so a tile could only be reloaded after its TMA store had
// drained: the first pair of every source and the second pair of every head waited ~2500 cycles for operands. Storing the
// read-outs straight from registers instead (64 contiguous bytes per thread) was worse: one L1 tag cycle per row and
// instruction, 0.76 instead of 0.54 ms.
#include <cmath>
#include <cstdlib>
#include <cuda_bf16.h>
#include <type_traits>

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

struct AttnBwdParams {
  int T, B, H, G, hs, n_src, n_qt, total_work;
  long long* trace;   // TA_ATTN_TRACE=<device address> (dev only): clock64 stamps of CTA 0, 4 roles x 1024 slots
  uint32_t drop_thresh;            // dropout on the probabilities: 16-bit lane threshold, 0 = off (philox.cuh)
  float drop_scale;                // 1 / (1 - p)
  unsigned long long drop_seed;
  const unsigned long long* drop_step;
  int debug;   // TA_ATTN_DEBUG (dev only): 1 = no D/lse loads, 2 = no read-out stores, 4 = no P/dS math
  float scale, scale_log2;
  const float* lse;            // [n_src][G][B][H][T]
  const float* dtab;           // [n_src][G][B][H][T]  D = rowsum(dO * O_src)
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
// warps 0-15 math, 16 loader, 17 S/dP issuer, 18 gradient issuer, 19 storer (640 x 96 registers = 512 x 104 + 128 x 64)
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
  AB_ST_FULL = 20, AB_ST_EMPTY = 21, AB_NBARS = 22
};
constexpr int AB_NTILES = 14;   // Q, dO: 2 slots each; K/V: 2 stages; P, dS: 2 column halves each; 2 read-out staging tiles

// DROPOUT is a template parameter: the mask code sat in the middle of the hot loop (two Philox draws and sixteen selects per
// 16-column piece) and was executed-around on every pair of the common p = 0 case
template <int DP, bool DROPOUT>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ AttnBwdMaps maps, const __grid_constant__ AttnBwdParams p) {
  constexpr int TILE = AB_TILE;
  constexpr int OC = DP / 2;   // gradient columns per math thread at read-out
  extern __shared__ __align__(1024) uint8_t smem[];
  // [Q 2 slots][dO 2 slots][2 stages x (K, V)][P: 2 column halves][dS: 2 column halves][staging: 2 tiles][barriers]
  // The staging tiles take the dK_j / dV_j read-out of a key/value tile, or the two dQ tiles of a head, for the storer's
  // TMA stores; only the next read-out waits for them: dK/dV read-outs are at least a pair apart, and the dQ read-out of a
  // head follows its last dK/dV read-out half a pair later (middle / end of the next head's first pair).
  const uint32_t sQ = smem_u32(smem), sdO = sQ + 2 * TILE, sKV = sQ + 4 * TILE, sP = sQ + 8 * TILE, sdS = sQ + 10 * TILE;
  const uint32_t sSt = sQ + 12 * TILE;
  const uint32_t bb = sQ + AB_NTILES * TILE;                 // barrier i at bb + 8 * i
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(smem + AB_NTILES * TILE + AB_NBARS * 8);
  if ((sQ & 1023u) != 0) __trap();

  // the shuffle tells the compiler that `warp` is warp-uniform: the issuer branches then compute their descriptors in
  // uniform registers instead of moving every tcgen05.mma operand across with R2UR
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    for (uint32_t x = 0; x < AB_NBARS; ++x) {
      const bool wide = x == AB_S_EMPTY || x == AB_S_EMPTY + 1 || x == AB_DP_EMPTY || x == AB_DP_EMPTY + 1 ||
                        x == AB_PDS_FULL || x == AB_DKV_EMPTY || x == AB_DQ_EMPTY || x == AB_ST_FULL;
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
  auto bar = [&](uint32_t i) { return reinterpret_cast<uint64_t*>(smem + AB_NTILES * TILE) + i; };

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
      // jobs in the order the math warps produce them: dK_j / dV_j after the last pair of a key/value tile, then - after
      // the last pair of a head - its dQ tiles
      if ((tid & 31) == 0) {
        uint32_t job = 0;
        auto drain = [&]() {
          bulk_commit_group();
          bulk_wait_read<0>();
          mbar_arrive_a(bb + 8 * AB_ST_EMPTY);      // the staging tiles may be overwritten
          ++job;
        };
        for (; it.work < total; it.next(n_qt, n_src, stride)) {
          if (it.i != n_qt - 1) continue;
          int w = it.work;
          const int h = w % p.H; w /= p.H;
          const int b = w % p.B;
          const int g = w / p.B;
          mbar_wait_a(bb + 8 * AB_ST_FULL, job & 1);
          if (!(p.debug & 2)) {
            tma_store_5d(&maps.dk[it.s], sSt, 0, h, it.j * 128, b, g);
            tma_store_5d(&maps.dv[it.s], sSt + TILE, 0, h, it.j * 128, b, g);
          }
          drain();
          if (it.j == n_qt - 1 && it.s == n_src - 1) {   // last pair of the head
            mbar_wait_a(bb + 8 * AB_ST_FULL, job & 1);
            if (!(p.debug & 2)) {
              tma_store_5d(&maps.dq, sSt, 0, h, 0, b, g);
              if (n_qt > 1) tma_store_5d(&maps.dq, sSt + TILE, 0, h, 128, b, g);
            }
            drain();
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
            AB_TRACE(1, tc, 6 + hh);   // S half issued
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
          // last pair that reads query tile i / key-value tile j: the slot / ring stage goes back to the loader when these
          // UMMAs are complete (the S / dP UMMAs of the pair completed before the math warps could produce P / dS)
          if (a.s == n_src - 1 && a.j == a.i) umma_commit_ws(bb + 8 * (AB_Q_EMPTY + sl));
          if (a.i == n_qt - 1) umma_commit_ws(bb + 8 * (AB_KV_EMPTY + st));
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ================================================================== math warps
    const int r = tid & 127, qd = tid >> 7, wq = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    // deferred read-outs, packed: bit 0 dK/dV pending, bit 1 dQ pending, bit 2/3 their g_done parities
    uint32_t pend = 0;
    uint32_t job = 0;          // read-outs staged so far (the storer counts the same sequence)
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
    // dK_j (quarters 2, 3) -> staging tile 0, dV_j (quarters 0, 1) -> staging tile 1: every UMMA that accumulates into
    // them is complete at g_done, and the storer has drained the previous read-out from the staging tiles
    auto readout_kv = [&]() {
      mbar_wait_a(bb + 8 * AB_G_DONE, (pend >> 2) & 1);
      mbar_wait_a(bb + 8 * AB_ST_EMPTY, (job & 1) ^ 1);
      tc_fence_after();
      const bool is_k = qd >= 2;
      read_stage(tmem_base + (is_k ? 320 : 256) + lane_off, sSt + (is_k ? 0 : TILE), (qd & 1) * OC, is_k ? p.scale : 1.f, true);
      tc_fence_before();
      mbar_arrive_a(bb + 8 * AB_DKV_EMPTY);
      fence_proxy_async_smem();
      mbar_arrive_a(bb + 8 * AB_ST_FULL);
      ++job;
      pend &= ~0x5u;
    };
    // dQ_0 (quarters 0, 1) -> staging tile 0, dQ_1 (quarters 2, 3) -> staging tile 1
    auto readout_q = [&]() {
      mbar_wait_a(bb + 8 * AB_G_DONE, (pend >> 3) & 1);
      mbar_wait_a(bb + 8 * AB_ST_EMPTY, (job & 1) ^ 1);
      tc_fence_after();
      const int ti = qd >> 1;
      read_stage(tmem_base + 384 + 64 * ti + lane_off, sSt + ti * TILE, (qd & 1) * OC, p.scale, ti < n_qt);
      tc_fence_before();
      mbar_arrive_a(bb + 8 * AB_DQ_EMPTY);
      fence_proxy_async_smem();
      mbar_arrive_a(bb + 8 * AB_ST_FULL);
      ++job;
      pend &= ~0xau;
    };
    int tc = 0;
    const bool ttr = warp == 0;
    // lse and D of this thread's row, loaded one pair ahead (an L2 round trip is as long as a fifth of a pair)
    const long long src_stride = (long long)p.G * p.B * p.H * p.T;
    auto table_base = [&](int work) {    // work = (g * B + b) * H + h
      return (long long)work * p.T;
    };
    float Ln = 0.f, Dn = 0.f;
    auto load_row = [&](const AbIter& a) {
      const int row = a.i * 128 + r;
      if (row < p.T && !(p.debug & 1)) {
        const long long idx = a.s * src_stride + table_base(a.work) + row;
        Ln = __ldg(p.lse + idx);
        Dn = __ldg(p.dtab + idx);
      } else {
        Ln = Dn = 0.f;
      }
    };
    if (it.work < total) load_row(it);

    for (; it.work < total; it.next(n_qt, n_src, stride)) {
      const int row = it.i * 128 + r;
      const bool diag = it.i == it.j;
      const float Lr = Ln * 1.4426950408889634f, Dr = Dn;
      {
        AbIter nx = it;
        nx.next(n_qt, n_src, stride);
        if (nx.work < total) load_row(nx);
      }
      if (ttr) AB_TRACE(2, tc, 16);         // pair top
      #pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col_base = it.j * 128 + hh * 64 + qd * 16;
        // these 16 columns are needed unless they lie above the diagonal for every row of the warp (warp-uniform)
        const bool need = (!diag || col_base <= it.i * 128 + wq * 32 + 31) && !(p.debug & 4);
        int nvalid = !diag ? 16 : (row - col_base + 1 < 0 ? 0 : (row - col_base + 1 > 16 ? 16 : row - col_base + 1));
        if (row >= p.T) nvalid = 0;
        uint32_t pp[8], dd[8];     // 16 columns of P and dS, packed bf16x2
        mbar_wait_a(bb + 8 * (AB_S_FULL + hh), it.n & 1);
        if (ttr) AB_TRACE(2, tc, 15);       // S half available
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
          // causal / ragged masking only where the warp's 32 rows x 16 columns touch the diagonal or the end of the
          // sequence (a select per element otherwise: 8 % of the kernel's instructions)
          auto pds16 = [&](auto masked) {
            constexpr bool MASK = decltype(masked)::value;
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float p0 = ex2_approx(__uint_as_float(sv[e]) * p.scale_log2 - Lr);
              float p1 = ex2_approx(__uint_as_float(sv[e + 1]) * p.scale_log2 - Lr);
              if constexpr (MASK) {
                p0 = e < nvalid ? p0 : 0.f;
                p1 = e + 1 < nvalid ? p1 : 0.f;
              }
              float x0 = __uint_as_float(dv[e]), x1 = __uint_as_float(dv[e + 1]), q0 = p0, q1 = p1;
              if constexpr (DROPOUT) {
                const float m0 = (keep >> e) & 1u ? dsc : 0.f, m1 = (keep >> (e + 1)) & 1u ? dsc : 0.f;
                x0 *= m0; x1 *= m1; q0 *= m0; q1 *= m1;
              }
              const float d0 = p0 * (x0 - Dr), d1 = p1 * (x1 - Dr);
              __nv_bfloat162 pv = __floats2bfloat162_rn(q0, q1), dh = __floats2bfloat162_rn(d0, d1);
              pp[e / 2] = *reinterpret_cast<uint32_t*>(&pv);
              dd[e / 2] = *reinterpret_cast<uint32_t*>(&dh);
            }
          };
          if ((!diag || col_base + 15 <= it.i * 128 + wq * 32) && it.i * 128 + 128 <= p.T) pds16(std::false_type{});
          else pds16(std::true_type{});
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
      // dQ of the previous head, at the end of this head's first pair: its dK/dV read-out went into the staging tiles half
      // a pair ago and has drained; the gradient UMMAs of this pair wait for it (AB_DQ_EMPTY)
      if (pend & 2) readout_q();
      if (it.i == n_qt - 1)   // last pair of this key/value tile: dK/dV read-out after the next pair's first half
        pend = (pend & ~0x4u) | 1u | ((it.n & 1u) << 2);
      if (it.i == n_qt - 1 && it.j == n_qt - 1 && it.s == n_src - 1)   // last pair of the head: dQ right behind it
        pend = (pend & ~0x8u) | 2u | ((it.n & 1u) << 3);
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
  constexpr int SMEM = AB_NTILES * AB_TILE + AB_NBARS * 8 + 16 + 64;
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

// D[s][g][b][h][t] = sum_c dO[g][b*T+t][h*hs+c] * O_s[...][h*hs+c]. Eight lanes per (g, b, h, four consecutive rows t), one
// 16-byte vector per row each (lanes beyond hs/8 idle); the dO vectors are reused for every source. One pass over dO and the
// per-source outputs: HBM-bound, so every thread keeps four independent 16-byte loads in flight and the grid is sized to the
// machine (one (row, head) per thread group and block launch ran at half the copy bandwidth).
__global__ void __launch_bounds__(256, 3)
attn_bwd_dtab_kernel(const __nv_bfloat16* __restrict__ d_o, const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ o_src,
                     long long ldo, long long o_gs, long long o_src_stride, float* __restrict__ dtab, int n_src, int G, int B, int T,
                     int H, int hs, int items) {      // items = G * B * ceil(T / 4) * H
  pdl_prologue();
  const int v = threadIdx.x & 7;
  const int TQ = (T + 3) >> 2;
  const long long table = (long long)G * B * H * T;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  for (int item = blockIdx.x * 32 + (threadIdx.x >> 3); item < items; item += gridDim.x * 32) {
    const int h = item % H, r1 = item / H;
    const int t0 = (r1 % TQ) * 4, r2 = r1 / TQ;
    const int b = r2 % B, g = r2 / B;
    const bool okv = v * 8 < hs;
    const long long ro = ((long long)b * T + t0) * ldo + h * hs + v * 8;
    const long long dst = (((long long)g * B + b) * H + h) * T + t0;
    uint4 dv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) dv[u] = okv && t0 + u < T ? *reinterpret_cast<const uint4*>(d_o + g * o_gs + ro + u * ldo) : z4;
    for (int s = 0; s < n_src; ++s) {
      const __nv_bfloat16* ob = o_src ? o_src + (long long)s * o_src_stride : o + g * o_gs;
      uint4 ov[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) ov[u] = okv && t0 + u < T ? *reinterpret_cast<const uint4*>(ob + ro + u * ldo) : z4;
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&dv[u]);
        const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov[u]);
        float a = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) a += __low2float(ah[e]) * __low2float(oh[e]) + __high2float(ah[e]) * __high2float(oh[e]);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        acc[u] = a;
      }
      if (v == 0) {
        float* d = dtab + s * table + dst;
        if ((T & 3) == 0) {
          *reinterpret_cast<float4*>(d) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (t0 + u < T) d[u] = acc[u];
        }
      }
    }
  }
}

}  // namespace ta

using namespace ta;

extern "C" int64_t ta_attention_bwd_workspace_bytes(int32_t n_src, int32_t G, int32_t B, int32_t T, int32_t H) {
  return (int64_t)n_src * G * B * H * T * (int64_t)sizeof(float);
}

extern "C" int32_t ta_attention_bwd(const ta_attn_tensor* q, const ta_attn_tensor* k, const ta_attn_tensor* v,
                                    int32_t n_src, const void* o_bf16, const void* o_src_bf16, int64_t o_src_stride,
                                    const void* do_bf16, int64_t ldo, int64_t o_gs, const float* lse, void* workspace,
                                    const ta_attn_tensor* dq, const ta_attn_tensor* dk, const ta_attn_tensor* dv,
                                    int32_t G, int32_t B, int32_t T, int32_t H, int32_t hs, float dropout_p, uint64_t seed,
                                    void* stream) {
  TA_CHECK_ARG(q && k && v && o_bf16 && do_bf16 && lse && dq && dk && dv && workspace, "ta_attention_bwd: null pointer");
  TA_CHECK_ARG((reinterpret_cast<uintptr_t>(o_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(do_bf16) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(o_src_bf16) & 15) == 0 && o_src_stride % 8 == 0 && hs % 8 == 0,
               "ta_attention_bwd: outputs must be 16-byte aligned");
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
    if (make_head_map(&maps.dk[s], &dk[s], hs, H, T, B, G, 128)) return 1;
    if (make_head_map(&maps.dv[s], &dv[s], hs, H, T, B, G, 128)) return 1;
  }
  if (make_head_map(&maps.dq, dq, hs, H, T, B, G, 128)) return 1;
  p.T = T; p.B = B; p.H = H; p.G = G; p.hs = hs; p.n_src = n_src; p.n_qt = (T + 127) / 128;
  p.scale = (float)(1.0 / sqrt((double)hs));
  p.scale_log2 = (float)(1.4426950408889634 / sqrt((double)hs));
  const long long total = (long long)G * B * H;
  TA_CHECK_ARG(total < (1ll << 26), "attention grid too large");
  p.total_work = (int)total;
  { const char* e = getenv("TA_ATTN_DEBUG"); p.debug = e ? atoi(e) : 0; }
  p.drop_thresh = dropout_p > 0.f ? dropout_thresh16(dropout_p) : 0u;
  p.drop_scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  p.drop_seed = seed;
  p.drop_step = dropout_step_counter();
  { const char* e = getenv("TA_ATTN_TRACE"); p.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr; }
  p.lse = lse;
  p.dtab = (const float*)workspace;
  auto st = (cudaStream_t)stream;
  {
    const long long items = total * ((T + 3) / 4);
    TA_CHECK_ARG(items < (1ll << 31), "ta_attention_bwd: G * B * H * T too large");
    long long blocks = (items + 31) / 32;
    if (blocks > 6ll * num_sms()) blocks = 6ll * num_sms();
    launch_kernel(attn_bwd_dtab_kernel, dim3((unsigned)blocks), dim3(256), 0, st, (const __nv_bfloat16*)do_bf16,
                  (const __nv_bfloat16*)o_bf16, n_src > 1 ? (const __nv_bfloat16*)o_src_bf16 : (const __nv_bfloat16*)nullptr,
                  (long long)ldo, (long long)o_gs, (long long)o_src_stride, (float*)workspace, (int)n_src, (int)G, (int)B, (int)T,
                  (int)H, (int)hs, (int)items);
    TA_CHECK_LAUNCH();
  }
  const int DP = hs <= 16 ? 16 : (hs <= 32 ? 32 : 64);
  if (DP == 16) return launch_attn_bwd<16>(maps, p, st);
  if (DP == 32) return launch_attn_bwd<32>(maps, p, st);
  return launch_attn_bwd<64>(maps, p, st);
}
