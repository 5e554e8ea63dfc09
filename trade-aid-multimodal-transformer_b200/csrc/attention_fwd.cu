// K3 forward — causal flash-style attention on tcgen05/TMEM for both the self-attention heads
// (/root/reference/model.py:60-73) and the selective cross-attention heads (model.py:136-159, one independent causal
// softmax per key/value source modality, outputs summed).
//
// Persistent, warp-specialised: one CTA per SM walks (group, batch, head) work items; inside an item the schedule is
// q tile i (128 rows) -> source s -> key/value block j <= i (128 rows): one (i, s) pass is a softmax "segment".
//   warp 16  loader : TMA Q tiles (2 slots) and K/V blocks (3-stage ring); runs ahead across segments and work items
//   warp 17  issuer : S = Q K^T (128x128x hs) into one of two TMEM S buffers, one pair AHEAD of the softmax;
//                     O += P V (128 x hs x 128, V as the MN-major B operand) into one of two TMEM O buffers
//   warps 0-15 math : thread (row r, column quarter) owns 32 S columns: tcgen05.ld, causal mask, online softmax with exp2
//                     (row max exchanged between the four quarters through smem, lazy rescale of O in TMEM only when the
//                     max moves by more than 2^8), P (bf16) into one of two 128B-swizzled smem buffers; the segment
//                     epilogue (O / l, += over sources, lse) is deferred behind the next pair's exponentials so the
//                     last P V of a segment never stalls the math warps.
// The first version ran these phases one after another in a 128-thread CTA (tensor pipe 8% active, warps mostly
// waiting on mbarriers: profiles/r01_ncu_attn_v2_*.csv). Nothing of size T x T ever touches HBM. head_size <= 64.
#include <cmath>
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace ta {

constexpr int ATT_MAX_SRC = 8;
constexpr int AF_MATH_WARPS = 16;
constexpr int AF_THREADS = (AF_MATH_WARPS + 4) * 32;   // warps 0-15 math, 16 loader, 17 issuer, 18-19 idle (whole warpgroup for setmaxnreg)
constexpr int AF_KV_STAGES = 3;
constexpr int AF_TILE = 16384;   // 128 rows x 128 bytes

struct AttnMaps {
  CUtensorMap q;
  CUtensorMap k[ATT_MAX_SRC];
  CUtensorMap v[ATT_MAX_SRC];
};

struct AttnFwdParams {
  int T, B, H, G, hs, n_src, n_qtiles, total_work;
  float scale_log2;
  __nv_bfloat16* o;       // [G][B*T][ldo], head h at column h*hs
  long long ldo, o_gs;
  __nv_bfloat16* o_src;   // optional per-source outputs [n_src][G][B*T][ldo] (needed by backward when n_src > 1)
  long long o_src_stride;
  float* lse;             // [n_src][G][B][H][T], natural log
};

// The (work, i, s, j) schedule; every warp role walks the same sequence with its own copy.
struct AfIter {
  int total, n_qt, n_src, stride;
  int work, i, s, j;
  int n, k, qc;   // running counters over the CTA's lifetime: pairs, segments, q tiles
  __device__ __forceinline__ void init(int total_, int n_qt_, int n_src_, int first, int stride_) {
    total = total_; n_qt = n_qt_; n_src = n_src_; stride = stride_;
    work = first; i = s = j = 0; n = k = qc = 0;
  }
  __device__ __forceinline__ bool valid() const { return work < total; }
  __device__ __forceinline__ bool first_in_seg() const { return j == 0; }
  __device__ __forceinline__ bool last_in_seg() const { return j == i; }
  __device__ __forceinline__ bool new_q() const { return j == 0 && s == 0; }
  __device__ __forceinline__ bool last_of_q() const { return j == i && s == n_src - 1; }
  __device__ __forceinline__ void next() {
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * AF_TILE);
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
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(s_full + 12);
  float* sMax = reinterpret_cast<float*>(s_full + 13);   // [2 parity][4 quarters][128]
  float* sL = sMax + 1024;                                // [2 parity][4 quarters][128]
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int x = 0; x < 2; ++x) {
      mbar_init(&q_full[x], 1); mbar_init(&q_empty[x], 1);
      mbar_init(&s_full[x], 1); mbar_init(&s_empty[x], AF_MATH_WARPS * 32);
      mbar_init(&p_full[x], AF_MATH_WARPS * 32); mbar_init(&p_empty[x], 1);
      mbar_init(&o_full[x], 1); mbar_init(&o_empty[x], AF_MATH_WARPS * 32);
    }
    for (int x = 0; x < NS; ++x) { mbar_init(&kv_full[x], 1); mbar_init(&kv_empty[x], 1); }
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
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,256+DP) O1 [320,320+DP)
  AfIter it;
  it.init(p.total_work, p.n_qtiles, p.n_src, blockIdx.x, gridDim.x);

  if (warp >= AF_MATH_WARPS) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == AF_MATH_WARPS) {
    // ================================================================== loader
    if ((tid & 31) == 0) {
      for (; it.valid(); it.next()) {
        int w = it.work;
        const int h = w % p.H; w /= p.H;
        const int b = w % p.B;
        const int g = w / p.B;
        if (it.new_q()) {
          const int qs = it.qc & 1;
          mbar_wait(&q_empty[qs], ((it.qc >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[qs], AF_TILE);
          tma_load_5d(sQ + qs * AF_TILE, &maps.q, &q_full[qs], 0, h, it.i * 128, b, g);
        }
        const int st = it.n % NS, u = it.n / NS;
        mbar_wait(&kv_empty[st], (u & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * AF_TILE);
        uint8_t* dK = sKV + st * 2 * AF_TILE;
        tma_load_5d(dK, &maps.k[it.s], &kv_full[st], 0, h, it.j * 128, b, g);
        tma_load_5d(dK + AF_TILE, &maps.v[it.s], &kv_full[st], 0, h, it.j * 128, b, g);
      }
    }
  } else if (warp == AF_MATH_WARPS + 1) {
    // ================================================================== UMMA issuer
    if ((tid & 31) == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DP, 0, 1);
      const int nk_qk = (p.hs + 15) / 16;
      auto issue_s = [&](const AfIter& a) {
        const int qs = a.qc & 1, st = a.n % NS, sb = a.n & 1;
        if (a.new_q()) mbar_wait(&q_full[qs], (a.qc >> 1) & 1);
        mbar_wait(&kv_full[st], (a.n / NS) & 1);
        mbar_wait(&s_empty[sb], ((a.n >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t aQ = smem_u32(sQ + qs * AF_TILE), aK = smem_u32(sKV + st * 2 * AF_TILE);
        for (int kk = 0; kk < nk_qk; ++kk)
          umma_bf16(tmem_base + sb * 128, umma_desc_sw128(aQ + kk * 32, 16, 1024), umma_desc_sw128(aK + kk * 32, 16, 1024),
                    idesc_s, kk != 0);
        umma_commit(&s_full[sb]);
      };
      auto issue_pv = [&](const AfIter& a) {
        const int st = a.n % NS, pb = a.n & 1, ob = a.k & 1;
        if (a.first_in_seg()) mbar_wait(&o_empty[ob], ((a.k >> 1) & 1) ^ 1);
        mbar_wait(&p_full[pb], (a.n >> 1) & 1);
        tc_fence_after();
        const uint32_t aP = smem_u32(sP + pb * 2 * AF_TILE), aV = smem_u32(sKV + st * 2 * AF_TILE + AF_TILE);
        // only the k-steps that can hold unmasked columns for some row of the tile
        const int last_row = (a.i * 128 + 127 < p.T - 1) ? a.i * 128 + 127 : p.T - 1;
        int nk = (last_row - a.j * 128) / 16 + 1;
        if (nk > 8) nk = 8;
        for (int kk = 0; kk < nk; ++kk)
          umma_bf16(tmem_base + 256 + ob * 64, umma_desc_sw128(aP + (kk >> 2) * AF_TILE + (kk & 3) * 32, 16, 1024),
                    umma_desc_sw128(aV + kk * 2048, 8192, 1024), idesc_o, !a.first_in_seg() || kk != 0);
        umma_commit(&p_empty[pb]);
        umma_commit(&kv_empty[st]);
        if (a.last_in_seg()) umma_commit(&o_full[ob]);
        if (a.last_of_q()) umma_commit(&q_empty[a.qc & 1]);
      };
      AfIter is = it;   // S runs one pair ahead of P V
      if (is.valid()) { issue_s(is); is.next(); }
      for (; it.valid(); it.next()) {
        if (is.valid()) { issue_s(is); is.next(); }
        issue_pv(it);
      }
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
    float o_tot[MULTI ? EC : 1];
#pragma unroll
    for (int e = 0; e < (MULTI ? EC : 1); ++e) o_tot[e] = 0.f;
    // deferred segment epilogue
    bool pend = false;
    int pe_work = 0, pe_i = 0, pe_s = 0, pe_k = 0;
    bool pe_last_src = false;
    float pe_m = 0.f, pe_l = 0.f;
    int xpar = 0;     // parity of the max-exchange buffer
    auto epilogue = [&]() {
      float* sl = sL + (pe_k & 1) * 512;
      sl[qd * 128 + r] = pe_l;
      named_bar_sync(bar_id, 128);
      const float l = (sl[r] + sl[128 + r]) + (sl[256 + r] + sl[384 + r]);
      const float inv = 1.f / l;
      const int ob = pe_k & 1;
      mbar_wait(&o_full[ob], (pe_k >> 1) & 1);
      tc_fence_after();
      uint32_t ov[EC];
      if (ep_active) {
        tmem_ld_cols<EC>(tmem_base + 256 + ob * 64 + lane_off + qd * EC, ov);
        tmem_wait_ld();
      }
      tc_fence_before();
      mbar_arrive(&o_empty[ob]);
      int w = pe_work;
      const int h = w % p.H; w /= p.H;
      const int b = w % p.B;
      const int g = w / p.B;
      const int row = pe_i * 128 + r;
      const bool row_ok = row < p.T;
      if (ep_active) {
        const long long off = (long long)g * p.o_gs + ((long long)b * p.T + row) * p.ldo + h * p.hs + qd * EC;
        float val[EC];
#pragma unroll
        for (int e = 0; e < EC; ++e) val[e] = __uint_as_float(ov[e]) * inv;
        auto store = [&](__nv_bfloat16* dst, const float (&x)[EC]) {
#pragma unroll
          for (int e = 0; e < EC; e += 8) {
            if (qd * EC + e < p.hs) {
              uint4 u;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int q = 0; q < 4; ++q) h2[q] = __floats2bfloat162_rn(x[e + 2 * q], x[e + 2 * q + 1]);
              *reinterpret_cast<uint4*>(dst + e) = u;
            }
          }
        };
        if constexpr (MULTI) {
          if (p.o_src && row_ok) store(p.o_src + (long long)pe_s * p.o_src_stride + off, val);
#pragma unroll
          for (int e = 0; e < EC; ++e) o_tot[e] += val[e];
          if (pe_last_src) {
#pragma unroll
            for (int e = 0; e < EC; ++e) val[e] = o_tot[e];
            if (row_ok) store(p.o + off, val);
#pragma unroll
            for (int e = 0; e < EC; ++e) o_tot[e] = 0.f;
          }
        } else if (row_ok) {
          store(p.o + off, val);
          if (p.o_src) store(p.o_src + off, val);
        }
      }
      if (qd == 0 && row_ok)
        p.lse[((((long long)pe_s * p.G + g) * p.B + b) * p.H + h) * p.T + row] = pe_m * 0.6931471805599453f + logf(l);
      pend = false;
    };

    for (; it.valid(); it.next()) {
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
      l_run = l_run * alpha + psum;
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
  constexpr int SMEM = (2 + 2 * AF_KV_STAGES + 4) * AF_TILE + 256 + 8192 + 64;
  static bool attr = false;
  if (!attr) {
    TA_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DP, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr = true;
  }
  const int grid = p.total_work < num_sms() ? p.total_work : num_sms();
  attn_fwd_kernel<DP, MULTI><<<grid, AF_THREADS, SMEM, st>>>(maps, p);
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
  AttnMaps maps;
  if (make_head_map(&maps.q, q, hs, H, T, B, G, 128)) return 1;
  for (int s = 0; s < n_src; ++s) {
    if (make_head_map(&maps.k[s], &k[s], hs, H, T, B, G, 128)) return 1;
    if (make_head_map(&maps.v[s], &v[s], hs, H, T, B, G, 128)) return 1;
  }
  AttnFwdParams p{};
  p.T = T; p.B = B; p.H = H; p.G = G; p.hs = hs; p.n_src = n_src;
  p.n_qtiles = (T + 127) / 128;
  p.scale_log2 = (float)(1.4426950408889634 / sqrt((double)hs));
  p.o = (__nv_bfloat16*)o_bf16; p.ldo = ldo; p.o_gs = o_gs;
  p.o_src = (__nv_bfloat16*)o_src_bf16; p.o_src_stride = o_src_stride;
  p.lse = lse;
  const long long total = (long long)G * B * H;
  TA_CHECK_ARG(total < (1ll << 31), "attention grid too large");
  p.total_work = (int)total;
  auto st = (cudaStream_t)stream;
  if (n_src > 1) {
    if (hs <= 16) return launch_attn_fwd<16, true>(maps, p, st);
    if (hs <= 32) return launch_attn_fwd<32, true>(maps, p, st);
    return launch_attn_fwd<64, true>(maps, p, st);
  }
  if (hs <= 16) return launch_attn_fwd<16, false>(maps, p, st);
  if (hs <= 32) return launch_attn_fwd<32, false>(maps, p, st);
  return launch_attn_fwd<64, false>(maps, p, st);
}
