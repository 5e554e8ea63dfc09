// K9 — one-shot data preparation on the device: a float64 column resident in HBM goes through the
// reference's processing steps and comes out as int32 token ids in a row of the token store, without a Python loop over
// the elements:
//   convert_to_percent_changes  /root/reference/data_utils.py:612-662, /root/reference/file_cache.py:329-385
//   range_numeric_data          /root/reference/data_utils.py:361-470
//   bin_numeric_data            /root/reference/data_utils.py:473-609  (np.percentile -> radix select of 4 order statistics)
//   numerical_representation    /root/reference/data_utils.py:212-225  (sorted(set(.)) -> hash set + rank by counting)
// All of it is HBM-bound integer / fp64 work: grid-stride kernels with 148 x k CTAs, one 8-byte load and one 8- or 4-byte
// store per element per step (algorithmic bytes per row: percent 16, range 16, select 8 x 8 passes, min/max 8, assign 16,
// vocabulary insert 8, lookup 8 + 4). The element arithmetic (exact CPython round(), libm-compatible decade, exact
// products) lives in dataprep_math.cuh and is shared with the CPU test harness.
// Error reporting: the kernels never trap; they record the FIRST offending row in a device status array that the host
// wrapper reads once per step and turns into the reference's exception.
#include "../../include/tradeaid_b200.h"
#include "common.cuh"
#include "dataprep_math.cuh"

namespace ta {
using namespace prep;

constexpr int DP_THREADS = 256;
constexpr long long NO_ROW = 0x7fffffffffffffffll;

// Grid of a grid-stride kernel: every CTA resident at once (SMs x the kernel's own occupancy), so that there is no
// partial second wave - a fixed 8 CTAs per SM left the 36-42 register kernels with a quarter-filled tail wave (ncu:
// 55 % achieved warps) - and never more CTAs than the rows need (`rows_per_thread` rows per thread).
template <typename Kernel>
static int dp_grid(Kernel kernel, int64_t n, int rows_per_thread = 2, size_t dyn_smem = 0) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, DP_THREADS, dyn_smem) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = 4;
  }
  const int64_t want = ceil_div(n, (int64_t)DP_THREADS * rows_per_thread);
  const int64_t cap = (int64_t)num_sms() * per_sm;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}
static inline int dp_parts(int64_t n) {   // partial results of the two-stage min/max: sized without a device query
  const int64_t want = ceil_div(n, (int64_t)DP_THREADS * 2), cap = (int64_t)num_sms() * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

__device__ __forceinline__ void note_row(long long* slot, long long row) {
  atomicMin(reinterpret_cast<long long*>(slot), row);
}

static Pow10 host_pow10(int nd) { return make_pow10(nd); }

// ------------------------------------------------------------------------------------------------ status
__global__ void prep_status_reset_kernel(long long* status) {
  pdl_prologue();
  if (threadIdx.x < TA_PREP_STATUS_WORDS) status[threadIdx.x] = NO_ROW;
}

// ------------------------------------------------------------------------------------------------ percent changes
// file_starts: ascending row index of the first row of every file (file_starts[0] = 0). A row that starts a file is 0.0
// and the last row of the previous file is never a "previous value" (the reference converts file by file).
__device__ __forceinline__ bool starts_a_file(long long i, const long long* __restrict__ file_starts, int F) {
  if (F == 1) return i == 0;
  int lo = 0, hi = F;   // first index with file_starts[k] >= i
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (file_starts[mid] < i) lo = mid + 1; else hi = mid;
  }
  return lo < F && file_starts[lo] == i;
}

__device__ __forceinline__ double percent_row(long long i, double cur, double prev, bool first, Pow10 P,
                                              long long* __restrict__ status) {
  if (!is_finite(cur)) note_row(status + TA_PREP_STATUS_NONFINITE, i);
  if (first) return 0.0;
  if (prev == 0.0) { note_row(status + TA_PREP_STATUS_ZERO_PREV, i - 1); return 0.0; }
  return percent_change(cur, prev, P);
}

// Every thread takes two rows a grid stride apart per iteration: four independent 8-byte loads in flight.
__global__ void __launch_bounds__(DP_THREADS, 8) prep_percent_kernel(const double* __restrict__ x, double* __restrict__ out,
                                                                    long long N, const long long* __restrict__ file_starts,
                                                                    int F, Pow10 P, long long* __restrict__ status) {
  pdl_prologue();
  const long long stride = (long long)gridDim.x * DP_THREADS;
  for (long long i0 = blockIdx.x * (long long)DP_THREADS + threadIdx.x; i0 < N; i0 += 2 * stride) {
    const long long i1 = i0 + stride;
    const bool two = i1 < N;
    const double c0 = x[i0], p0 = i0 > 0 ? x[i0 - 1] : 0.0;
    const double c1 = two ? x[i1] : 0.0, p1 = two ? x[i1 - 1] : 0.0;
    const bool f0 = starts_a_file(i0, file_starts, F);
    const bool f1 = two && starts_a_file(i1, file_starts, F);
    out[i0] = percent_row(i0, c0, p0, f0, P, status);
    if (two) out[i1] = percent_row(i1, c1, p1, f1, P, status);
  }
}

// ------------------------------------------------------------------------------------------------ ranging
__device__ __forceinline__ double range_row(long long i, double v, const RangeParams& rp, const double* s_thr,
                                            const double* s_fac, long long* __restrict__ status) {
  if (!is_finite(v)) { note_row(status + TA_PREP_STATUS_NONFINITE, i); return v; }
  const double r = range_element(v, rp, s_thr, s_fac);
  if (!is_finite(r)) note_row(status + TA_PREP_STATUS_OVERFLOW, i);   // factor table entry was +inf (10**k > DBL_MAX)
  return r;
}

__global__ void __launch_bounds__(DP_THREADS, 8) prep_range_kernel(const double* __restrict__ x, double* __restrict__ out,
                                                                  long long N, RangeParams rp,
                                                                  const double* __restrict__ thresholds,
                                                                  const double* __restrict__ factors,
                                                                  long long* __restrict__ status) {
  pdl_prologue();
  __shared__ double s_thr[P10_COUNT];
  __shared__ double s_fac[P10_COUNT];
  if (rp.has_whole_digits)
    for (int k = threadIdx.x; k < P10_COUNT; k += DP_THREADS) { s_thr[k] = thresholds[k]; s_fac[k] = factors[k]; }
  __syncthreads();
  const long long stride = (long long)gridDim.x * DP_THREADS;
  for (long long i0 = blockIdx.x * (long long)DP_THREADS + threadIdx.x; i0 < N; i0 += 2 * stride) {
    const long long i1 = i0 + stride;
    const bool two = i1 < N;
    const double v0 = x[i0];
    const double v1 = two ? x[i1] : 0.0;
    out[i0] = range_row(i0, v0, rp, s_thr, s_fac, status);
    if (two) out[i1] = range_row(i1, v1, rp, s_thr, s_fac, status);
  }
}

// decimal_places = None: per-element decimals from the element's own shortest repr (dataprep_math.cuh). The parameter block
// (23 powers of ten + clamps, ~1 KB) is read from global memory into shared memory once per CTA.
__global__ void __launch_bounds__(DP_THREADS, 4) prep_range_auto_kernel(const double* __restrict__ x,
                                                                       const unsigned char* __restrict__ is_int,
                                                                       double* __restrict__ out, long long N,
                                                                       const RangeAutoParams* __restrict__ rp_g,
                                                                       const double* __restrict__ thresholds,
                                                                       const double* __restrict__ factors,
                                                                       long long* __restrict__ status,
                                                                       long long* __restrict__ n_flagged) {
  pdl_prologue();
  __shared__ double s_thr[P10_COUNT];
  __shared__ double s_fac[P10_COUNT];
  __shared__ RangeAutoParams s_rp;
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(rp_g);
    unsigned* dst = reinterpret_cast<unsigned*>(&s_rp);
    for (int k = threadIdx.x; k < (int)(sizeof(RangeAutoParams) / 4); k += DP_THREADS) dst[k] = src[k];
  }
  __syncthreads();
  if (s_rp.has_whole_digits)
    for (int k = threadIdx.x; k < P10_COUNT; k += DP_THREADS) { s_thr[k] = thresholds[k]; s_fac[k] = factors[k]; }
  __syncthreads();
  long long flagged = 0;
  for (long long i = blockIdx.x * (long long)DP_THREADS + threadIdx.x; i < N; i += (long long)gridDim.x * DP_THREADS) {
    const double v = x[i];
    if (!is_finite(v)) { note_row(status + TA_PREP_STATUS_NONFINITE, i); out[i] = v; continue; }
    int flag = 0;
    const double r = range_element_auto(v, is_int && is_int[i], s_rp, s_thr, s_fac, &flag);
    if (!flag && !is_finite(r)) note_row(status + TA_PREP_STATUS_OVERFLOW, i);
    flagged += flag;
    out[i] = r;
  }
  if (flagged) atomicAdd(reinterpret_cast<unsigned long long*>(n_flagged), (unsigned long long)flagged);
}

// ------------------------------------------------------------------------------------------------ order statistics
// Radix select, most significant byte first, for up to TA_PREP_MAX_RANKS ranks at once. state: per rank {prefix, k}.
// Ranks whose prefixes agree on the bytes decided so far see the same rows, so they share one histogram (the neighbouring
// order statistics a percentile interpolates between stay together until the last bytes): leader[r] = the first rank
// with rank r's prefix; only leaders count.
__device__ __forceinline__ void select_leaders(const u64* prefix, int R, int shift, int* leader) {
  for (int r = 0; r < R; ++r) {
    leader[r] = r;
    for (int q = 0; q < r; ++q)
      if (shift == 56 || (prefix[q] >> (shift + 8)) == (prefix[r] >> (shift + 8))) { leader[r] = leader[q]; break; }
  }
}

__global__ void __launch_bounds__(DP_THREADS, 8) prep_select_hist_kernel(const double* __restrict__ x, long long N, int R,
                                                                        int shift, const u64* __restrict__ state,
                                                                        u64* __restrict__ hist,
                                                                        long long* __restrict__ status) {
  pdl_prologue();
  __shared__ unsigned s_h[TA_PREP_MAX_RANKS][256];
  __shared__ u64 s_prefix[TA_PREP_MAX_RANKS];
  __shared__ int s_group[TA_PREP_MAX_RANKS];
  __shared__ int s_groups;
  for (int k = threadIdx.x; k < TA_PREP_MAX_RANKS * 256; k += DP_THREADS) (&s_h[0][0])[k] = 0u;
  if (threadIdx.x == 0) {
    int leader[TA_PREP_MAX_RANKS];
    for (int r = 0; r < R; ++r) s_prefix[r] = state[2 * r];
    select_leaders(s_prefix, R, shift, leader);
    int n = 0;
    for (int r = 0; r < R; ++r)
      if (leader[r] == r) s_group[n++] = r;
    s_groups = n;
  }
  __syncthreads();
  const int groups = s_groups;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * DP_THREADS;
  // warp-uniform trip count so that __match_any_sync sees the whole warp; two rows a stride apart per iteration
  for (long long base = blockIdx.x * (long long)DP_THREADS + threadIdx.x - lane; base < N; base += 2 * stride) {
    const long long i0 = base + lane, i1 = i0 + stride;
    const bool valid0 = i0 < N, valid1 = i1 < N;
    const double v0 = valid0 ? x[i0] : 0.0;
    const double v1 = valid1 ? x[i1] : 0.0;
    if (shift == 56) {   // the first pass sees every row: report non-finite ones
      if (valid0 && !is_finite(v0)) note_row(status + TA_PREP_STATUS_NONFINITE, i0);
      if (valid1 && !is_finite(v1)) note_row(status + TA_PREP_STATUS_NONFINITE, i1);
    }
    const u64 key0 = order_key(v0), key1 = order_key(v1);
    const unsigned byte0 = (unsigned)(key0 >> shift) & 255u, byte1 = (unsigned)(key1 >> shift) & 255u;
    for (int g = 0; g < groups; ++g) {
      const int r = s_group[g];
      const u64 want = shift == 56 ? 0 : s_prefix[r] >> (shift + 8);
      const bool in0 = valid0 && (shift == 56 || (key0 >> (shift + 8)) == want);
      const bool in1 = valid1 && (shift == 56 || (key1 >> (shift + 8)) == want);
      const unsigned peers0 = __match_any_sync(0xffffffffu, in0 ? byte0 : 256u);
      const unsigned peers1 = __match_any_sync(0xffffffffu, in1 ? byte1 : 256u);
      if (in0 && lane == __ffs(peers0) - 1) atomicAdd(&s_h[r][byte0], (unsigned)__popc(peers0));
      if (in1 && lane == __ffs(peers1) - 1) atomicAdd(&s_h[r][byte1], (unsigned)__popc(peers1));
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < R * 256; k += DP_THREADS) {
    const unsigned c = (&s_h[0][0])[k];
    if (c) atomicAdd(&hist[k], (u64)c);
  }
}

// One CTA of 256 threads = one thread per bin: block-wide prefix sum of the leader's histogram, the bin that holds the
// rank extends the prefix. (A serial walk over the 256 bins cost 19 us per pass, a quarter of a pass over 20 M rows.)
__global__ void __launch_bounds__(256) prep_select_pick_kernel(int R, int shift, u64* __restrict__ state,
                                                               u64* __restrict__ hist, double* __restrict__ out) {
  pdl_prologue();
  __shared__ u64 s_prefix[TA_PREP_MAX_RANKS], s_k[TA_PREP_MAX_RANKS];
  __shared__ int s_leader[TA_PREP_MAX_RANKS];
  __shared__ u64 s_c[256];
  const int t = threadIdx.x;
  if (t == 0) {
    for (int r = 0; r < R; ++r) { s_prefix[r] = state[2 * r]; s_k[r] = state[2 * r + 1]; }
    select_leaders(s_prefix, R, shift, s_leader);
  }
  __syncthreads();
  for (int r = 0; r < R; ++r) {
    const u64 c = hist[s_leader[r] * 256 + t];
    s_c[t] = c;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
      const u64 v = t >= off ? s_c[t - off] : 0;
      __syncthreads();
      s_c[t] += v;
      __syncthreads();
    }
    const u64 incl = s_c[t], excl = incl - c, k = s_k[r];
    if ((excl <= k && k < incl) || (t == 255 && k >= incl)) {
      const u64 prefix = s_prefix[r] | ((u64)t << shift);
      state[2 * r] = prefix;
      state[2 * r + 1] = k - excl;
      if (shift == 0) out[r] = order_value(prefix);
    }
    __syncthreads();
  }
  for (int k = t; k < R * 256; k += 256) hist[k] = 0;
}

__global__ void prep_select_init_kernel(int R, const long long* __restrict__ ranks, u64* __restrict__ state,
                                        u64* __restrict__ hist) {
  pdl_prologue();
  if (threadIdx.x < R) { state[2 * threadIdx.x] = 0; state[2 * threadIdx.x + 1] = (u64)ranks[threadIdx.x]; }
  for (int k = threadIdx.x; k < TA_PREP_MAX_RANKS * 256; k += blockDim.x) hist[k] = 0;
}

// ------------------------------------------------------------------------------------------------ filtered min / max
// min and max of the rows with lo <= x <= hi (bin_numeric_data's outlier filter, data_utils.py:503-509) + their count.
__global__ void __launch_bounds__(DP_THREADS) prep_minmax_partial_kernel(const double* __restrict__ x, long long N,
                                                                           double lo, double hi, double* __restrict__ part_min,
                                                                           double* __restrict__ part_max,
                                                                           long long* __restrict__ part_cnt) {
  pdl_prologue();
  double mn = INFINITY, mx = -INFINITY;
  long long cnt = 0;
  for (long long i = blockIdx.x * (long long)DP_THREADS + threadIdx.x; i < N; i += (long long)gridDim.x * DP_THREADS) {
    const double v = x[i];
    if (v >= lo && v <= hi) { mn = fmin(mn, v); mx = fmax(mx, v); ++cnt; }
  }
  __shared__ double s_mn[DP_THREADS], s_mx[DP_THREADS];
  __shared__ long long s_c[DP_THREADS];
  s_mn[threadIdx.x] = mn; s_mx[threadIdx.x] = mx; s_c[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = DP_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_mn[threadIdx.x] = fmin(s_mn[threadIdx.x], s_mn[threadIdx.x + o]);
      s_mx[threadIdx.x] = fmax(s_mx[threadIdx.x], s_mx[threadIdx.x + o]);
      s_c[threadIdx.x] += s_c[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part_min[blockIdx.x] = s_mn[0]; part_max[blockIdx.x] = s_mx[0]; part_cnt[blockIdx.x] = s_c[0]; }
}

__global__ void __launch_bounds__(DP_THREADS) prep_minmax_final_kernel(const double* __restrict__ part_min,
                                                                         const double* __restrict__ part_max,
                                                                         const long long* __restrict__ part_cnt, int parts,
                                                                         double* __restrict__ out) {
  pdl_prologue();
  double mn = INFINITY, mx = -INFINITY;
  long long cnt = 0;
  for (int k = threadIdx.x; k < parts; k += DP_THREADS) { mn = fmin(mn, part_min[k]); mx = fmax(mx, part_max[k]); cnt += part_cnt[k]; }
  __shared__ double s_mn[DP_THREADS], s_mx[DP_THREADS];
  __shared__ long long s_c[DP_THREADS];
  s_mn[threadIdx.x] = mn; s_mx[threadIdx.x] = mx; s_c[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = DP_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_mn[threadIdx.x] = fmin(s_mn[threadIdx.x], s_mn[threadIdx.x + o]);
      s_mx[threadIdx.x] = fmax(s_mx[threadIdx.x], s_mx[threadIdx.x + o]);
      s_c[threadIdx.x] += s_c[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = s_mn[0]; out[1] = s_mx[0]; out[2] = (double)s_c[0]; }
}

// ------------------------------------------------------------------------------------------------ bin assignment
__global__ void __launch_bounds__(DP_THREADS) prep_bin_kernel(const double* __restrict__ x, double* __restrict__ out,
                                                                long long N, const double* __restrict__ pos,
                                                                const double* __restrict__ neg, int G, int sorted) {
  pdl_prologue();
  extern __shared__ double s_bounds[];
  double* s_pos = s_bounds;
  double* s_neg = s_bounds + (G + 1);
  for (int k = threadIdx.x; k <= G; k += DP_THREADS) { s_pos[k] = pos[k]; s_neg[k] = neg[k]; }
  __syncthreads();
  const long long stride = (long long)gridDim.x * DP_THREADS;
  for (long long i0 = blockIdx.x * (long long)DP_THREADS + threadIdx.x; i0 < N; i0 += 2 * stride) {
    const long long i1 = i0 + stride;
    const bool two = i1 < N;
    const double v0 = x[i0];
    const double v1 = two ? x[i1] : 0.0;
    out[i0] = (double)bin_of(v0, s_pos, s_neg, G, sorted != 0);
    if (two) out[i1] = (double)bin_of(v1, s_pos, s_neg, G, sorted != 0);
  }
}

// ------------------------------------------------------------------------------------------------ vocabulary
constexpr u64 SLOT_EMPTY = 0xffffffffffffffffull;   // a NaN pattern: never a key (non-finite rows are reported)
constexpr long long VOCAB_MAX_PROBES = 4096;        // linear probing at <= 50 % load never gets near this

__global__ void __launch_bounds__(DP_THREADS) prep_vocab_insert_kernel(const double* __restrict__ x, long long N,
                                                                         u64* __restrict__ table, long long cap_mask,
                                                                         long long* __restrict__ counters,
                                                                         long long* __restrict__ status) {
  pdl_prologue();
  long long added = 0;
  const long long probe_limit = cap_mask + 1 < VOCAB_MAX_PROBES ? cap_mask + 1 : VOCAB_MAX_PROBES;
  for (long long i = blockIdx.x * (long long)DP_THREADS + threadIdx.x; i < N; i += (long long)gridDim.x * DP_THREADS) {
    const double v = x[i];
    if (!is_finite(v)) { note_row(status + TA_PREP_STATUS_NONFINITE, i); continue; }
    const u64 key = vocab_key(v);
    u64 h = hash64(key) & (u64)cap_mask;
    bool placed = false;
    for (long long probe = 0; probe < probe_limit; ++probe) {
      u64 cur = *reinterpret_cast<volatile u64*>(table + h);
      if (cur == SLOT_EMPTY) {
        cur = atomicCAS(table + h, SLOT_EMPTY, key);
        if (cur == SLOT_EMPTY) { ++added; placed = true; break; }
      }
      if (cur == key) { placed = true; break; }
      h = (h + 1) & (u64)cap_mask;
    }
    if (!placed) {   // table too small for this column: the host retries with a larger one; stop working on this one
      note_row(status + TA_PREP_STATUS_TABLE_FULL, i);
      break;
    }
  }
  if (added) atomicAdd(reinterpret_cast<unsigned long long*>(counters), (unsigned long long)added);
}

__global__ void __launch_bounds__(DP_THREADS) prep_vocab_compact_kernel(const u64* __restrict__ table, long long cap,
                                                                          u64* __restrict__ keys, int* __restrict__ slots,
                                                                          long long* __restrict__ counters) {
  pdl_prologue();
  for (long long h = blockIdx.x * (long long)DP_THREADS + threadIdx.x; h < cap; h += (long long)gridDim.x * DP_THREADS) {
    const u64 k = table[h];
    if (k != SLOT_EMPTY) {
      const long long pos = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(counters + 1), 1ull);
      keys[pos] = k;
      slots[pos] = (int)h;
    }
  }
}

// token of a key = number of distinct keys below it (= its index in sorted(set(data))); vocabulary[token] = value.
__global__ void __launch_bounds__(DP_THREADS) prep_vocab_rank_kernel(const u64* __restrict__ keys,
                                                                       const int* __restrict__ slots, int V,
                                                                       double* __restrict__ vocabulary,
                                                                       int* __restrict__ token_of_slot) {
  pdl_prologue();
  __shared__ double s_v[DP_THREADS];
  const int me = blockIdx.x * DP_THREADS + threadIdx.x;
  const double mine = me < V ? double_of(keys[me]) : 0.0;
  int rank = 0;
  for (int base = 0; base < V; base += DP_THREADS) {
    __syncthreads();
    s_v[threadIdx.x] = base + threadIdx.x < V ? double_of(keys[base + threadIdx.x]) : INFINITY;
    __syncthreads();
    const int n = V - base < DP_THREADS ? V - base : DP_THREADS;
    for (int q = 0; q < n; ++q) rank += s_v[q] < mine;
  }
  if (me < V) {
    vocabulary[rank] = mine;
    token_of_slot[slots[me]] = rank;
  }
}

__global__ void __launch_bounds__(DP_THREADS) prep_vocab_lookup_kernel(const double* __restrict__ x, long long N,
                                                                         const u64* __restrict__ table, long long cap_mask,
                                                                         const int* __restrict__ token_of_slot,
                                                                         int* __restrict__ tokens) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)DP_THREADS + threadIdx.x; i < N; i += (long long)gridDim.x * DP_THREADS) {
    const u64 key = vocab_key(x[i]);
    u64 h = hash64(key) & (u64)cap_mask;
    int tok = -1;
    for (long long probe = 0; probe <= cap_mask && probe < VOCAB_MAX_PROBES; ++probe) {
      const u64 cur = table[h];
      if (cur == key) { tok = token_of_slot[h]; break; }
      if (cur == SLOT_EMPTY) break;
      h = (h + 1) & (u64)cap_mask;
    }
    tokens[i] = tok;
  }
}

}  // namespace ta

using namespace ta;

extern "C" {

int32_t ta_prep_status_reset(int64_t* status, void* stream) {
  TA_CHECK_ARG(status != nullptr, "ta_prep_status_reset: null status");
  TA_CHECK_CUDA(launch_kernel(prep_status_reset_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (long long*)status));
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_prep_percent_changes(const double* x, double* out, int64_t N, const int64_t* file_starts, int32_t n_files,
                                int32_t decimals, int64_t* status, void* stream) {
  TA_CHECK_ARG(x && out && file_starts && status, "ta_prep_percent_changes: null pointer");
  TA_CHECK_ARG(x != out, "ta_prep_percent_changes: not an in-place operation (row i reads row i-1)");
  TA_CHECK_ARG(N > 0 && n_files >= 1, "ta_prep_percent_changes: N=%lld n_files=%d", (long long)N, n_files);
  TA_CHECK_ARG(decimals >= 0 && decimals <= MAX_DECIMALS, "ta_prep_percent_changes: decimals %d outside [0, %d]", decimals,
               MAX_DECIMALS);
  TA_CHECK_CUDA(launch_kernel(prep_percent_kernel, dim3(dp_grid(prep_percent_kernel, N)), dim3(DP_THREADS), 0, (cudaStream_t)stream, x, out,
                              (long long)N, (const long long*)file_starts, (int)n_files, host_pow10(decimals),
                              (long long*)status));
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_prep_range(const double* x, double* out, int64_t N, int32_t has_whole_digits, int32_t decimals, double lower,
                      double upper, double upper_clamp, const double* thresholds, const double* factors, int64_t* status,
                      void* stream) {
  TA_CHECK_ARG(x && out && status, "ta_prep_range: null pointer");
  TA_CHECK_ARG(N > 0, "ta_prep_range: N=%lld", (long long)N);
  TA_CHECK_ARG(decimals >= 0 && decimals <= MAX_DECIMALS, "ta_prep_range: decimals %d outside [0, %d]", decimals, MAX_DECIMALS);
  TA_CHECK_ARG(!has_whole_digits || (thresholds && factors), "ta_prep_range: decade tables required with whole digits");
  RangeParams rp;
  rp.has_whole_digits = has_whole_digits ? 1 : 0;
  rp.decimals = decimals;
  rp.P = host_pow10(decimals);
  rp.lower = lower;
  rp.upper = upper;
  rp.upper_clamp = upper_clamp;
  TA_CHECK_CUDA(launch_kernel(prep_range_kernel, dim3(dp_grid(prep_range_kernel, N)), dim3(DP_THREADS), 0, (cudaStream_t)stream, x, out,
                              (long long)N, rp, thresholds, factors, (long long*)status));
  TA_CHECK_LAUNCH();
  return 0;
}

int64_t ta_prep_range_auto_params_bytes(void) { return (int64_t)sizeof(RangeAutoParams); }

int32_t ta_prep_range_auto_params(void* params_host, int32_t has_whole_digits, double lower, double upper,
                                  const double* upper_clamps_host) {
  TA_CHECK_ARG(params_host && upper_clamps_host, "ta_prep_range_auto_params: null pointer");
  RangeAutoParams* rp = static_cast<RangeAutoParams*>(params_host);
  rp->has_whole_digits = has_whole_digits ? 1 : 0;
  rp->lower = lower;
  rp->upper = upper;
  for (int d = 0; d <= AUTO_MAX_DECIMALS; ++d) {
    rp->P[d] = host_pow10(d);
    rp->upper_clamp[d] = upper_clamps_host[d];
  }
  return 0;
}

int32_t ta_prep_range_auto(const double* x, const uint8_t* is_int, double* out, int64_t N, const void* params_dev,
                           const double* thresholds, const double* factors, int64_t* status, int64_t* n_flagged, void* stream) {
  TA_CHECK_ARG(x && out && params_dev && status && n_flagged, "ta_prep_range_auto: null pointer");
  TA_CHECK_ARG(N > 0, "ta_prep_range_auto: N=%lld", (long long)N);
  TA_CHECK_CUDA(launch_kernel(prep_range_auto_kernel, dim3(dp_grid(prep_range_auto_kernel, N)), dim3(DP_THREADS), 0,
                              (cudaStream_t)stream, x, (const unsigned char*)is_int, out, (long long)N,
                              (const RangeAutoParams*)params_dev, thresholds, factors, (long long*)status, (long long*)n_flagged));
  TA_CHECK_LAUNCH();
  return 0;
}

int64_t ta_prep_select_workspace_bytes(void) { return (int64_t)(2 * TA_PREP_MAX_RANKS + TA_PREP_MAX_RANKS * 256) * 8; }

int32_t ta_prep_order_statistics(const double* x, int64_t N, const int64_t* ranks_dev, int32_t n_ranks, double* out,
                                 void* workspace, int64_t* status, void* stream) {
  TA_CHECK_ARG(x && ranks_dev && out && workspace && status, "ta_prep_order_statistics: null pointer");
  TA_CHECK_ARG(N > 0 && n_ranks >= 1 && n_ranks <= TA_PREP_MAX_RANKS, "ta_prep_order_statistics: N=%lld n_ranks=%d",
               (long long)N, n_ranks);
  u64* state = (u64*)workspace;
  u64* hist = state + 2 * TA_PREP_MAX_RANKS;
  cudaStream_t st = (cudaStream_t)stream;
  TA_CHECK_CUDA(launch_kernel(prep_select_init_kernel, dim3(1), dim3(256), 0, st, (int)n_ranks, (const long long*)ranks_dev,
                              state, hist));
  TA_CHECK_LAUNCH();
  const int grid = dp_grid(prep_select_hist_kernel, N);
  for (int shift = 56; shift >= 0; shift -= 8) {
    TA_CHECK_CUDA(launch_kernel(prep_select_hist_kernel, dim3(grid), dim3(DP_THREADS), 0, st, x, (long long)N, (int)n_ranks,
                                shift, (const u64*)state, hist, (long long*)status));
    TA_CHECK_LAUNCH();
    TA_CHECK_CUDA(launch_kernel(prep_select_pick_kernel, dim3(1), dim3(256), 0, st, (int)n_ranks, shift, state, hist, out));
    TA_CHECK_LAUNCH();
  }
  return 0;
}

int64_t ta_prep_minmax_workspace_bytes(int64_t N) { return (int64_t)dp_parts(N) * 24; }

int32_t ta_prep_filtered_minmax(const double* x, int64_t N, double lo, double hi, double* out3, void* workspace,
                                void* stream) {
  TA_CHECK_ARG(x && out3 && workspace, "ta_prep_filtered_minmax: null pointer");
  TA_CHECK_ARG(N > 0, "ta_prep_filtered_minmax: N=%lld", (long long)N);
  const int grid = dp_parts(N);
  double* pmin = (double*)workspace;
  double* pmax = pmin + grid;
  long long* pcnt = (long long*)(pmax + grid);
  cudaStream_t st = (cudaStream_t)stream;
  TA_CHECK_CUDA(launch_kernel(prep_minmax_partial_kernel, dim3(grid), dim3(DP_THREADS), 0, st, x, (long long)N, lo, hi, pmin,
                              pmax, pcnt));
  TA_CHECK_LAUNCH();
  TA_CHECK_CUDA(launch_kernel(prep_minmax_final_kernel, dim3(1), dim3(DP_THREADS), 0, st, (const double*)pmin,
                              (const double*)pmax, (const long long*)pcnt, grid, out3));
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_prep_bin_assign(const double* x, double* out, int64_t N, const double* pos_bounds, const double* neg_bounds,
                           int32_t num_groups, int32_t bounds_sorted, void* stream) {
  TA_CHECK_ARG(x && out && pos_bounds && neg_bounds, "ta_prep_bin_assign: null pointer");
  TA_CHECK_ARG(N > 0, "ta_prep_bin_assign: N=%lld", (long long)N);
  TA_CHECK_ARG(num_groups >= 1 && num_groups <= TA_PREP_MAX_GROUPS, "ta_prep_bin_assign: num_groups %d outside [1, %d]",
               num_groups, TA_PREP_MAX_GROUPS);
  const size_t smem = (size_t)2 * (num_groups + 1) * sizeof(double);
  if (smem > 48 * 1024)
    TA_CHECK_CUDA(cudaFuncSetAttribute(prep_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TA_CHECK_CUDA(launch_kernel(prep_bin_kernel, dim3(dp_grid(prep_bin_kernel, N, 2, smem)), dim3(DP_THREADS), smem, (cudaStream_t)stream, x, out,
                              (long long)N, pos_bounds, neg_bounds, (int)num_groups, (int)(bounds_sorted != 0)));
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_prep_vocab_insert(const double* x, int64_t N, uint64_t* table, int64_t capacity, int64_t* counters,
                             int64_t* status, void* stream) {
  TA_CHECK_ARG(x && table && counters && status, "ta_prep_vocab_insert: null pointer");
  TA_CHECK_ARG(N > 0, "ta_prep_vocab_insert: N=%lld", (long long)N);
  TA_CHECK_ARG(capacity >= 2 && capacity <= (1ll << 31) && (capacity & (capacity - 1)) == 0,
               "ta_prep_vocab_insert: capacity %lld is not a power of two in [2, 2^31]", (long long)capacity);
  cudaStream_t st = (cudaStream_t)stream;
  TA_CHECK_CUDA(cudaMemsetAsync(table, 0xff, (size_t)capacity * 8, st));
  TA_CHECK_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(int64_t), st));
  TA_CHECK_CUDA(launch_kernel(prep_vocab_insert_kernel, dim3(dp_grid(prep_vocab_insert_kernel, N, 1)), dim3(DP_THREADS), 0, st, x, (long long)N,
                              (u64*)table, (long long)(capacity - 1), (long long*)counters, (long long*)status));
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_prep_vocab_build(const uint64_t* table, int64_t capacity, int64_t n_distinct, uint64_t* keys_ws, int32_t* slots_ws,
                            double* vocabulary, int32_t* token_of_slot, int64_t* counters, void* stream) {
  TA_CHECK_ARG(table && keys_ws && slots_ws && vocabulary && token_of_slot && counters, "ta_prep_vocab_build: null pointer");
  TA_CHECK_ARG(n_distinct >= 1 && n_distinct <= capacity && n_distinct < (1ll << 31), "ta_prep_vocab_build: n_distinct=%lld",
               (long long)n_distinct);
  cudaStream_t st = (cudaStream_t)stream;
  TA_CHECK_CUDA(launch_kernel(prep_vocab_compact_kernel, dim3(dp_grid(prep_vocab_compact_kernel, capacity, 1)), dim3(DP_THREADS), 0, st, (const u64*)table,
                              (long long)capacity, (u64*)keys_ws, (int*)slots_ws, (long long*)counters));
  TA_CHECK_LAUNCH();
  TA_CHECK_CUDA(launch_kernel(prep_vocab_rank_kernel, dim3((unsigned)ceil_div(n_distinct, DP_THREADS)), dim3(DP_THREADS), 0, st,
                              (const u64*)keys_ws, (const int*)slots_ws, (int)n_distinct, vocabulary, (int*)token_of_slot));
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_prep_vocab_lookup(const double* x, int64_t N, const uint64_t* table, int64_t capacity, const int32_t* token_of_slot,
                             int32_t* tokens, void* stream) {
  TA_CHECK_ARG(x && table && token_of_slot && tokens, "ta_prep_vocab_lookup: null pointer");
  TA_CHECK_ARG(N > 0 && capacity >= 2 && (capacity & (capacity - 1)) == 0, "ta_prep_vocab_lookup: N=%lld capacity=%lld",
               (long long)N, (long long)capacity);
  TA_CHECK_CUDA(launch_kernel(prep_vocab_lookup_kernel, dim3(dp_grid(prep_vocab_lookup_kernel, N, 1)), dim3(DP_THREADS), 0, (cudaStream_t)stream, x,
                              (long long)N, (const u64*)table, (long long)(capacity - 1), (const int*)token_of_slot,
                              (int*)tokens));
  TA_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
