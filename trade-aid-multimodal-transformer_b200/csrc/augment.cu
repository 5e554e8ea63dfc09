// K7b — the training-set augmentation of get_batch on the device-resident token store, bit-exact to the reference's
// host loop (/root/reference/data_utils.py:293-358, called from training_utils.py:353-360 with all_modality_params[r][2]):
//
//     rand_list = [0, 1, -1, ..., s, -s]
//     for n in range(len(data)):                       # index order
//         if s < data[n] < vocab_size - s:             # eligible
//             data[n] += random.choice(rand_list)      # CPython: rand_list[_randbelow(2s+1)]
//
// CPython's _randbelow(n) draws k = n.bit_length() bits (getrandbits(k) = next 32-bit Mersenne-Twister word >> (32 - k))
// and REJECTS values >= n, so an eligible element consumes one or more words and an ineligible one consumes none. The
// host pre-draws the word stream in bulk from the same generator (random.getrandbits(32 * W): its 32-bit limbs are the
// successive words); here the j-th ACCEPTED word is matched with the j-th ELIGIBLE element through two prefix sums:
//   aug_count    per 2048-item tile: accepted words / eligible elements
//   aug_scan     exclusive scan of the tile counts (one block), totals E and A
//   aug_compact  accepted word of rank j -> draws[j]; the word of rank E-1 defines `consumed` (the host re-synchronises
//                Python's generator to exactly that many words)
//   aug_apply    eligible element of rank j: token += delta(draws[j])      (nothing is touched if A < E: too few words)
// All integer work, HBM-bound: 4 B/word + 4 B/token read, 4 B/token written back, 1 B/draw.
#include "../../include/tradeaid_b200.h"
#include "common.cuh"

namespace ta {

constexpr int AUG_THREADS = 256;
constexpr int AUG_PER = 8;
constexpr int AUG_TILE = AUG_THREADS * AUG_PER;

struct AugParams {
  int* tokens;
  long long N;
  const unsigned* words;
  long long W;
  int s, V, n, shift;          // rand_size, vocab size, 2s+1, 32 - bit_length(n)
  unsigned* counts_w;          // [tiles_w] then exclusive offsets
  unsigned* counts_e;          // [tiles_e]
  unsigned char* draws;        // [N]
  long long* stats;            // [0] eligible E, [1] accepted A, [2] consumed words
  int tiles_w, tiles_e;
};

__device__ __forceinline__ bool aug_accept(const AugParams& p, long long i) { return i < p.W && (p.words[i] >> p.shift) < (unsigned)p.n; }
__device__ __forceinline__ bool aug_eligible(const AugParams& p, long long i) {
  if (i >= p.N) return false;
  const int v = p.tokens[i];
  return p.s < v && v < p.V - p.s;
}

// block-wide exclusive prefix of one value per thread (AUG_THREADS threads); returns the exclusive prefix, total in `total`
__device__ __forceinline__ unsigned block_exclusive(unsigned v, unsigned& total) {
  __shared__ unsigned warp_sums[AUG_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();                       // warp_sums may still be read by the previous call
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  unsigned base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < AUG_THREADS / 32; ++w) {
    if (w < warp) base += warp_sums[w];
    tot += warp_sums[w];
  }
  total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(AUG_THREADS) aug_count_kernel(const AugParams p) {
  pdl_prologue();
  const bool is_w = (int)blockIdx.x < p.tiles_w;
  const long long tile = is_w ? blockIdx.x : blockIdx.x - p.tiles_w;
  const long long i0 = tile * AUG_TILE + (long long)threadIdx.x * AUG_PER;
  unsigned c = 0;
#pragma unroll
  for (int j = 0; j < AUG_PER; ++j) c += is_w ? aug_accept(p, i0 + j) : aug_eligible(p, i0 + j);
  unsigned total;
  block_exclusive(c, total);
  if (threadIdx.x == 0) (is_w ? p.counts_w : p.counts_e)[tile] = total;
}

__global__ void __launch_bounds__(1024) aug_scan_kernel(const AugParams p) {
  pdl_prologue();
  // one block: exclusive scan of counts_w and counts_e in place, totals to stats
  __shared__ unsigned long long carry_s;
  __shared__ unsigned warp_sums[32];
  for (int which = 0; which < 2; ++which) {
    unsigned* a = which == 0 ? p.counts_e : p.counts_w;
    const int len = which == 0 ? p.tiles_e : p.tiles_w;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < len; c0 += 1024) {
      const int i = c0 + threadIdx.x;
      const unsigned v = i < len ? a[i] : 0u;
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      unsigned inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) warp_sums[warp] = inc;
      __syncthreads();
      unsigned base = 0, tot = 0;
      for (int w = 0; w < 32; ++w) {
        if (w < warp) base += warp_sums[w];
        tot += warp_sums[w];
      }
      const unsigned long long carry = carry_s;
      // tile offsets stay below 2^32: the binding limits N and W to < 2^31 items
      if (i < len) a[i] = (unsigned)(carry + base + inc - v);
      __syncthreads();
      if (threadIdx.x == 0) carry_s = carry + tot;
      __syncthreads();
    }
    if (threadIdx.x == 0) p.stats[which == 0 ? 0 : 1] = (long long)carry_s;
    __syncthreads();
  }
  if (threadIdx.x == 0) p.stats[2] = 0;      // E == 0: nothing consumed; otherwise aug_compact overwrites it
}

__global__ void __launch_bounds__(AUG_THREADS) aug_compact_kernel(const AugParams p) {
  pdl_prologue();
  const long long E = p.stats[0];
  const long long i0 = (long long)blockIdx.x * AUG_TILE + (long long)threadIdx.x * AUG_PER;
  unsigned flags = 0, c = 0;
#pragma unroll
  for (int j = 0; j < AUG_PER; ++j)
    if (aug_accept(p, i0 + j)) { flags |= 1u << j; ++c; }
  unsigned total;
  long long rank = (long long)p.counts_w[blockIdx.x] + block_exclusive(c, total);
#pragma unroll
  for (int j = 0; j < AUG_PER; ++j) {
    if (flags & (1u << j)) {
      if (rank < E) {
        p.draws[rank] = (unsigned char)(p.words[i0 + j] >> p.shift);
        if (rank == E - 1) p.stats[2] = i0 + j + 1;
      }
      ++rank;
    }
  }
}

__global__ void __launch_bounds__(AUG_THREADS) aug_apply_kernel(const AugParams p) {
  pdl_prologue();
  if (p.stats[1] < p.stats[0]) return;        // too few accepted words for the eligible elements: leave the tokens alone
  const long long i0 = (long long)blockIdx.x * AUG_TILE + (long long)threadIdx.x * AUG_PER;
  unsigned flags = 0, c = 0;
  int v[AUG_PER];
#pragma unroll
  for (int j = 0; j < AUG_PER; ++j) {
    v[j] = i0 + j < p.N ? p.tokens[i0 + j] : 0;
    if (i0 + j < p.N && p.s < v[j] && v[j] < p.V - p.s) { flags |= 1u << j; ++c; }
  }
  unsigned total;
  long long rank = (long long)p.counts_e[blockIdx.x] + block_exclusive(c, total);
#pragma unroll
  for (int j = 0; j < AUG_PER; ++j) {
    if (flags & (1u << j)) {
      const int r = p.draws[rank];             // index into [0, 1, -1, 2, -2, 3, -3]
      p.tokens[i0 + j] = v[j] + ((r & 1) ? (r + 1) / 2 : -(r / 2));
      ++rank;
    }
  }
}

}  // namespace ta

extern "C" {
int64_t ta_augment_workspace_bytes(int64_t N, int64_t W) {
  using namespace ta;
  const int64_t tw = ceil_div(W > 0 ? W : 1, AUG_TILE), te = ceil_div(N > 0 ? N : 1, AUG_TILE);
  return (tw + te) * 4 + ((N + 15) / 16) * 16 + 64;
}

int32_t ta_augment_tokens(int32_t* tokens, int64_t N, const uint32_t* words, int64_t W, int32_t rand_size, int32_t vocab_size,
                          void* workspace, int64_t* stats, void* stream) {
  using namespace ta;
  TA_CHECK_ARG(tokens && words && workspace && stats, "ta_augment_tokens: null pointer");
  TA_CHECK_ARG(rand_size >= 1 && rand_size <= 3, "rand_size must be an integer between 1 and 3, or null.");
  TA_CHECK_ARG(vocab_size > 0, "vocab_size must be a positive integer.");
  TA_CHECK_ARG(N > 0 && W > 0 && N < (1ll << 31) && W < (1ll << 31), "ta_augment_tokens: sizes must be in [1, 2^31)");
  AugParams p{};
  p.tokens = tokens; p.N = N; p.words = words; p.W = W;
  p.s = rand_size; p.V = vocab_size; p.n = 2 * rand_size + 1;
  p.shift = 32 - (rand_size == 1 ? 2 : 3);       // bit_length(3) = 2, bit_length(5) = bit_length(7) = 3
  p.tiles_w = (int)ceil_div(W, AUG_TILE); p.tiles_e = (int)ceil_div(N, AUG_TILE);
  p.counts_w = static_cast<unsigned*>(workspace);
  p.counts_e = p.counts_w + p.tiles_w;
  p.draws = reinterpret_cast<unsigned char*>(p.counts_e + p.tiles_e);
  p.stats = reinterpret_cast<long long*>(stats);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  launch_kernel(aug_count_kernel, dim3(p.tiles_w + p.tiles_e), dim3(AUG_THREADS), 0, st, p);
  TA_CHECK_LAUNCH();
  launch_kernel(aug_scan_kernel, dim3(1), dim3(1024), 0, st, p);
  TA_CHECK_LAUNCH();
  launch_kernel(aug_compact_kernel, dim3(p.tiles_w), dim3(AUG_THREADS), 0, st, p);
  TA_CHECK_LAUNCH();
  launch_kernel(aug_apply_kernel, dim3(p.tiles_e), dim3(AUG_THREADS), 0, st, p);
  TA_CHECK_LAUNCH();
  return 0;
}
}
