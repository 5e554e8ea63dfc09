// K8 — directional evaluation metrics of one modality on the device (/root/reference/training_utils.py:183-330).
// Per sample of the batch, from the logits of the LAST position: arg-max token -> predicted value, direction sign of
// the prediction and of the actual next value (against the previous input value, or against zero for percentage
// data, training_utils.py:198-211), a win when the signs agree, and the softmax mass on every token whose own direction
// sign equals the predicted one ("certainty"). The reference walks batch x vocabulary in Python with an .item() per
// element; here one warp takes a sample, and the per-sample results are summed in sample order by one thread, so the
// counts are exact and the certainty sum is reproducible. The totals are ADDED to device accumulators: estimate_loss
// runs eval_iters batches per split and reads them back once.
#include "../../include/tradeaid_b200.h"
#include "common.cuh"

namespace ta {

constexpr int DM_THREADS = 256;
constexpr int DM_ROWS = 1024;      // samples per pass through shared memory

__device__ __forceinline__ int sign_of(double x) { return (x > 0.0) - (x < 0.0); }
// ordering in which NaN is the largest value (torch.argmax / torch.max semantics)
__device__ __forceinline__ bool nan_max_gt(float a, float b) { return a > b || (a != a && b == b); }
__device__ __forceinline__ bool nan_max_eq(float a, float b) { return a == b || (a != a && b != b); }
__device__ __forceinline__ long long clamp_token(long long t, int V) { return t < 0 ? 0 : (t >= V ? V - 1 : t); }

__global__ void __launch_bounds__(DM_THREADS) directional_metrics_kernel(
    const float* __restrict__ logits, long long ld, const double* __restrict__ vals, const long long* __restrict__ prev,
    long long prev_stride, const long long* __restrict__ actual, long long actual_stride, int rows, int V, int is_pct,
    long long* __restrict__ counts, double* __restrict__ certainty) {
  pdl_prologue();
  __shared__ int s_win[DM_ROWS];
  __shared__ double s_cert[DM_ROWS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long wins = 0, losses = 0;
  double cert = 0.0;
  for (int r0 = 0; r0 < rows; r0 += DM_ROWS) {
    const int n = rows - r0 < DM_ROWS ? rows - r0 : DM_ROWS;
    for (int r = warp; r < n; r += DM_THREADS / 32) {
      const float* x = logits + (long long)(r0 + r) * ld;
      // arg-max as torch.argmax: lowest index among equal maxima, a NaN counts as the maximum (first NaN wins). The
      // reference survives NaN logits (main.py:606-614 warns and goes on), so a row of NaNs must not index out of range:
      // `arg` is always a valid column, and the token ids read from the batch are clamped into the vocabulary.
      float best = x[0];
      int arg = 0;
      for (int c = lane; c < V; c += 32) {
        const float v = x[c];
        if (nan_max_gt(v, best) || (nan_max_eq(v, best) && c < arg)) { best = v; arg = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (nan_max_gt(ob, best) || (nan_max_eq(ob, best) && oa < arg)) { best = ob; arg = oa; }
      }
      const double ref = is_pct ? 0.0 : vals[clamp_token(prev[(long long)(r0 + r) * prev_stride], V)];
      const int ps = sign_of(vals[arg] - ref);
      const int as = sign_of(vals[clamp_token(actual[(long long)(r0 + r) * actual_stride], V)] - ref);
      // softmax mass on the tokens whose direction equals the predicted one
      float tot = 0.f, dir = 0.f;
      for (int c = lane; c < V; c += 32) {
        const float e = __expf(x[c] - best);
        tot += e;
        if (sign_of(vals[c] - ref) == ps) dir += e;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
        dir += __shfl_xor_sync(0xffffffffu, dir, o);
      }
      if (lane == 0) {
        s_win[r] = ps == as;
        s_cert[r] = (double)dir / (double)tot;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int r = 0; r < n; ++r) {
        wins += s_win[r];
        losses += 1 - s_win[r];
        cert += s_cert[r];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[0] += wins;
    counts[1] += losses;
    certainty[0] += cert;
  }
}

}  // namespace ta

extern "C" int32_t ta_directional_metrics(const float* last_logits, int64_t ld, const double* vocab_values,
                                          const int64_t* prev_tokens, int64_t prev_stride, const int64_t* actual_tokens,
                                          int64_t actual_stride, int32_t rows, int32_t V, int32_t is_percentage,
                                          int64_t* counts, double* certainty, void* stream) {
  using namespace ta;
  TA_CHECK_ARG(last_logits && vocab_values && actual_tokens && counts && certainty, "ta_directional_metrics: null pointer");
  TA_CHECK_ARG(is_percentage || prev_tokens, "ta_directional_metrics: value data needs the previous input tokens");
  TA_CHECK_ARG(rows > 0 && V > 0, "ta_directional_metrics: bad sizes");
  launch_kernel(directional_metrics_kernel, dim3(1), dim3(DM_THREADS), 0, (cudaStream_t)stream, 
      last_logits, ld, vocab_values, (const long long*)prev_tokens, prev_stride, (const long long*)actual_tokens,
      actual_stride, rows, V, is_percentage ? 1 : 0, (long long*)counts, certainty);
  TA_CHECK_LAUNCH();
  return 0;
}
