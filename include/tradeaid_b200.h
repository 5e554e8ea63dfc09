/* tradeaid_b200.h — C ABI of libtradeaid_b200.so, the sm_100a kernel library under the trade-AId hot path.
 *
 * The reference (tsnuk/trade-AId-multimodal-transformer) is pure Python/PyTorch and has NO FFI, plugin or operator
 * interface of its own: its hot path is the nn.Module tree of model.py plus training_utils.get_batch. The drop-in
 * boundary is therefore the Python surface (model.MultimodalTransformer, training_utils.get_batch / estimate_loss,
 * see INTEGRATION.md); this header is the C ABI that sits UNDER that surface. Each entry point names the reference
 * call site (file:line in /root/reference) whose ATen work it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; sizes are element counts unless named *_bytes;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never allocates, never synchronises;
 *   - return value: 0 = ok, non-zero = error; ta_last_error() returns a thread-local message;
 *   - bf16 = __nv_bfloat16 (uint16_t storage), row-major, "ld" = elements between consecutive rows;
 *   - a "group" is an independent problem instance of identical shape (modalities, heads): g = g0 + G0*g1 and each
 *     operand carries two element strides (gs0, gs1) for it.
 */
#ifndef TRADEAID_B200_H
#define TRADEAID_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TA_ABI_VERSION 1

const char* ta_last_error(void);
int32_t ta_abi_version(void);
/* Number of kernels this library has launched since load (or since ta_reset_launch_count). */
int64_t ta_launch_count(void);
void ta_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------ K2: GEMM
 * D[g] = epilogue( alpha * A[g] (M x K) * B[g]^T (N x K) ), bf16 operands, fp32 accumulation in TMEM (tcgen05.mma),
 * operands staged by TMA (128B swizzle), persistent grid. Replaces every nn.Linear on the path:
 *   model.py:36-50 (q/k/v tanh-MLPs), :82-86 / :102-106 (projection MLPs), :127-131 (cross q / kv), :167-172 (FFN),
 *   :331-337 (output heads) and their autograd backward (dX = dY*W, dW = dY^T*X).
 * major = 0: the reduction dim is contiguous  (operand stored as [rows = M or N][K], ld >= K)     "K-major"
 * major = 1: the M/N dim is contiguous        (operand stored as [K][M or N],      ld >= M or N)  "MN-major"
 * so  Y = X W^T   : A = X (major 0), B = W  (major 0)
 *     dX = dY W   : A = dY (major 0), B = W  (major 1; reduction over W's rows)
 *     dW = dY^T X : A = dY (major 1), B = X  (major 1; reduction over the token rows)
 * ld and the group strides must be multiples of 8 elements (TMA: 16-byte global strides); ptr 16-byte aligned.
 */
typedef struct ta_operand {
  const void* ptr;
  int32_t major;
  int32_t _pad;
  int64_t ld;
  int64_t gs0, gs1;
} ta_operand;

enum { TA_ACT_NONE = 0, TA_ACT_TANH = 1, TA_ACT_RELU = 2 };
enum { TA_AUX_NONE = 0, TA_AUX_TANH_BWD = 1, TA_AUX_RELU_BWD = 2 };

typedef struct ta_epilogue {
  float alpha;        /* v = alpha * acc                                                     */
  int32_t act;        /* TA_ACT_*: v = act(v + bias)                                         */
  const float* bias;  /* [N] fp32 or NULL                                                    */
  int64_t bias_gs0, bias_gs1;
  const void* aux;    /* bf16 [M, ld_aux] saved activation h or NULL                         */
  int64_t ld_aux, aux_gs0, aux_gs1;
  int32_t aux_mode;   /* TA_AUX_TANH_BWD: v *= 1 - h^2 ; TA_AUX_RELU_BWD: v = h > 0 ? v : 0   */
  int32_t _pad;
  const float* residual; /* fp32 [M, ld_res] or NULL: v += residual                          */
  int64_t ld_res, res_gs0, res_gs1;
  float* out_f32;     /* fp32 [M, ld_f32] or NULL                                            */
  int64_t ld_f32, f32_gs0, f32_gs1;
  void* out_bf16;     /* bf16 [M, ld_bf16] or NULL                                           */
  int64_t ld_bf16, bf16_gs0, bf16_gs1;
  /* inverted dropout on v before the residual add (model.py:91, 116, 171); p = 0 disables. The keep mask of
   * element (g, row, col) is a pure function of (seed, g, row, col) so backward regenerates it. */
  float dropout_p;
  int32_t _pad2;
  uint64_t dropout_seed;
} ta_epilogue;

int32_t ta_gemm_bf16(const ta_operand* A, const ta_operand* B, int64_t M, int64_t N, int64_t K, int32_t G0,
                     int32_t G1, const ta_epilogue* ep, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRADEAID_B200_H */
