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

#define TA_ABI_VERSION 9

const char* ta_last_error(void);
int32_t ta_abi_version(void);
/* Number of kernels this library has launched since load (or since ta_reset_launch_count). */
int64_t ta_launch_count(void);
void ta_reset_launch_count(void);
/* Persistent kernels (GEMM, attention, LayerNorm backward) size their grids from the SM count. While a collective
 * (the NCCL gradient all-reduce of the data-parallel step, main.py:646-650 under DP) runs on a side stream, its CTAs
 * hold SMs; a persistent grid of one CTA per SM would then run its displaced CTAs as a second wave. The margin (SMs,
 * rounded to whole SM pairs) is subtracted from the grid of every FOLLOWING launch made by the CALLING THREAD (two engines
 * driven from two threads do not see each other's margin); 0 restores full grids. */
int32_t ta_set_sm_margin(int32_t sms);
int32_t ta_get_sm_margin(void);
/* Programmatic dependent launch for the FOLLOWING launches of the calling thread: 1 = every kernel is launched with the
 * programmatic-stream-serialization attribute (its CTAs are scheduled while the previous kernel drains; griddepcontrol.wait
 * keeps the data dependence), 0 = plain launches, -1 = follow the TA_PDL environment variable (default off). It pays where a
 * step is a few hundred small launches (c2: -4.5 %) and costs a little where one persistent CTA per SM holds all of shared
 * memory (c4: +1.8 %), so the engine switches it on for small problems only. Results are bit-identical either way. */
int32_t ta_set_pdl(int32_t on);

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

/* Dropout (model.py:69, 91, 116, 151, 171). Masks are a pure function of (seed, coordinates): Philox4x32-7 keyed by the
 * 64-bit seed of the call, one draw per 8 consecutive columns of a row (csrc/philox.cuh), so backward regenerates them.
 * ta_dropout_set_step_counter registers a device counter (or NULL) that the FOLLOWING launches of the calling THREAD
 * capture in their parameters: its value is mixed into the seed when the kernel RUNS, so a captured CUDA graph draws
 * fresh masks on every replay as long as the graph increments the counter. Bindings set it around each launch.
 * ta_dropout_bf16: dst = keep ? src / (1 - p) : 0 with the GEMM-epilogue mask of the same seed (backward of the three
 * projection / feed-forward dropout sites). ta_dropout_attn_mask: the attention keep mask as bytes, for tests. */
int32_t ta_dropout_set_step_counter(const uint64_t* device_counter);
int32_t ta_dropout_bf16(const void* src, int64_t ld_src, int64_t gs_src, void* dst, int64_t ld_dst, int64_t gs_dst, int32_t G,
                        int64_t rows, int32_t cols, float p, uint64_t seed, void* stream);
int32_t ta_dropout_attn_mask(uint8_t* mask, int32_t n_src, int32_t G, int32_t B, int32_t H, int32_t T, float p, uint64_t seed,
                             void* stream);

int32_t ta_gemm_bf16(const ta_operand* A, const ta_operand* B, int64_t M, int64_t N, int64_t K, int32_t G0,
                     int32_t G1, const ta_epilogue* ep, void* stream);


/* ------------------------------------------------------------------------------------------------ K3: attention
 * Causal softmax(q k^T * hs^-0.5) v per (group, batch, head) without materialising T x T in HBM (tcgen05 + TMEM,
 * online softmax). Self-attention: model.py:60-73 (n_src = 1). Cross-attention: model.py:136-159 — one independent
 * causal softmax per key/value source, outputs summed (n_src = number of other modalities).
 * Every tensor is a bf16 activation [G][B*T][ld] whose head h occupies columns [h*head_stride, h*head_stride + hs);
 * ptr already points at column 0 of head 0 (so q/k/v slices of a packed buffer are expressed by ptr offsets).
 * lse: fp32 [n_src][G][B][H][T] (natural log), saved for backward. o_src (optional, required by backward when
 * n_src > 1): per-source normalised outputs, bf16 [n_src] x (same layout as o), o_src_stride elements apart.
 */
typedef struct ta_attn_tensor {
  const void* ptr;
  int64_t ld;          /* elements between consecutive token rows */
  int64_t head_stride; /* elements between consecutive heads      */
  int64_t gs;          /* elements between consecutive groups     */
} ta_attn_tensor;

int32_t ta_attention_fwd(const ta_attn_tensor* q, const ta_attn_tensor* k, const ta_attn_tensor* v, int32_t n_src,
                         void* o_bf16, int64_t ldo, int64_t o_gs, void* o_src_bf16, int64_t o_src_stride, float* lse,
                         int32_t G, int32_t B, int32_t T, int32_t H, int32_t hs, float dropout_p, uint64_t seed,
                         void* stream);

/* The same forward in plain fp32 (no tensor cores; q / k / v / o / o_src are FLOAT tensors, strides in elements): the exact
 * attention of the bf16x3 precision mode, where the bf16 operands and output of the tensor-core kernels are all that is left
 * of the logits error against the reference's fp32 forward. hs <= 64, T <= 1024, no dropout. */
int32_t ta_attention_fwd_f32(const ta_attn_tensor* q, const ta_attn_tensor* k, const ta_attn_tensor* v, int32_t n_src,
                             float* o, int64_t ldo, int64_t o_gs, float* o_src, int64_t o_src_stride, float* lse, int32_t G,
                             int32_t B, int32_t T, int32_t H, int32_t hs, void* stream);

/* Decode step of `generate` (model.py:404-446) while the sequence still fits block_size: ONE new query row per
 * (group, batch) against the first `len` rows of a key/value cache — the reference recomputes the whole window for
 * every new token; with causal masks and absolute positions the earlier rows do not change, so their keys / values are
 * cached. q: bf16 [G][B][ld] (ta_attn_tensor.ld = elements between batch rows). k[s], v[s]: caches [G][B][capacity][ld]
 * (ld = elements between cached positions, kv_batch_stride[s] = elements between batches; k and v of a source share
 * strides). Per source an independent softmax over positions [0, len), outputs summed; o: bf16 [G][B][ldo], head h at
 * columns [h*hs, (h+1)*hs). n_src <= 16, hs a multiple of 8 up to 64 (as the forward / backward kernels). */
int32_t ta_attention_decode(const ta_attn_tensor* q, const ta_attn_tensor* k, const ta_attn_tensor* v,
                            const int64_t* kv_batch_stride, int32_t n_src, void* o_bf16, int64_t ldo, int64_t o_gs,
                            int32_t G, int32_t B, int32_t len, const int32_t* len_dev, int32_t H, int32_t hs, void* stream);
/* The decode step as a replayable CUDA graph keeps the number of cached positions in device memory: with len_dev != NULL the
 * decode kernel attends to *len_dev + len positions; ta_embed_fwd_at embeds one new position per sequence at index *pos_dev
 * (x[m, r, :] = tok[voff[m] + idx[m, r], :] + pos[*pos_dev, :]); ta_kv_append writes the new position's K/V rows (W_words
 * 32-bit words per (group, batch), element strides in words) at row *pos_dev of a cache [G][B][capacity][...]. */
int32_t ta_embed_fwd_at(const int64_t* idx, const float* tok, const int64_t* voff, const float* pos, float* x, int32_t M,
                        int64_t rows, int32_t C, const int32_t* pos_dev, void* stream);
int32_t ta_kv_append(const void* src, int64_t s_g, int64_t s_b, void* cache, int64_t c_g, int64_t c_b, int64_t c_t, int32_t G,
                     int32_t B, int32_t W_words, const int32_t* pos_dev, void* stream);

/* Backward of the above (recompute): given dO (layout of o) produces dq (layout of q), and per source dk, dv (layouts
 * of k[s], v[s]); all bf16. T <= 256 in this build. Two launches on `stream`: one pass over dO and the per-source outputs
 * writes D = rowsum(dO * O_s) into `workspace` (ta_attention_bwd_workspace_bytes(n_src, G, B, T, H) bytes of device memory,
 * fp32 [n_src][G][B][H][T], the layout of lse), then the tile kernel. */
int64_t ta_attention_bwd_workspace_bytes(int32_t n_src, int32_t G, int32_t B, int32_t T, int32_t H);
int32_t ta_attention_bwd(const ta_attn_tensor* q, const ta_attn_tensor* k, const ta_attn_tensor* v, int32_t n_src,
                         const void* o_bf16, const void* o_src_bf16, int64_t o_src_stride, const void* do_bf16,
                         int64_t ldo, int64_t o_gs, const float* lse, void* workspace, const ta_attn_tensor* dq,
                         const ta_attn_tensor* dk, const ta_attn_tensor* dv, int32_t G, int32_t B, int32_t T, int32_t H,
                         int32_t hs, float dropout_p, uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------------------ K7: get_batch
 * xb[m,b,t] = store[m, ix[b]+t], yb[m,b,t] = store[m, ix[b]+t+1] (training_utils.py:378-379). store: int32 token ids
 * [M][store_stride] resident in HBM; ix: int64 [B] (drawn on the host by torch.randint, training_utils.py:104/150, so
 * the index stream is bit-exact); xb/yb: int64 [M][B][T]. */
int32_t ta_gather_batch(const int32_t* store, int64_t store_stride, int32_t M, const int64_t* ix, int32_t B, int32_t T,
                        int64_t* xb, int64_t* yb, void* stream);

/* ------------------------------------------------------------------------------------------------ K1: embedding
 * x[m, r, :] = tok[voff[m] + idx[m, r], :] + pos[r % T, :] (model.py:300-319). idx int64 [M][rows]; tok fp32
 * [sum V, C] (tables of all modalities stacked, voff int64 [M+1] row offsets); pos fp32 [>=T, C]; x fp32 [M][rows][C].
 * Backward: dtok rows (deterministic, ordered accumulation) and dpos[t] = sum over modalities and batch. */
int32_t ta_embed_fwd(const int64_t* idx, const float* tok, const int64_t* voff, const float* pos, float* x, int32_t M,
                     int64_t rows, int32_t T, int32_t C, void* stream);
int32_t ta_embed_bwd(const int64_t* idx, const float* dx, const int64_t* voff, int32_t max_vocab, float* dtok,
                     float* dpos, int32_t M, int32_t B, int32_t T, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------ K4: LayerNorm
 * nn.LayerNorm(C), eps inside the sqrt, affine (model.py:189-190, 210, 330). x fp32 [rows_total][C] -> y bf16 (ldy);
 * gamma/beta fp32 [G][C] with G = rows_total / rows_per_group; mean/rstd fp32 [rows_total] saved for backward.
 * Backward: dy bf16; dstream_out = dstream_in (nullable) + dx (fp32) with an optional bf16 copy; dgamma/dbeta [G][C].
 */
int32_t ta_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, int64_t ldy, float* mean,
                         float* rstd, int64_t rows_total, int64_t rows_per_group, int32_t C, float eps, void* stream);
/* same, fp32 output (the bf16x3 precision mode splits it into hi + lo operands next: ta_split_bf16) */
int32_t ta_layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y_f32, int64_t ldy, float* mean,
                             float* rstd, int64_t rows_total, int64_t rows_per_group, int32_t C, float eps, void* stream);
int64_t ta_layernorm_bwd_workspace_floats(int64_t rows_total, int64_t rows_per_group, int32_t C);
int32_t ta_layernorm_bwd(const void* dy_bf16, int64_t lddy, const float* x, const float* gamma, const float* mean,
                         const float* rstd, const float* dstream_in, float* dstream_out, void* dstream_bf16,
                         int64_t ld_dbf, float* dgamma, float* dbeta, float* workspace, int64_t rows_total,
                         int64_t rows_per_group, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------ K5: cross-entropy
 * F.cross_entropy(logits.view(B*T, V), targets.view(B*T)) mean reduction (model.py:394-399), fused with its backward:
 * row_loss[r] = logsumexp - logit[target]; loss_mean[0] = mean_r row_loss (fixed-order reduction);
 * dlogits (bf16, nullable) = (softmax - onehot) * grad_scale / rows. */
int32_t ta_cross_entropy(const float* logits, int64_t ldl, const int64_t* targets, float* row_loss, float* loss_mean,
                         void* dlogits_bf16, int64_t ldd, int64_t rows, int32_t V, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------ K7b: augmentation
 * data_utils.add_rand_to_data_points (data_utils.py:293-358; the `[2]` quirk of get_batch, training_utils.py:353-360) on
 * one modality of the device-resident token store, bit-exact to the host loop INCLUDING the consumption of Python's
 * Mersenne-Twister stream: words = the next W 32-bit outputs of the generator (pre-drawn by the host); every eligible
 * token (rand_size < v < vocab_size - rand_size), in index order, takes the next word whose top bit_length(2*rand_size+1)
 * bits are < 2*rand_size+1 (CPython's _randbelow rejection) and adds [0, 1, -1, 2, -2, 3, -3][that value].
 * stats[0] = eligible tokens E, stats[1] = accepted words A among the W given, stats[2] = words consumed (index of the
 * E-th accepted word + 1; 0 if E = 0). If A < E nothing is modified: call again with more words. */
int64_t ta_augment_workspace_bytes(int64_t N, int64_t W);
int32_t ta_augment_tokens(int32_t* tokens, int64_t N, const uint32_t* words, int64_t W, int32_t rand_size, int32_t vocab_size,
                          void* workspace, int64_t* stats, void* stream);

/* ------------------------------------------------------------------------------------------------ K8: evaluation metrics
 * Directional metrics of one modality for one batch (training_utils.py:183-330, calculate_evaluation_metrics): from the
 * LAST-position logits (fp32, row r at last_logits + r*ld) the arg-max token's value, its direction sign and the actual
 * next value's sign against the previous input value (vocab_values[prev_tokens[r*prev_stride]]; against 0 when
 * is_percentage, prev_tokens may then be NULL), and the softmax mass on the tokens of the predicted direction.
 * counts[0] += wins, counts[1] += losses, certainty[0] += sum of masses (device accumulators: estimate_loss adds
 * eval_iters batches and reads once). vocab_values: fp64 [V], the numeric vocabulary (python floats are doubles). */
int32_t ta_directional_metrics(const float* last_logits, int64_t ld, const double* vocab_values, const int64_t* prev_tokens,
                               int64_t prev_stride, const int64_t* actual_tokens, int64_t actual_stride, int32_t rows,
                               int32_t V, int32_t is_percentage, int64_t* counts, double* certainty, void* stream);

/* ------------------------------------------------------------------------------------------------ K9: data preparation
 * One-shot preparation of one modality's column on the device. The column is float64 [N] in device
 * memory (python floats are doubles); every step is bit-exact to the reference's Python arithmetic, including CPython's
 * round(x, ndigits) (exact decimal rounding, half to even) — see csrc/dataprep_math.cuh.
 * status: int64 [TA_PREP_STATUS_WORDS] device array, reset with ta_prep_status_reset; word w holds the FIRST row for which
 * condition w occurred (INT64_MAX = never). The kernels do not fail on bad rows; the host wrapper reads status and raises
 * what the reference raises. */
#define TA_PREP_STATUS_WORDS 4
#define TA_PREP_STATUS_NONFINITE 0   /* NaN / inf element */
#define TA_PREP_STATUS_ZERO_PREV 1   /* percent change with a zero previous value (row of that previous value) */
#define TA_PREP_STATUS_OVERFLOW 2    /* ranged value not finite (10**k factor beyond DBL_MAX) */
#define TA_PREP_STATUS_TABLE_FULL 3  /* vocabulary hash table too small: retry with a larger capacity */
#define TA_PREP_MAX_RANKS 4
#define TA_PREP_MAX_GROUPS 8192
int32_t ta_prep_status_reset(int64_t* status, void* stream);

/* convert_to_percent_changes (data_utils.py:612-662; file_cache.py:329-385 is the same arithmetic):
 * out[i] = round(((x[i] - x[i-1]) / x[i-1]) * 100, decimals), 0.0 for the first row of every file (file_starts: ascending
 * int64 [n_files] device array of first rows, file_starts[0] = 0). A zero previous value gives 0.0 and is recorded in
 * status[ZERO_PREV]: data_utils raises ZeroDivisionError for it (:652), file_cache keeps the 0.0 (:366-372).
 * Not in place. 0 <= decimals <= 22. */
int32_t ta_prep_percent_changes(const double* x, double* out, int64_t N, const int64_t* file_starts, int32_t n_files,
                                int32_t decimals, int64_t* status, void* stream);

/* range_numeric_data (data_utils.py:361-470) with decimal_places given: per element
 * round(x * 10**(nwd-1 - floor(log10|x|)), decimals), then the magnitude clamps of :440-455 (lower = 10**(nwd-1),
 * upper = 10**nwd, upper_clamp = what a value >= upper becomes). has_whole_digits = 0: only the rounding (:434).
 * thresholds / factors: fp64 [633] device tables over the decades -324..308 — thresholds[k] = the smallest double whose
 * floor(math.log10(.)) is >= k-324 under the host's libm, factors[k] = float(10**(nwd-1-(k-324))) (inf if not
 * representable); both built by the host wrapper with Python's own expressions. May run in place. */
int32_t ta_prep_range(const double* x, double* out, int64_t N, int32_t has_whole_digits, int32_t decimals, double lower,
                      double upper, double upper_clamp, const double* thresholds, const double* factors, int64_t* status,
                      void* stream);

/* range_numeric_data with decimal_places = None (data_utils.py:431): every element is rounded to the length of the fraction its
 * own str() shows - max(1, shortest round-tripping number of decimals) for a float in repr's positional range, 0 for a Python
 * int (is_int: optional uint8 [N] mask). Elements repr() prints in exponent notation (|x| >= 1e16 or 0 < |x| < 1e-4) come back
 * as NaN and are counted in n_flagged[0] (device int64, += ): the host wrapper patches those few with Python's own expression.
 * params: ta_prep_range_auto_params fills a host block of ta_prep_range_auto_params_bytes() bytes (powers of ten, the clamps
 * upper_clamps_host[d] = what a value >= 10**nwd becomes at d decimals, d = 0..22, built with Python's expressions); the caller
 * copies it to the device. thresholds / factors as ta_prep_range (needed with whole digits). May run in place. */
int64_t ta_prep_range_auto_params_bytes(void);
int32_t ta_prep_range_auto_params(void* params_host, int32_t has_whole_digits, double lower, double upper,
                                  const double* upper_clamps_host);
int32_t ta_prep_range_auto(const double* x, const uint8_t* is_int, double* out, int64_t N, const void* params_dev,
                           const double* thresholds, const double* factors, int64_t* status, int64_t* n_flagged, void* stream);

/* Order statistics for np.percentile (data_utils.py:503-504): out[r] = the ranks[r]-th smallest element (0-based) of x,
 * n_ranks <= TA_PREP_MAX_RANKS, by an 8-pass radix select (no sort, x is not modified). ranks / out are device arrays. */
int64_t ta_prep_select_workspace_bytes(void);
int32_t ta_prep_order_statistics(const double* x, int64_t N, const int64_t* ranks_dev, int32_t n_ranks, double* out,
                                 void* workspace, int64_t* status, void* stream);

/* Outlier filter of bin_numeric_data (:505-512): out3 = {min, max, count} over the rows with lo <= x <= hi. */
int64_t ta_prep_minmax_workspace_bytes(int64_t N);
int32_t ta_prep_filtered_minmax(const double* x, int64_t N, double lo, double hi, double* out3, void* workspace,
                                void* stream);

/* Group assignment of bin_numeric_data (:527-557): pos_bounds = [0, b_1 .. b_G], neg_bounds = [-b_G .. -b_1, 0] (fp64
 * [G+1] device arrays built by the host with the reference's expressions); out[i] = the group number as a double
 * (-G .. G). bounds_sorted != 0: the caller has checked that both arrays are non-decreasing, the reference's linear walk
 * (:531-536, :548-553) is then done as a bisection with the same result; 0 keeps the walk. May run in place. */
int32_t ta_prep_bin_assign(const double* x, double* out, int64_t N, const double* pos_bounds, const double* neg_bounds,
                           int32_t num_groups, int32_t bounds_sorted, void* stream);

/* numerical_representation (data_utils.py:212-225): vocabulary = sorted(set(x)), token = index in it.
 * insert: distinct values of x into an open-addressing table (uint64 [capacity], capacity a power of two; cleared here);
 *   counters[0] = number of distinct values. The caller retries with a larger table if status[TABLE_FULL] is set or the
 *   load exceeds 1/2.
 * build: ranks the n_distinct values by counting (O(V^2) compares, V is a vocabulary size): vocabulary fp64 [n_distinct]
 *   ascending, token_of_slot int32 [capacity]; keys_ws uint64 [n_distinct], slots_ws int32 [n_distinct].
 * lookup: tokens[i] = token of x[i] (int32: a row of the token store the feeder gathers from). */
int32_t ta_prep_vocab_insert(const double* x, int64_t N, uint64_t* table, int64_t capacity, int64_t* counters,
                             int64_t* status, void* stream);
int32_t ta_prep_vocab_build(const uint64_t* table, int64_t capacity, int64_t n_distinct, uint64_t* keys_ws, int32_t* slots_ws,
                            double* vocabulary, int32_t* token_of_slot, int64_t* counters, void* stream);
int32_t ta_prep_vocab_lookup(const double* x, int64_t N, const uint64_t* table, int64_t capacity, const int32_t* token_of_slot,
                             int32_t* tokens, void* stream);

/* bias gradients: out[g, c] = sum_r dy[g, r, c] (dy bf16 [G][rows][ld]); deterministic two-stage reduction. */
int64_t ta_colsum_workspace_floats(int64_t rows, int32_t N, int32_t G);
int32_t ta_colsum_bf16(const void* dy_bf16, int64_t ld, int64_t gs, float* out, int64_t out_gs, float* workspace,
                       int64_t rows, int32_t N, int32_t G, void* stream);

/* fp32 -> bf16 casts (bf16 GEMM-operand shadows of weights and of the un-normalised stream, model.py:143). */
int32_t ta_cast_pad_bf16(const float* src, int64_t sg, int64_t sr, void* dst_bf16, int64_t dg, int64_t dr, int32_t G,
                         int32_t R, int32_t Cs, int32_t Rd, int32_t Cd, void* stream);
int32_t ta_unpad_f32(const float* src, int64_t sg, int64_t sr, float* dst, int64_t dg, int64_t dr, int32_t G, int32_t R,
                     int32_t Cn, void* stream);
int32_t ta_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------ bf16x3 precision mode
 * The north star holds logits to 2e-3 of the reference's fp32 forward (model.py:380-402); single-pass bf16 operands give
 * ~5e-3. An fp32 value is written as hi + lo (hi = bf16(x), lo = bf16(x - hi)) and a product of two such values as
 * a_hi*b_hi + a_lo*b_hi + a_hi*b_lo: ONE ta_gemm_bf16 launch over K-concatenated operands [A_hi|A_lo|A_hi] x [B_hi|B_hi|B_lo].
 * ta_split_bf16 writes `terms` slots of a [G1][G0][R][C] fp32 view (unit inner stride; g1/g0/r element strides) side by
 * side along the columns (concat_rows = 0: slot k at columns [k*C, (k+1)*C)) or stacked along the rows (concat_rows = 1:
 * slot k at rows [k*R, (k+1)*R), for MN-major operands); bit k of lo_mask selects lo for slot k (A: 0b010, B: 0b100).
 * ta_convert: strided fp32 <-> bf16 conversion / copy of such a view (the attention kernels keep bf16 I/O in every mode). */
int32_t ta_split_bf16(const float* src, int64_t s_g1, int64_t s_g0, int64_t s_r, void* dst_bf16, int64_t d_g1, int64_t d_g0,
                      int64_t d_r, int32_t G1, int32_t G0, int32_t R, int32_t C, int32_t terms, int32_t lo_mask,
                      int32_t concat_rows, void* stream);
int32_t ta_convert(const void* src, int32_t src_is_bf16, int64_t s_g1, int64_t s_g0, int64_t s_r, void* dst, int32_t dst_is_bf16,
                   int64_t d_g1, int64_t d_g0, int64_t d_r, int32_t G1, int32_t G0, int32_t R, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------ K6: AdamW
 * torch.optim.AdamW defaults (main.py:464, 650): betas (0.9, 0.999), eps 1e-8, decoupled weight_decay 0.01, on the flat
 * fp32 master buffer; also writes the bf16 shadow (nullable). ta_adamw_tick advances the device-resident step counter
 * (int64[1]) and derives the bias-corrected scalars (float[4]) so the step replays inside a CUDA graph. */
int32_t ta_adamw_tick(int64_t* step_dev, float* scalars_dev, double lr, double beta1, double beta2, double weight_decay,
                      void* stream);
int32_t ta_adamw_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, const float* scalars_dev,
                      float beta1, float beta2, float eps, float grad_scale, void* stream);
/* same with the gradient read as bf16: the data-parallel step may exchange its gradient buckets in bf16 (half the NVLink
 * bytes of the all-reduce that follows main.py:646 under DP); the reduced bf16 buckets are consumed here without a cast back. */
int32_t ta_adamw_step_bf16g(float* p, const void* g_bf16, float* m, float* v, void* shadow_bf16, int64_t n,
                            const float* scalars_dev, float beta1, float beta2, float eps, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRADEAID_B200_H */
