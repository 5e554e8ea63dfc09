// Types shared by the two forward attention kernels (attention_fwd.cu: one tile at a time on 16 math warps;
// attention_fwd2.cu: two independent tile streams, one thread per query row, P kept in tensor memory).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/tradeaid_b200.h"

namespace ta {

constexpr int ATT_MAX_SRC = 8;
constexpr int AF_TILE = 16384;   // 128 rows x 128 bytes

struct AttnMaps {
  CUtensorMap q;
  CUtensorMap k[ATT_MAX_SRC];
  CUtensorMap v[ATT_MAX_SRC];
  CUtensorMap o;                     // TMA store targets
  CUtensorMap o_src[ATT_MAX_SRC];
};

struct AttnFwdParams {
  int T, B, H, G, hs, n_src, n_qtiles, total_work;
  int debug;   // TA_ATTN_DEBUG (dev only): 16 = no TMEM source sum, 32 = no output staging / stores
  uint32_t drop_thresh;            // dropout on the probabilities (model.py:69, 151): 16-bit lane threshold, 0 = off
  float drop_scale;                // 1 / (1 - p)
  unsigned long long drop_seed;
  const unsigned long long* drop_step;
  float scale_log2;
  __nv_bfloat16* o;       // [G][B*T][ldo], head h at column h*hs
  long long ldo, o_gs;
  __nv_bfloat16* o_src;   // optional per-source outputs [n_src][G][B*T][ldo] (needed by backward when n_src > 1)
  long long o_src_stride;
  float* lse;             // [n_src][G][B][H][T], natural log
};

// 5-D view {d < hs, head, t < T, batch, group} of a [G][B*T][ld] bf16 activation; box = {64, 1, box_rows, 1, 1}
int make_head_map(CUtensorMap* out, const ta_attn_tensor* t, int hs, int H, int T, int B, int G, int box_rows);
// attention_fwd2.cu
int launch_attn_fwd2(const AttnMaps& maps, const AttnFwdParams& p, cudaStream_t st);

}  // namespace ta
