// HBM-bound row kernels of the hot path: K1 embedding gather (+ backward), K4 LayerNorm (+ backward),
// K5 softmax-cross-entropy (+ backward), K7 get_batch window gather, bias-gradient column sums, strided casts.
// All are bandwidth kernels: 128-bit loads/stores, one warp per row where a row fits a warp, grids sized from the
// row count (>> 148 SMs for every configuration of BASELINE.json), deterministic reductions (no float atomics).
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "common.cuh"
#include "philox.cuh"

namespace ta {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------ K7 get_batch
// x[m,b,t] = store[m, ix[b] + t], y[m,b,t] = store[m, ix[b] + t + 1]   (training_utils.py:378-379)
__global__ void gather_batch_kernel(const int32_t* __restrict__ store, long long store_stride, int M,
                                    const long long* __restrict__ ix, int B, int T, long long* __restrict__ xb,
                                    long long* __restrict__ yb) {
  pdl_prologue();
  const long long total = (long long)M * B * (T + 1);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(e % (T + 1));
    const long long mb = e / (T + 1);
    const int b = (int)(mb % B);
    const int m = (int)(mb / B);
    const long long v = store[m * store_stride + ix[b] + t];
    if (t < T) xb[mb * T + t] = v;
    if (t > 0) yb[mb * T + t - 1] = v;
  }
}

// ------------------------------------------------------------------------------------------------ K1 embedding
// x[m, r, :] = E[voff[m] + idx[m, r], :] + P[r % T, :]        (model.py:310-316)
__global__ void embed_fwd_kernel(const long long* __restrict__ idx, const float* __restrict__ tok,
                                 const long long* __restrict__ voff, const float* __restrict__ pos, float* __restrict__ x,
                                 int M, long long rows, int T, int C) {
  pdl_prologue();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)M * rows;
  for (long long r = blockIdx.x * (long long)warps_per_block + (threadIdx.x >> 5); r < total;
       r += (long long)gridDim.x * warps_per_block) {
    const int m = (int)(r / rows);
    const int t = (int)((r % rows) % T);
    const float* e = tok + (voff[m] + idx[r]) * C;
    const float* p = pos + (long long)t * C;
    float* o = x + r * C;
    if ((C & 3) == 0) {
      for (int c = lane * 4; c < C; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(e + c);
        const float4 b = *reinterpret_cast<const float4*>(p + c);
        *reinterpret_cast<float4*>(o + c) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      }
    } else {
      for (int c = lane; c < C; c += 32) o[c] = e[c] + p[c];
    }
  }
}

// Decode step under a CUDA graph: one new position per sequence whose index lives on the device.
// x[m, r, :] = E[voff[m] + idx[m, r], :] + P[*pos_dev, :]
__global__ void embed_fwd_at_kernel(const long long* __restrict__ idx, const float* __restrict__ tok,
                                    const long long* __restrict__ voff, const float* __restrict__ pos, float* __restrict__ x,
                                    int M, long long rows, int C, const int* __restrict__ pos_dev) {
  pdl_prologue();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)M * rows;
  const float* p = pos + (long long)(*pos_dev) * C;
  for (long long r = blockIdx.x * (long long)warps_per_block + (threadIdx.x >> 5); r < total;
       r += (long long)gridDim.x * warps_per_block) {
    const int m = (int)(r / rows);
    const float* e = tok + (voff[m] + idx[r]) * C;
    float* o = x + r * C;
    for (int c = lane; c < C; c += 32) o[c] = e[c] + p[c];
  }
}

// cache[g, b, *pos_dev, 0:W) = src[g, b, 0:W)  (32-bit words; the K/V rows of the new position)
__global__ void kv_append_kernel(const uint32_t* __restrict__ src, long long s_g, long long s_b, uint32_t* __restrict__ cache,
                                 long long c_g, long long c_b, long long c_t, int G, int B, int W, const int* __restrict__ pos_dev) {
  pdl_prologue();
  const long long t = *pos_dev;
  const long long total = (long long)G * B * W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(e % W);
    const long long gb = e / W;
    const int b = (int)(gb % B);
    const long long g = gb / B;
    cache[g * c_g + b * c_b + t * c_t + w] = src[g * s_g + b * s_b + w];
  }
}

// dP[t, c] = sum_m sum_b dx[m, b*T + t, c]   (position table shared by all modalities, model.py:298,313)
__global__ void embed_bwd_pos_kernel(const float* __restrict__ dx, float* __restrict__ dpos, int M, int B, int T, int C) {
  pdl_prologue();
  const int t = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int m = 0; m < M; ++m)
      for (int b = 0; b < B; ++b) acc += dx[(((long long)m * B + b) * T + t) * C + c];
    dpos[(long long)t * C + c] = acc;
  }
}

// dE[voff[m] + v, :] = sum over rows r of modality m with idx == v of dx[m, r, :], rows visited in increasing order
// (deterministic). One block per (m, v): scan the modality's indices in chunks, compact matches, accumulate.
__global__ void embed_bwd_tok_kernel(const long long* __restrict__ idx, const float* __restrict__ dx,
                                     const long long* __restrict__ voff, float* __restrict__ dtok, long long rows, int C) {
  pdl_prologue();
  constexpr int CHUNK = 1024;
  __shared__ int match[CHUNK];
  __shared__ int warp_cnt[32];
  __shared__ int n_match;
  const int m = blockIdx.y;
  const long long v = blockIdx.x;
  const long long vbeg = voff[m], vend = voff[m + 1];
  if (vbeg + v >= vend) return;
  const long long* id = idx + (long long)m * rows;
  const float* g = dx + (long long)m * rows * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  // each thread owns columns c = threadIdx.x + k*blockDim.x, k < 8 (C <= 8*blockDim.x), else loops
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (long long base = 0; base < rows; base += CHUNK) {
    // ordered compaction of matches in [base, base+CHUNK)
    int mine[CHUNK / 256];
    int cnt = 0;
    // blockDim.x == 256: thread handles positions tid*4 .. tid*4+3 (contiguous => order preserved)
    const int per = CHUNK / blockDim.x;
    for (int j = 0; j < per; ++j) {
      const long long pos = base + threadIdx.x * per + j;
      const bool hit = pos < rows && id[pos] == v;
      if (hit) mine[cnt++] = (int)(pos - base);
    }
    // exclusive scan of cnt over the block (warp shuffle + smem)
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) warp_cnt[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = lane < nwarp ? warp_cnt[lane] : 0;
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += n;
      }
      if (lane < nwarp) warp_cnt[lane] = wi - w;
      if (lane == 31) n_match = wi;
    }
    __syncthreads();
    const int off = warp_cnt[warp] + incl - cnt;
    for (int j = 0; j < cnt; ++j) match[off + j] = mine[j];
    __syncthreads();
    const int nm = n_match;
    for (int j = 0; j < nm; ++j) {
      const float* row = g + (base + match[j]) * C;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = threadIdx.x + k * blockDim.x;
        if (c < C) acc[k] += row[c];
      }
    }
    __syncthreads();
  }
  float* out = dtok + (vbeg + v) * C;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = threadIdx.x + k * blockDim.x;
    if (c < C) out[c] = acc[k];
  }
}

// ------------------------------------------------------------------------------------------------ K4 LayerNorm
// y = (x - mean) * rstd * gamma + beta, eps inside the sqrt (torch.nn.LayerNorm, model.py:189-190, 210, 330).
// One warp per row; the row is held in registers (C <= 32*4*NV) so x is read from HBM exactly once.
// y32 != nullptr: the output is written in fp32 instead (bf16x3 precision mode: it is split into hi + lo operands next).
template <int NV>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ y32, long long ldy, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, long long rows_total, long long rows_per_group, int C, float eps) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows_total) return;
  const int g = (int)(row / rows_per_group);
  const float* xr = x + row * C;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 4;
    v[k] = c < C ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += v[k].x + v[k].y + v[k].z + v[k].w;
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 4;
    if (c < C) {
      const float a = v[k].x - mean, b = v[k].y - mean, cc = v[k].z - mean, d = v[k].w - mean;
      q += a * a + b * b + cc * cc + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  const float* gm = gamma + (long long)g * C;
  const float* bt = beta + (long long)g * C;
  __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 4;
    if (c < C) {
      const float4 gg = *reinterpret_cast<const float4*>(gm + c);
      const float4 bb = *reinterpret_cast<const float4*>(bt + c);
      if (y32 != nullptr) {
        *reinterpret_cast<float4*>(y32 + row * ldy + c) =
            make_float4((v[k].x - mean) * rstd * gg.x + bb.x, (v[k].y - mean) * rstd * gg.y + bb.y,
                        (v[k].z - mean) * rstd * gg.z + bb.z, (v[k].w - mean) * rstd * gg.w + bb.w);
        continue;
      }
      __nv_bfloat162 lo = __floats2bfloat162_rn((v[k].x - mean) * rstd * gg.x + bb.x, (v[k].y - mean) * rstd * gg.y + bb.y);
      __nv_bfloat162 hi = __floats2bfloat162_rn((v[k].z - mean) * rstd * gg.z + bb.z, (v[k].w - mean) * rstd * gg.w + bb.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(yr + c) = u;
    }
  }
}

// LayerNorm backward. dy bf16 (gradient w.r.t. the LN output), x fp32 (saved LN input), mean/rstd saved.
//   dx = rstd * (g*dy - mean_c(g*dy) - xhat * mean_c(g*dy*xhat))
//   dstream_out = dstream_in + dx   (fp32, residual-stream gradient; dstream_in may be NULL => 0) and its bf16 copy
//   partial dgamma/dbeta per block, reduced in fixed order by partial_reduce_kernel.
// Layout: TPR threads share one row (thread t owns float4 columns (k*TPR + t)*4), 256/TPR rows side by side ("slots"),
// R independent row sets per pass so that R x slots rows (x, dy and the incoming stream gradient: 10 bytes/element) are
// in flight before the first reduction. The first version kept a whole row plus its dgamma/dbeta accumulators in one
// warp (193 registers, 8 warps per SM) and reached 27% of the HBM roofline on c4.
template <int TPR, int NV, int R>
// (three resident blocks cap a thread at 80 registers: fine for TPR = 256, where a thread owns one float4 per row; the
// narrower rows keep SLOTS x R rows per thread block in flight and spilled 124-164 bytes at 80, so they run two blocks)
__global__ void __launch_bounds__(256, NV == 1 ? (TPR == 256 ? 3 : 2) : 1)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const float* __restrict__ x,
              const float* __restrict__ gamma, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const float* __restrict__ dstream_in, float* __restrict__ dstream_out, __nv_bfloat16* __restrict__ dstream_bf16,
              long long ld_dbf, float* __restrict__ part_dgamma, float* __restrict__ part_dbeta, long long rows_per_group,
              int C) {
  pdl_prologue();
  constexpr int SLOTS = 256 / TPR;
  constexpr int WPR = TPR / 32;                 // warps per row
  __shared__ float red[2][2 * R][SLOTS][WPR > 1 ? WPR : 1];
  __shared__ float4 slot_sum[SLOTS > 1 ? 2 * SLOTS * TPR * NV : 1];
  const int tid = threadIdx.x, lane = tid & 31;
  const int slot = tid / TPR, t = tid % TPR, wr = t >> 5;
  const int g = blockIdx.y;
  const long long grow0 = (long long)g * rows_per_group;
  const float* gm = gamma + (long long)g * C;
  float4 gg[NV], dg[NV], db[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * TPR + t) * 4;
    gg[k] = c < C ? *reinterpret_cast<const float4*>(gm + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    dg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  constexpr int RPP = R * SLOTS;                // rows per pass
  const float inv_c = 1.f / C;
  int par = 0;
  for (long long pass = blockIdx.x; pass * RPP < rows_per_group; pass += gridDim.x, par ^= 1) {
    float4 xv[R][NV], din[R][NV];
    uint2 dv[R][NV];
    float mean[R], rstd[R];
    long long row[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long rg = pass * RPP + r * SLOTS + slot;
      const bool ok = rg < rows_per_group;
      row[r] = ok ? grow0 + rg : -1;
      mean[r] = ok ? mean_in[row[r]] : 0.f;
      rstd[r] = ok ? rstd_in[row[r]] : 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * TPR + t) * 4;
        const bool in = ok && c < C;
        xv[r][k] = in ? *reinterpret_cast<const float4*>(x + row[r] * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        dv[r][k] = in ? *reinterpret_cast<const uint2*>(dy + row[r] * lddy + c) : make_uint2(0u, 0u);
        din[r][k] = in && dstream_in ? *reinterpret_cast<const float4*>(dstream_in + row[r] * C + c)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float s1[R], s2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      s1[r] = s2[r] = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&dv[r][k].x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&dv[r][k].y);
        const float4 d = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
        float4& xh = xv[r][k];     // x -> xhat in place (zero-filled columns/rows: rstd = 0 or dy = 0 keep them inert)
        xh = make_float4((xh.x - mean[r]) * rstd[r], (xh.y - mean[r]) * rstd[r], (xh.z - mean[r]) * rstd[r],
                         (xh.w - mean[r]) * rstd[r]);
        dg[k].x += d.x * xh.x; dg[k].y += d.y * xh.y; dg[k].z += d.z * xh.z; dg[k].w += d.w * xh.w;
        db[k].x += d.x; db[k].y += d.y; db[k].z += d.z; db[k].w += d.w;
        const float4 gd = make_float4(d.x * gg[k].x, d.y * gg[k].y, d.z * gg[k].z, d.w * gg[k].w);
        s1[r] += gd.x + gd.y + gd.z + gd.w;
        s2[r] += gd.x * xh.x + gd.y * xh.y + gd.z * xh.z + gd.w * xh.w;
      }
      s1[r] = warp_sum(s1[r]);
      s2[r] = warp_sum(s2[r]);
    }
    if (WPR > 1) {
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          red[par][2 * r][slot][wr] = s1[r];
          red[par][2 * r + 1][slot][wr] = s2[r];
        }
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < WPR; ++w) {
          a += red[par][2 * r][slot][w];
          b += red[par][2 * r + 1][slot][w];
        }
        s1[r] = a;
        s2[r] = b;
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row[r] < 0) continue;
      const float m1 = s1[r] * inv_c, m2 = s2[r] * inv_c;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * TPR + t) * 4;
        if (c >= C) continue;
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&dv[r][k].x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&dv[r][k].y);
        const float4 xh = xv[r][k];
        float4 o;
        o.x = rstd[r] * (__low2float(lo) * gg[k].x - m1 - xh.x * m2) + din[r][k].x;
        o.y = rstd[r] * (__high2float(lo) * gg[k].y - m1 - xh.y * m2) + din[r][k].y;
        o.z = rstd[r] * (__low2float(hi) * gg[k].z - m1 - xh.z * m2) + din[r][k].z;
        o.w = rstd[r] * (__high2float(hi) * gg[k].w - m1 - xh.w * m2) + din[r][k].w;
        *reinterpret_cast<float4*>(dstream_out + row[r] * C + c) = o;
        if (dstream_bf16) {
          __nv_bfloat162 plo = __floats2bfloat162_rn(o.x, o.y), phi = __floats2bfloat162_rn(o.z, o.w);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&plo);
          u.y = *reinterpret_cast<uint32_t*>(&phi);
          *reinterpret_cast<uint2*>(dstream_bf16 + row[r] * ld_dbf + c) = u;
        }
      }
    }
  }
  // dgamma / dbeta partial of this block: sum over slots in fixed order, one row of the partial buffers per block
  const long long prow = ((long long)g * gridDim.x + blockIdx.x) * C;
  if (SLOTS > 1) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      slot_sum[(slot * NV + k) * TPR + t] = dg[k];
      slot_sum[((SLOTS + slot) * NV + k) * TPR + t] = db[k];
    }
    __syncthreads();
    if (slot == 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        for (int sl = 0; sl < SLOTS; ++sl) {
          const float4 u = slot_sum[(sl * NV + k) * TPR + t], w = slot_sum[((SLOTS + sl) * NV + k) * TPR + t];
          a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
          b.x += w.x; b.y += w.y; b.z += w.z; b.w += w.w;
        }
        dg[k] = a;
        db[k] = b;
      }
    }
  }
  if (slot == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * TPR + t) * 4;
      if (c < C) {
        *reinterpret_cast<float4*>(part_dgamma + prow + c) = dg[k];
        *reinterpret_cast<float4*>(part_dbeta + prow + c) = db[k];
      }
    }
  }
}

// out[g, c] = sum_{p < parts_per_group} part[g*parts_per_group + p, c]
__global__ void partial_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int parts_per_group, int C,
                                      long long out_gs) {
  pdl_prologue();
  const int g = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* p = part + (long long)g * parts_per_group * C + c;
  float a = 0.f;
  for (int i = 0; i < parts_per_group; ++i) a += p[(long long)i * C];
  out[(long long)g * out_gs + c] = a;
}

// ------------------------------------------------------------------------------------------------ K5 cross-entropy
// per row: lse = logsumexp(logits), loss_row = lse - logits[target]; dlogits = (softmax - onehot) * scale (bf16)
// (F.cross_entropy mean reduction, model.py:398; the mean is taken by row_mean_kernel in fixed order).
__global__ void __launch_bounds__(256)
ce_fwd_bwd_kernel(const float* __restrict__ logits, long long ldl, const long long* __restrict__ targets,
                  float* __restrict__ row_loss, __nv_bfloat16* __restrict__ dlogits, long long ldd, long long rows, int V,
                  float scale) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* l = logits + row * ldl;
  float mx = -INFINITY;
  for (int c = lane; c < V; c += 32) mx = fmaxf(mx, l[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < V; c += 32) s += expf(l[c] - mx);
  s = warp_sum(s);
  const float lse = mx + logf(s);
  const long long t = targets[row];
  if (lane == 0) row_loss[row] = lse - l[t];
  if (dlogits) {
    __nv_bfloat16* d = dlogits + row * ldd;
    const float inv = 1.f / s;
    for (int c = lane; c < V; c += 32) {
      float p = expf(l[c] - mx) * inv;
      if (c == t) p -= 1.f;
      d[c] = __float2bfloat16_rn(p * scale);
    }
  }
}

// out[g] = mean over rows of in[g, :rows]; one block per group, fixed summation tree => deterministic.
__global__ void row_mean_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows) {
  pdl_prologue();
  __shared__ float sh[32];
  const float* p = in + (long long)blockIdx.x * rows;
  float a = 0.f;
  for (long long i = threadIdx.x; i < rows; i += blockDim.x) a += p[i];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    float b = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    b = warp_sum(b);
    if (threadIdx.x == 0) out[blockIdx.x] = b / (float)rows;
  }
}

// ------------------------------------------------------------------------------------------------ bias gradients
// part[g, split, c] = sum over the split's rows of dy[g, r, c]  (dy bf16); reduced by partial_reduce_kernel.
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ dy, long long ld, long long gs, float* __restrict__ part,
                      long long rows, int N, int splits) {
  pdl_prologue();
  __shared__ float sh[8][64];
  const int g = blockIdx.z, split = blockIdx.y;
  const int c0 = blockIdx.x * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rows_per_split = (rows + splits - 1) / splits;
  const long long r0 = split * rows_per_split;
  const long long r1 = r0 + rows_per_split < rows ? r0 + rows_per_split : rows;
  const __nv_bfloat16* base = dy + (long long)g * gs;
  const int c = c0 + lane * 2;
  float a = 0.f, b = 0.f;
  if (c + 1 < N && (ld & 1) == 0) {
    for (long long r = r0 + warp; r < r1; r += 8) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(base + r * ld + c);
      a += __low2float(v);
      b += __high2float(v);
    }
  } else {
    for (long long r = r0 + warp; r < r1; r += 8) {
      if (c < N) a += __bfloat162float(base[r * ld + c]);
      if (c + 1 < N) b += __bfloat162float(base[r * ld + c + 1]);
    }
  }
  sh[warp][lane * 2] = a;
  sh[warp][lane * 2 + 1] = b;
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
    part[((long long)g * splits + split) * N + c0 + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------ strided casts
// dst[g, r, c] (bf16, zero padded) = src[g, r, c] (fp32) for r < R, c < Cs
__global__ void cast_pad_kernel(const float* __restrict__ src, long long sg, long long sr, __nv_bfloat16* __restrict__ dst,
                                long long dg, long long dr, int G, int R, int Cs, int Rd, int Cd) {
  pdl_prologue();
  const long long total = (long long)G * Rd * Cd;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cd);
    const long long gr = e / Cd;
    const int r = (int)(gr % Rd);
    const long long g = gr / Rd;
    const float v = (r < R && c < Cs) ? src[g * sg + (long long)r * sr + c] : 0.f;
    dst[g * dg + (long long)r * dr + c] = __float2bfloat16_rn(v);
  }
}

// dst (fp32, ld = dr) = src (fp32 padded, ld = sr) unpadded copy: dst[g, r, c] = src[g, r, c], r < R, c < Cn
__global__ void unpad_f32_kernel(const float* __restrict__ src, long long sg, long long sr, float* __restrict__ dst,
                                 long long dg, long long dr, int G, int R, int Cn) {
  pdl_prologue();
  const long long total = (long long)G * R * Cn;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cn);
    const long long gr = e / Cn;
    const int r = (int)(gr % R);
    const long long g = gr / R;
    dst[g * dg + (long long)r * dr + c] = src[g * sg + (long long)r * sr + c];
  }
}

// contiguous fp32 -> bf16 cast, 8 elements per thread (n % 8 == 0 handled by the vector body, tail scalar)
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_prologue();
  const long long nv = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i];
    const float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
    h[0] = __floats2bfloat162_rn(a.x, a.y);
    h[1] = __floats2bfloat162_rn(a.z, a.w);
    h[2] = __floats2bfloat162_rn(b.x, b.y);
    h[3] = __floats2bfloat162_rn(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = u;
  }
  if (blockIdx.x == 0)
    for (long long i = nv * 8 + threadIdx.x; i < n; i += blockDim.x) dst[i] = __float2bfloat16_rn(src[i]);
}

static int grid_for(long long work_items, int per_block, int max_blocks_per_sm = 16) {
  long long b = ceil_div(work_items, per_block);
  const long long cap = (long long)num_sms() * max_blocks_per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}


// ------------------------------------------------------------------------------------------------ dropout helpers
// dst[g, r, c] = keep(seed; plane g, row r, col c) ? src[g, r, c] / (1 - p) : 0 — the mask of the GEMM epilogue with the
// same seed, applied to the upstream gradient in backward (model.py:91, 116, 171). One thread per 8 columns.
__global__ void dropout_bf16_kernel(const __nv_bfloat16* __restrict__ src, long long ld_src, long long gs_src,
                                    __nv_bfloat16* __restrict__ dst, long long ld_dst, long long gs_dst, long long rows, int cols,
                                    unsigned thresh, float scale, unsigned long long seed, const unsigned long long* step) {
  pdl_prologue();
  const int c8n = (cols + 7) >> 3;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (idx >= rows * c8n) return;
  const long long row = idx / c8n;
  const int c8 = (int)(idx - row * c8n);
  const unsigned long long sd = seed + (step ? *step * 0x9E3779B97F4A7C15ull : 0ull);
  const unsigned m = dropout_keep8(sd, 0u, (unsigned)g, (unsigned)row, (unsigned)c8, thresh);
  const __nv_bfloat16* s = src + (long long)g * gs_src + row * ld_src + c8 * 8;
  __nv_bfloat16* d = dst + (long long)g * gs_dst + row * ld_dst + c8 * 8;
#pragma unroll
  for (int e = 0; e < 8; ++e)
    if (c8 * 8 + e < cols) d[e] = (m >> e) & 1u ? __float2bfloat16_rn(__bfloat162float(s[e]) * scale) : __float2bfloat16_rn(0.f);
}

// attention keep mask as bytes [n_src][G][B][H][T][T] (tests): plane = ((s*G + g)*B + b)*H + h, kind 1
__global__ void dropout_attn_mask_kernel(uint8_t* __restrict__ mask, long long planes, int T, unsigned thresh, unsigned long long seed,
                                         const unsigned long long* step) {
  pdl_prologue();
  const int c8n = (T + 7) >> 3;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= planes * T * c8n) return;
  const int c8 = (int)(idx % c8n);
  const long long pr = idx / c8n;
  const int row = (int)(pr % T);
  const long long plane = pr / T;
  const unsigned long long sd = seed + (step ? *step * 0x9E3779B97F4A7C15ull : 0ull);
  const unsigned m = dropout_keep8(sd, 1u, (unsigned)plane, (unsigned)row, (unsigned)c8, thresh);
  for (int e = 0; e < 8; ++e)
    if (c8 * 8 + e < T) mask[(plane * T + row) * T + c8 * 8 + e] = (m >> e) & 1u;
}

}  // namespace ta

using namespace ta;

extern "C" {

int32_t ta_gather_batch(const int32_t* store, int64_t store_stride, int32_t M, const int64_t* ix, int32_t B, int32_t T,
                        int64_t* xb, int64_t* yb, void* stream) {
  TA_CHECK_ARG(store && ix && xb && yb, "ta_gather_batch: null pointer");
  TA_CHECK_ARG(M > 0 && B > 0 && T > 0, "ta_gather_batch: non-positive size");
  const long long total = (long long)M * B * (T + 1);
  launch_kernel(gather_batch_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
      store, store_stride, M, (const long long*)ix, B, T, (long long*)xb, (long long*)yb);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_embed_fwd(const int64_t* idx, const float* tok, const int64_t* voff, const float* pos, float* x, int32_t M,
                     int64_t rows, int32_t T, int32_t C, void* stream) {
  TA_CHECK_ARG(idx && tok && voff && pos && x, "ta_embed_fwd: null pointer");
  TA_CHECK_ARG(M > 0 && rows > 0 && T > 0 && C > 0 && rows % T == 0, "ta_embed_fwd: bad sizes");
  launch_kernel(embed_fwd_kernel, dim3(grid_for((long long)M * rows, 8)), dim3(256), 0, (cudaStream_t)stream, 
      (const long long*)idx, tok, (const long long*)voff, pos, x, M, rows, T, C);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_embed_fwd_at(const int64_t* idx, const float* tok, const int64_t* voff, const float* pos, float* x, int32_t M,
                        int64_t rows, int32_t C, const int32_t* pos_dev, void* stream) {
  TA_CHECK_ARG(idx && tok && voff && pos && x && pos_dev, "ta_embed_fwd_at: null pointer");
  TA_CHECK_ARG(M > 0 && rows > 0 && C > 0, "ta_embed_fwd_at: bad sizes");
  const long long total = (long long)M * rows;
  launch_kernel(embed_fwd_at_kernel, dim3(grid_for(total, 8)), dim3(256), 0, (cudaStream_t)stream, (const long long*)idx, tok,
                (const long long*)voff, pos, x, (int)M, (long long)rows, (int)C, (const int*)pos_dev);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_kv_append(const void* src, int64_t s_g, int64_t s_b, void* cache, int64_t c_g, int64_t c_b, int64_t c_t, int32_t G,
                     int32_t B, int32_t W_words, const int32_t* pos_dev, void* stream) {
  TA_CHECK_ARG(src && cache && pos_dev, "ta_kv_append: null pointer");
  TA_CHECK_ARG(G > 0 && B > 0 && W_words > 0, "ta_kv_append: bad sizes");
  TA_CHECK_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(cache)) & 3) == 0, "ta_kv_append: 4-byte alignment");
  const long long total = (long long)G * B * W_words;
  launch_kernel(kv_append_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const uint32_t*)src,
                (long long)s_g, (long long)s_b, (uint32_t*)cache, (long long)c_g, (long long)c_b, (long long)c_t, (int)G, (int)B,
                (int)W_words, (const int*)pos_dev);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_embed_bwd(const int64_t* idx, const float* dx, const int64_t* voff, int32_t max_vocab, float* dtok,
                     float* dpos, int32_t M, int32_t B, int32_t T, int32_t C, void* stream) {
  TA_CHECK_ARG(idx && dx && voff && dtok && dpos, "ta_embed_bwd: null pointer");
  TA_CHECK_ARG(C <= 8 * 256, "ta_embed_bwd: n_embd > 2048 not supported");
  launch_kernel(embed_bwd_pos_kernel, dim3(T), dim3(256), 0, (cudaStream_t)stream, dx, dpos, M, B, T, C);
  TA_CHECK_LAUNCH();
  dim3 grid(max_vocab, M);
  launch_kernel(embed_bwd_tok_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const long long*)idx, dx, (const long long*)voff, dtok,
                                                               (long long)B * T, C);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_dropout_bf16(const void* src, int64_t ld_src, int64_t gs_src, void* dst, int64_t ld_dst, int64_t gs_dst, int32_t G,
                        int64_t rows, int32_t cols, float p, uint64_t seed, void* stream) {
  TA_CHECK_ARG(src && dst && G > 0 && rows > 0 && cols > 0, "ta_dropout_bf16: bad arguments");
  TA_CHECK_ARG(p >= 0.f && p < 1.f, "ta_dropout_bf16: p must be in [0, 1)");
  const long long n = rows * ((cols + 7) / 8);
  dim3 grid((unsigned)ceil_div(n, 256), (unsigned)G);
  launch_kernel(dropout_bf16_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)src, ld_src, gs_src, (__nv_bfloat16*)dst, ld_dst,
                                                            gs_dst, rows, cols, dropout_thresh16(p), 1.f / (1.f - p), seed,
                                                            dropout_step_counter());
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_dropout_attn_mask(uint8_t* mask, int32_t n_src, int32_t G, int32_t B, int32_t H, int32_t T, float p, uint64_t seed,
                             void* stream) {
  TA_CHECK_ARG(mask && n_src > 0 && G > 0 && B > 0 && H > 0 && T > 0, "ta_dropout_attn_mask: bad arguments");
  const long long planes = (long long)n_src * G * B * H;
  const long long n = planes * T * ((T + 7) / 8);
  launch_kernel(dropout_attn_mask_kernel, dim3((unsigned)ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, mask, planes, T, dropout_thresh16(p), seed,
                                                                                         dropout_step_counter());
  TA_CHECK_LAUNCH();
  return 0;
}

static int32_t layernorm_fwd_any(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y32, int64_t ldy,
                                 float* mean, float* rstd, int64_t rows_total, int64_t rows_per_group, int32_t C, float eps,
                                 void* stream) {
  TA_CHECK_ARG(x && gamma && beta && (y_bf16 || y32) && mean && rstd, "ta_layernorm_fwd: null pointer");
  TA_CHECK_ARG(C % 4 == 0 && C <= 4096 && ldy % 4 == 0, "ta_layernorm_fwd: n_embd must be a multiple of 4 and <= 4096");
  TA_CHECK_ARG(rows_total > 0 && rows_per_group > 0 && rows_total % rows_per_group == 0, "ta_layernorm_fwd: bad rows");
  TA_CHECK_ARG(!y32 || (reinterpret_cast<uintptr_t>(y32) & 15) == 0, "ta_layernorm_fwd_f32: y must be 16-byte aligned");
  const int grid = (int)ceil_div(rows_total, 8);
  auto st = (cudaStream_t)stream;
  auto* y = (__nv_bfloat16*)y_bf16;
#define TA_LN_FWD(NV) launch_kernel(ln_fwd_kernel<NV>, dim3(grid), dim3(256), 0, st, x, gamma, beta, y, y32, ldy, mean, rstd, rows_total, rows_per_group, C, eps)
  if (C <= 128) TA_LN_FWD(1);
  else if (C <= 256) TA_LN_FWD(2);
  else if (C <= 512) TA_LN_FWD(4);
  else if (C <= 1024) TA_LN_FWD(8);
  else if (C <= 2048) TA_LN_FWD(16);
  else TA_LN_FWD(32);
#undef TA_LN_FWD
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, int64_t ldy, float* mean,
                         float* rstd, int64_t rows_total, int64_t rows_per_group, int32_t C, float eps, void* stream) {
  TA_CHECK_ARG(y_bf16, "ta_layernorm_fwd: null pointer");
  return layernorm_fwd_any(x, gamma, beta, y_bf16, nullptr, ldy, mean, rstd, rows_total, rows_per_group, C, eps, stream);
}

int32_t ta_layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y_f32, int64_t ldy, float* mean,
                             float* rstd, int64_t rows_total, int64_t rows_per_group, int32_t C, float eps, void* stream) {
  TA_CHECK_ARG(y_f32, "ta_layernorm_fwd_f32: null pointer");
  return layernorm_fwd_any(x, gamma, beta, nullptr, y_f32, ldy, mean, rstd, rows_total, rows_per_group, C, eps, stream);
}

// launch shape of ln_bwd_kernel: TPR threads per row, R row sets per pass, grid.x blocks per group
static void ln_bwd_shape(int64_t rows_per_group, int64_t G, int C, int* tpr, int* r, int* gx) {
  *tpr = C <= 128 ? 32 : C <= 256 ? 64 : C <= 512 ? 128 : 256;
  *r = 4;
  const int64_t rpp = (int64_t)*r * (256 / *tpr);
  const int resident = C <= 512 ? 2 : 3;                    // the kernel's launch bounds
  int64_t want = ((int64_t)num_sms() * resident) / G;       // every block resident, over all groups: one wave
  const int64_t passes = ceil_div(rows_per_group, rpp);
  if (want > passes) want = passes;
  if (want < 1) want = 1;
  *gx = (int)want;
}

int64_t ta_layernorm_bwd_workspace_floats(int64_t rows_total, int64_t rows_per_group, int32_t C) {
  int tpr, r, gx;
  const int64_t G = rows_total / rows_per_group;
  ln_bwd_shape(rows_per_group, G, C, &tpr, &r, &gx);
  return 2 * G * gx * C;
}

int32_t ta_layernorm_bwd(const void* dy_bf16, int64_t lddy, const float* x, const float* gamma, const float* mean,
                         const float* rstd, const float* dstream_in, float* dstream_out, void* dstream_bf16,
                         int64_t ld_dbf, float* dgamma, float* dbeta, float* workspace, int64_t rows_total,
                         int64_t rows_per_group, int32_t C, void* stream) {
  TA_CHECK_ARG(dy_bf16 && x && gamma && mean && rstd && dstream_out && dgamma && dbeta && workspace,
               "ta_layernorm_bwd: null pointer");
  TA_CHECK_ARG(C % 4 == 0 && C <= 2048 && lddy % 4 == 0 && ld_dbf % 4 == 0, "ta_layernorm_bwd: n_embd % 4 != 0 or > 2048");
  TA_CHECK_ARG(rows_total > 0 && rows_per_group > 0 && rows_total % rows_per_group == 0, "ta_layernorm_bwd: bad rows");
  const int G = (int)(rows_total / rows_per_group);
  int tpr, r, gx;
  ln_bwd_shape(rows_per_group, G, C, &tpr, &r, &gx);
  float* pg = workspace;
  float* pb = workspace + (long long)G * gx * C;
  auto st = (cudaStream_t)stream;
  dim3 grid((unsigned)gx, (unsigned)G);
#define TA_LN_BWD(TPR, NV)                                                                                             \
  launch_kernel(ln_bwd_kernel<TPR, NV, 4>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)dy_bf16, lddy, x, gamma, mean, rstd, dstream_in, \
                                                  dstream_out, (__nv_bfloat16*)dstream_bf16, ld_dbf, pg, pb,            \
                                                  rows_per_group, C)
  if (tpr == 32) TA_LN_BWD(32, 1);
  else if (tpr == 64) TA_LN_BWD(64, 1);
  else if (tpr == 128) TA_LN_BWD(128, 1);
  else if (C <= 1024) TA_LN_BWD(256, 1);
  else TA_LN_BWD(256, 2);
#undef TA_LN_BWD
  TA_CHECK_LAUNCH();
  dim3 rg((unsigned)ceil_div(C, 256), G);
  launch_kernel(partial_reduce_kernel, dim3(rg), dim3(256), 0, st, pg, dgamma, gx, C, C);
  TA_CHECK_LAUNCH();
  launch_kernel(partial_reduce_kernel, dim3(rg), dim3(256), 0, st, pb, dbeta, gx, C, C);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_cross_entropy(const float* logits, int64_t ldl, const int64_t* targets, float* row_loss, float* loss_mean,
                         void* dlogits_bf16, int64_t ldd, int64_t rows, int32_t V, float grad_scale, void* stream) {
  TA_CHECK_ARG(logits && targets && row_loss && loss_mean, "ta_cross_entropy: null pointer");
  TA_CHECK_ARG(rows > 0 && V > 0, "ta_cross_entropy: bad sizes");
  auto st = (cudaStream_t)stream;
  launch_kernel(ce_fwd_bwd_kernel, dim3((int)ceil_div(rows, 8)), dim3(256), 0, st, logits, ldl, (const long long*)targets, row_loss,
                                                          (__nv_bfloat16*)dlogits_bf16, ldd, rows, V,
                                                          grad_scale / (float)rows);
  TA_CHECK_LAUNCH();
  launch_kernel(row_mean_kernel, dim3(1), dim3(1024), 0, st, row_loss, loss_mean, rows);
  TA_CHECK_LAUNCH();
  return 0;
}

int64_t ta_colsum_workspace_floats(int64_t rows, int32_t N, int32_t G) {
  const int splits = (int)(rows >= 4096 ? 16 : (rows >= 512 ? 4 : 1));
  return (int64_t)G * splits * N;
}

int32_t ta_colsum_bf16(const void* dy_bf16, int64_t ld, int64_t gs, float* out, int64_t out_gs, float* workspace,
                       int64_t rows, int32_t N, int32_t G, void* stream) {
  TA_CHECK_ARG(dy_bf16 && out && workspace, "ta_colsum_bf16: null pointer");
  const int splits = (int)(rows >= 4096 ? 16 : (rows >= 512 ? 4 : 1));
  auto st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div(N, 64), splits, G);
  launch_kernel(colsum_partial_kernel, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)dy_bf16, ld, gs, workspace, rows, N, splits);
  TA_CHECK_LAUNCH();
  dim3 rg((unsigned)ceil_div(N, 256), G);
  launch_kernel(partial_reduce_kernel, dim3(rg), dim3(256), 0, st, workspace, out, splits, N, out_gs);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_cast_pad_bf16(const float* src, int64_t sg, int64_t sr, void* dst_bf16, int64_t dg, int64_t dr, int32_t G,
                         int32_t R, int32_t Cs, int32_t Rd, int32_t Cd, void* stream) {
  TA_CHECK_ARG(src && dst_bf16, "ta_cast_pad_bf16: null pointer");
  const long long total = (long long)G * Rd * Cd;
  if (total == 0) return 0;
  launch_kernel(cast_pad_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, src, sg, sr, (__nv_bfloat16*)dst_bf16, dg, dr,
                                                                         G, R, Cs, Rd, Cd);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_unpad_f32(const float* src, int64_t sg, int64_t sr, float* dst, int64_t dg, int64_t dr, int32_t G, int32_t R,
                     int32_t Cn, void* stream) {
  TA_CHECK_ARG(src && dst, "ta_unpad_f32: null pointer");
  const long long total = (long long)G * R * Cn;
  if (total == 0) return 0;
  launch_kernel(unpad_f32_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, src, sg, sr, dst, dg, dr, G, R, Cn);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream) {
  TA_CHECK_ARG(src && dst_bf16, "ta_cast_bf16: null pointer");
  TA_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst_bf16 & 15) == 0, "ta_cast_bf16: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  launch_kernel(cast_bf16_kernel, dim3(grid_for(n / 8 + 1, 256)), dim3(256), 0, (cudaStream_t)stream, src, (__nv_bfloat16*)dst_bf16, n);
  TA_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
