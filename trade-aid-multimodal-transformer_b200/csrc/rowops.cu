// HBM-bound row kernels of the hot path: K1 embedding gather (+ backward), K4 LayerNorm (+ backward),
// K5 softmax-cross-entropy (+ backward), K7 get_batch window gather, bias-gradient column sums, strided casts.
// All are bandwidth kernels: 128-bit loads/stores, one warp per row where a row fits a warp, grids sized from the
// row count (>> 148 SMs for every configuration of BASELINE.json), deterministic reductions (no float atomics).
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "common.cuh"

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

// dP[t, c] = sum_m sum_b dx[m, b*T + t, c]   (position table shared by all modalities, model.py:298,313)
__global__ void embed_bwd_pos_kernel(const float* __restrict__ dx, float* __restrict__ dpos, int M, int B, int T, int C) {
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
template <int NV>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, long long ldy, float* __restrict__ mean_out, float* __restrict__ rstd_out,
              long long rows_total, long long rows_per_group, int C, float eps) {
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
//   partial dgamma/dbeta per block (rows handled by the block), reduced by ln_bwd_reduce_kernel.
template <int NV>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const float* __restrict__ x,
              const float* __restrict__ gamma, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const float* __restrict__ dstream_in, float* __restrict__ dstream_out, __nv_bfloat16* __restrict__ dstream_bf16,
              long long ld_dbf, float* __restrict__ part_dgamma, float* __restrict__ part_dbeta, long long rows_total,
              long long rows_per_group, int rows_per_block, int C) {
  extern __shared__ float sh[];  // [2][warps][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  // a block never straddles two groups: rows_per_group % rows_per_block == 0 is enforced by the host
  const int g = (int)(row0 / rows_per_group);
  const float* gm = gamma + (long long)g * C;
  float4 gg[NV], dg[NV], db[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 4;
    gg[k] = c < C ? *reinterpret_cast<const float4*>(gm + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    dg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int rr = warp; rr < rows_per_block; rr += nwarp) {
    const long long row = row0 + rr;
    if (row >= rows_total) break;
    const float mean = mean_in[row], rstd = rstd_in[row];
    const float* xr = x + row * C;
    const __nv_bfloat16* dyr = dy + row * lddy;
    float4 xh[NV], d[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 4;
      if (c < C) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + c);
        const uint2 u = *reinterpret_cast<const uint2*>(dyr + c);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        d[k] = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
        xh[k] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        dg[k].x += d[k].x * xh[k].x; dg[k].y += d[k].y * xh[k].y; dg[k].z += d[k].z * xh[k].z; dg[k].w += d[k].w * xh[k].w;
        db[k].x += d[k].x; db[k].y += d[k].y; db[k].z += d[k].z; db[k].w += d[k].w;
        d[k].x *= gg[k].x; d[k].y *= gg[k].y; d[k].z *= gg[k].z; d[k].w *= gg[k].w;
        s1 += d[k].x + d[k].y + d[k].z + d[k].w;
        s2 += d[k].x * xh[k].x + d[k].y * xh[k].y + d[k].z * xh[k].z + d[k].w * xh[k].w;
      } else {
        xh[k] = d[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    const float* din = dstream_in ? dstream_in + row * C : nullptr;
    float* dout = dstream_out + row * C;
    __nv_bfloat16* dbf = dstream_bf16 ? dstream_bf16 + row * ld_dbf : nullptr;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 4;
      if (c < C) {
        float4 r = make_float4(rstd * (d[k].x - s1 - xh[k].x * s2), rstd * (d[k].y - s1 - xh[k].y * s2),
                               rstd * (d[k].z - s1 - xh[k].z * s2), rstd * (d[k].w - s1 - xh[k].w * s2));
        if (din) {
          const float4 a = *reinterpret_cast<const float4*>(din + c);
          r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
        }
        *reinterpret_cast<float4*>(dout + c) = r;
        if (dbf) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&lo);
          u.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(dbf + c) = u;
        }
      }
    }
  }
  // block reduction of dgamma / dbeta partials over warps (fixed order => deterministic)
  float* sg = sh;
  float* sb = sh + (size_t)nwarp * C;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 4;
    if (c < C) {
      *reinterpret_cast<float4*>(sg + (size_t)warp * C + c) = dg[k];
      *reinterpret_cast<float4*>(sb + (size_t)warp * C + c) = db[k];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < nwarp; ++w) {
      a += sg[(size_t)w * C + c];
      b += sb[(size_t)w * C + c];
    }
    part_dgamma[(long long)blockIdx.x * C + c] = a;
    part_dbeta[(long long)blockIdx.x * C + c] = b;
  }
}

// out[g, c] = sum_{p < parts_per_group} part[g*parts_per_group + p, c]
__global__ void partial_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int parts_per_group, int C,
                                      long long out_gs) {
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

}  // namespace ta

using namespace ta;

extern "C" {

int32_t ta_gather_batch(const int32_t* store, int64_t store_stride, int32_t M, const int64_t* ix, int32_t B, int32_t T,
                        int64_t* xb, int64_t* yb, void* stream) {
  TA_CHECK_ARG(store && ix && xb && yb, "ta_gather_batch: null pointer");
  TA_CHECK_ARG(M > 0 && B > 0 && T > 0, "ta_gather_batch: non-positive size");
  const long long total = (long long)M * B * (T + 1);
  gather_batch_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      store, store_stride, M, (const long long*)ix, B, T, (long long*)xb, (long long*)yb);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_embed_fwd(const int64_t* idx, const float* tok, const int64_t* voff, const float* pos, float* x, int32_t M,
                     int64_t rows, int32_t T, int32_t C, void* stream) {
  TA_CHECK_ARG(idx && tok && voff && pos && x, "ta_embed_fwd: null pointer");
  TA_CHECK_ARG(M > 0 && rows > 0 && T > 0 && C > 0 && rows % T == 0, "ta_embed_fwd: bad sizes");
  embed_fwd_kernel<<<grid_for((long long)M * rows, 8), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)idx, tok, (const long long*)voff, pos, x, M, rows, T, C);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_embed_bwd(const int64_t* idx, const float* dx, const int64_t* voff, int32_t max_vocab, float* dtok,
                     float* dpos, int32_t M, int32_t B, int32_t T, int32_t C, void* stream) {
  TA_CHECK_ARG(idx && dx && voff && dtok && dpos, "ta_embed_bwd: null pointer");
  TA_CHECK_ARG(C <= 8 * 256, "ta_embed_bwd: n_embd > 2048 not supported");
  embed_bwd_pos_kernel<<<T, 256, 0, (cudaStream_t)stream>>>(dx, dpos, M, B, T, C);
  TA_CHECK_LAUNCH();
  dim3 grid(max_vocab, M);
  embed_bwd_tok_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)idx, dx, (const long long*)voff, dtok,
                                                               (long long)B * T, C);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, int64_t ldy, float* mean,
                         float* rstd, int64_t rows_total, int64_t rows_per_group, int32_t C, float eps, void* stream) {
  TA_CHECK_ARG(x && gamma && beta && y_bf16 && mean && rstd, "ta_layernorm_fwd: null pointer");
  TA_CHECK_ARG(C % 4 == 0 && C <= 4096 && ldy % 4 == 0, "ta_layernorm_fwd: n_embd must be a multiple of 4 and <= 4096");
  TA_CHECK_ARG(rows_total > 0 && rows_per_group > 0 && rows_total % rows_per_group == 0, "ta_layernorm_fwd: bad rows");
  const int grid = (int)ceil_div(rows_total, 8);
  auto st = (cudaStream_t)stream;
  auto* y = (__nv_bfloat16*)y_bf16;
#define TA_LN_FWD(NV) ln_fwd_kernel<NV><<<grid, 256, 0, st>>>(x, gamma, beta, y, ldy, mean, rstd, rows_total, rows_per_group, C, eps)
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

static int ln_bwd_rows_per_block(int64_t rows_total, int64_t rows_per_group) {
  int rpb = 64;
  if (rows_total != rows_per_group)  // grouped: a block must not straddle two groups
    while (rows_per_group % rpb) rpb >>= 1;
  return rpb;
}

int64_t ta_layernorm_bwd_workspace_floats(int64_t rows_total, int64_t rows_per_group, int32_t C) {
  const int rpb = ln_bwd_rows_per_block(rows_total, rows_per_group);
  return 2 * (rows_total / rows_per_group) * ceil_div(rows_per_group, rpb) * C;
}

int32_t ta_layernorm_bwd(const void* dy_bf16, int64_t lddy, const float* x, const float* gamma, const float* mean,
                         const float* rstd, const float* dstream_in, float* dstream_out, void* dstream_bf16,
                         int64_t ld_dbf, float* dgamma, float* dbeta, float* workspace, int64_t rows_total,
                         int64_t rows_per_group, int32_t C, void* stream) {
  TA_CHECK_ARG(dy_bf16 && x && gamma && mean && rstd && dstream_out && dgamma && dbeta && workspace,
               "ta_layernorm_bwd: null pointer");
  TA_CHECK_ARG(C % 4 == 0 && C <= 2048 && lddy % 4 == 0 && ld_dbf % 4 == 0, "ta_layernorm_bwd: n_embd % 4 != 0 or > 2048");
  TA_CHECK_ARG(rows_total > 0 && rows_per_group > 0 && rows_total % rows_per_group == 0, "ta_layernorm_bwd: bad rows");
  const int rpb = ln_bwd_rows_per_block(rows_total, rows_per_group);
  const long long blocks_per_group = ceil_div(rows_per_group, rpb);
  const int G = (int)(rows_total / rows_per_group);
  const int grid = (int)(blocks_per_group * G);
  float* pg = workspace;
  float* pb = workspace + (long long)grid * C;
  auto st = (cudaStream_t)stream;
  const size_t shm = (size_t)2 * 8 * C * sizeof(float);
#define TA_LN_BWD(NV)                                                                                                  \
  do {                                                                                                                 \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      cudaFuncSetAttribute(ln_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * 128 * NV * 4);     \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    ln_bwd_kernel<NV><<<grid, 256, shm, st>>>((const __nv_bfloat16*)dy_bf16, lddy, x, gamma, mean, rstd, dstream_in,   \
                                              dstream_out, (__nv_bfloat16*)dstream_bf16, ld_dbf, pg, pb, rows_total,   \
                                              rows_per_group, rpb, C);                                                 \
  } while (0)
  if (C <= 128) TA_LN_BWD(1);
  else if (C <= 256) TA_LN_BWD(2);
  else if (C <= 512) TA_LN_BWD(4);
  else if (C <= 1024) TA_LN_BWD(8);
  else TA_LN_BWD(16);
#undef TA_LN_BWD
  TA_CHECK_LAUNCH();
  dim3 rg((unsigned)ceil_div(C, 256), G);
  partial_reduce_kernel<<<rg, 256, 0, st>>>(pg, dgamma, (int)blocks_per_group, C, C);
  TA_CHECK_LAUNCH();
  partial_reduce_kernel<<<rg, 256, 0, st>>>(pb, dbeta, (int)blocks_per_group, C, C);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_cross_entropy(const float* logits, int64_t ldl, const int64_t* targets, float* row_loss, float* loss_mean,
                         void* dlogits_bf16, int64_t ldd, int64_t rows, int32_t V, float grad_scale, void* stream) {
  TA_CHECK_ARG(logits && targets && row_loss && loss_mean, "ta_cross_entropy: null pointer");
  TA_CHECK_ARG(rows > 0 && V > 0, "ta_cross_entropy: bad sizes");
  auto st = (cudaStream_t)stream;
  ce_fwd_bwd_kernel<<<(int)ceil_div(rows, 8), 256, 0, st>>>(logits, ldl, (const long long*)targets, row_loss,
                                                          (__nv_bfloat16*)dlogits_bf16, ldd, rows, V,
                                                          grad_scale / (float)rows);
  TA_CHECK_LAUNCH();
  row_mean_kernel<<<1, 1024, 0, st>>>(row_loss, loss_mean, rows);
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
  colsum_partial_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy_bf16, ld, gs, workspace, rows, N, splits);
  TA_CHECK_LAUNCH();
  dim3 rg((unsigned)ceil_div(N, 256), G);
  partial_reduce_kernel<<<rg, 256, 0, st>>>(workspace, out, splits, N, out_gs);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_cast_pad_bf16(const float* src, int64_t sg, int64_t sr, void* dst_bf16, int64_t dg, int64_t dr, int32_t G,
                         int32_t R, int32_t Cs, int32_t Rd, int32_t Cd, void* stream) {
  TA_CHECK_ARG(src && dst_bf16, "ta_cast_pad_bf16: null pointer");
  const long long total = (long long)G * Rd * Cd;
  if (total == 0) return 0;
  cast_pad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, sg, sr, (__nv_bfloat16*)dst_bf16, dg, dr,
                                                                         G, R, Cs, Rd, Cd);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_unpad_f32(const float* src, int64_t sg, int64_t sr, float* dst, int64_t dg, int64_t dr, int32_t G, int32_t R,
                     int32_t Cn, void* stream) {
  TA_CHECK_ARG(src && dst, "ta_unpad_f32: null pointer");
  const long long total = (long long)G * R * Cn;
  if (total == 0) return 0;
  unpad_f32_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, sg, sr, dst, dg, dr, G, R, Cn);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream) {
  TA_CHECK_ARG(src && dst_bf16, "ta_cast_bf16: null pointer");
  TA_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst_bf16 & 15) == 0, "ta_cast_bf16: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  cast_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst_bf16, n);
  TA_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
