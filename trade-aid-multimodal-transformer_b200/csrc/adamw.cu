// K6 — fused AdamW over the flat fp32 master buffer, emitting the bf16 shadow the GEMMs read.
// Replaces torch.optim.AdamW(m.parameters(), lr) + optimizer.step() (/root/reference/main.py:464, 650) with torch's
// default hyper-parameters (betas 0.9/0.999, eps 1e-8, decoupled weight_decay 0.01 on every tensor, amsgrad off) and
// the same operation order as torch's implementation:
//   p *= 1 - lr*wd ; m += (g - m)*(1-b1) ; v = v*b2 + (1-b2)*g*g ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// HBM-bound: 16 B read + 14 B written per parameter, 128-bit accesses. The step counter and the bias corrections live
// in device memory (ta_adamw_tick) so the whole step can sit inside a replayed CUDA graph.
#include <cuda_bf16.h>

#include "../../include/tradeaid_b200.h"
#include "common.cuh"

namespace ta {

struct AdamScalars {  // device-resident, written by tick
  float step_size;    // lr / (1 - beta1^t)
  float bc2_sqrt;     // sqrt(1 - beta2^t)
  float decay;        // 1 - lr * wd
  float _pad;
};

__global__ void adamw_tick_kernel(long long* step, AdamScalars* sc, double lr, double beta1, double beta2, double wd) {
  pdl_prologue();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const long long t = *step + 1;
    *step = t;
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    sc->step_size = (float)(lr / bc1);
    sc->bc2_sqrt = (float)sqrt(bc2);
    sc->decay = (float)(1.0 - lr * wd);
  }
}

__device__ __forceinline__ void adam_one(float& p, float& m, float& v, float g, const AdamScalars& s, float b1w, float beta2,
                                         float b2w, float eps) {
  p *= s.decay;
  m = m + (g - m) * b1w;
  v = v * beta2 + b2w * g * g;
  const float denom = sqrtf(v) / s.bc2_sqrt + eps;
  p = p - s.step_size * (m / denom);
}

// GB16: the gradient is read as bf16 (the exchanged copy of a data-parallel step: trainer.py, grad_dtype = "bf16")
template <bool GB16>
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const void* __restrict__ gv, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ shadow, long long n, const AdamScalars* __restrict__ sc, float beta1, float beta2,
             float eps, float grad_scale) {
  pdl_prologue();
  const AdamScalars s = *sc;
  const float b1w = 1.f - beta1, b2w = 1.f - beta2;
  const long long nv = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg;
    if constexpr (GB16) {
      const uint2 u = reinterpret_cast<const uint2*>(gv)[i];
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      gg = make_float4(a.x, a.y, b.x, b.y);
    } else {
      gg = reinterpret_cast<const float4*>(gv)[i];
    }
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pp.x, mm.x, vv.x, gg.x * grad_scale, s, b1w, beta2, b2w, eps);
    adam_one(pp.y, mm.y, vv.y, gg.y * grad_scale, s, b1w, beta2, b2w, eps);
    adam_one(pp.z, mm.z, vv.z, gg.z * grad_scale, s, b1w, beta2, b2w, eps);
    adam_one(pp.w, mm.w, vv.w, gg.w * grad_scale, s, b1w, beta2, b2w, eps);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(shadow)[i] = u;
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = nv * 4 + threadIdx.x; i < n; i += blockDim.x) {
      float pp = p[i], mm = m[i], vv = v[i];
      const float gi = GB16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(gv)[i]) : reinterpret_cast<const float*>(gv)[i];
      adam_one(pp, mm, vv, gi * grad_scale, s, b1w, beta2, b2w, eps);
      p[i] = pp; m[i] = mm; v[i] = vv;
      if (shadow) shadow[i] = __float2bfloat16_rn(pp);
    }
  }
}

}  // namespace ta

using namespace ta;

extern "C" {

int32_t ta_adamw_tick(int64_t* step_dev, float* scalars_dev, double lr, double beta1, double beta2, double weight_decay,
                      void* stream) {
  TA_CHECK_ARG(step_dev && scalars_dev, "ta_adamw_tick: null pointer");
  launch_kernel(adamw_tick_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (long long*)step_dev, (AdamScalars*)scalars_dev, lr, beta1, beta2,
                                                       weight_decay);
  TA_CHECK_LAUNCH();
  return 0;
}

static int32_t adamw_step_any(float* p, const void* g, bool g_bf16, float* m, float* v, void* shadow_bf16, int64_t n,
                              const float* scalars_dev, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  TA_CHECK_ARG(p && g && m && v && scalars_dev, "ta_adamw_step: null pointer");
  TA_CHECK_ARG((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)g & (g_bf16 ? 7 : 15)) == 0 &&
               ((uintptr_t)shadow_bf16 & 7) == 0, "ta_adamw_step: buffers must be 16-byte aligned (bf16 buffers 8-byte)");
  if (n == 0) return 0;
  long long blocks = ceil_div(n / 4 + 1, 256);
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (g_bf16)
    launch_kernel(adamw_kernel<true>, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, (__nv_bfloat16*)shadow_bf16, n,
                  (const AdamScalars*)scalars_dev, beta1, beta2, eps, grad_scale);
  else
    launch_kernel(adamw_kernel<false>, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, (__nv_bfloat16*)shadow_bf16, n,
                  (const AdamScalars*)scalars_dev, beta1, beta2, eps, grad_scale);
  TA_CHECK_LAUNCH();
  return 0;
}

int32_t ta_adamw_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, const float* scalars_dev,
                      float beta1, float beta2, float eps, float grad_scale, void* stream) {
  return adamw_step_any(p, g, false, m, v, shadow_bf16, n, scalars_dev, beta1, beta2, eps, grad_scale, stream);
}

int32_t ta_adamw_step_bf16g(float* p, const void* g_bf16, float* m, float* v, void* shadow_bf16, int64_t n,
                            const float* scalars_dev, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  return adamw_step_any(p, g_bf16, true, m, v, shadow_bf16, n, scalars_dev, beta1, beta2, eps, grad_scale, stream);
}

}  // extern "C"
