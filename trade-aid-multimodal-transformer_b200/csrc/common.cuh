// Host-side helpers shared by all translation units of libtradeaid_b200.so.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

namespace ta {

// thread-local error text returned by ta_last_error()
char* err_buf();
int set_err(const char* fmt, ...);
void count_launch(int n = 1);

#define TA_CHECK_ARG(cond, ...)          \
  do {                                   \
    if (!(cond)) return ta::set_err(__VA_ARGS__); \
  } while (0)

#define TA_CHECK_CUDA(expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return ta::set_err("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define TA_CHECK_LAUNCH()                                                                                   \
  do {                                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                                    \
    if (_e != cudaSuccess) return ta::set_err("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    ta::count_launch();                                                                                     \
  } while (0)

int num_sms();
const unsigned long long* dropout_step_counter();   // see ta_dropout_set_step_counter

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE setting: a process that drives a second GPU after the first
// must opt in there too. `done` is the launcher's own static flag array (one per kernel instantiation), indexed by device.
constexpr int TA_MAX_DEVICES = 64;
template <typename Kernel>
static inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, bool (&done)[TA_MAX_DEVICES]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < TA_MAX_DEVICES && done[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < TA_MAX_DEVICES) done[dev] = true;
  return e;
}

// Programmatic dependent launch (PDL), opt-in with TA_PDL=1: every kernel of the library can be launched with the
// programmatic-stream-serialization attribute and starts with pdl_prologue() - `griddepcontrol.launch_dependents` lets the
// NEXT kernel of the stream be scheduled as soon as all CTAs of this one have started (its CTAs become resident as SMs
// free up and run their on-chip set-up), `griddepcontrol.wait` blocks until the PREVIOUS kernel has completed and its
// memory is visible, so no dependent global access can move ahead of it (results are bit-identical: tests). Inside a
// captured CUDA graph the edges become programmatic edges. Measured (CUDA-graph step, same box, PDL on vs off):
// c2 2.67 vs 2.79 ms (282 small kernels: launch-latency bound, -4.5 %), c3 13.60 vs 13.42 ms, c4 263.3 vs 258.6 ms
// (+1.8 %: with one persistent CTA per SM holding all of shared memory and TMEM the early-scheduled CTAs gain nothing
// and cost a little) - hence off by default; without the attribute both instructions are no-ops.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_launch_dependents(); pdl_wait(); }
#endif

}  // namespace ta
