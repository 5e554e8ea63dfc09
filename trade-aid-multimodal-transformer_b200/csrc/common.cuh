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

}  // namespace ta
