// Error reporting, launch counting and device queries for libtradeaid_b200.so.
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>

#include "../../include/tradeaid_b200.h"

namespace ta {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static thread_local const unsigned long long* g_dropout_step = nullptr;   // device counter mixed into every dropout seed; per calling thread
static thread_local int g_sm_margin = 0;                      // SMs left free for a concurrent kernel (ta_set_sm_margin); per calling thread

char* err_buf() { return g_err; }

int set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

// -1: not set through ta_set_pdl -> the TA_PDL environment variable decides (default off)
static thread_local int g_pdl = -1;
bool pdl_enabled() {
  static const int env = [] { const char* e = getenv("TA_PDL"); return e ? atoi(e) : 0; }();
  return (g_pdl >= 0 ? g_pdl : env) != 0;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

const unsigned long long* dropout_step_counter() { return g_dropout_step; }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  // persistent grids leave `margin` SMs (whole SM pairs, so CTA-pair clusters still fit) to whatever runs beside them
  const int margin = g_sm_margin;
  const int n = (cached[dev] - margin) & ~1;
  return n < 2 ? 2 : n;
}

}  // namespace ta

extern "C" {
const char* ta_last_error(void) { return ta::err_buf(); }
int32_t ta_abi_version(void) { return TA_ABI_VERSION; }

int32_t ta_dropout_set_step_counter(const uint64_t* device_counter) {
  ta::g_dropout_step = reinterpret_cast<const unsigned long long*>(device_counter);
  return 0;
}
int32_t ta_set_sm_margin(int32_t sms) {
  TA_CHECK_ARG(sms >= 0 && sms <= 128, "ta_set_sm_margin: margin %d out of range [0, 128]", (int)sms);
  ta::g_sm_margin = sms;
  return 0;
}
int32_t ta_get_sm_margin(void) { return ta::g_sm_margin; }
int32_t ta_set_pdl(int32_t on) {
  ta::g_pdl = on < 0 ? -1 : (on ? 1 : 0);
  return 0;
}
int64_t ta_launch_count(void) { return ta::g_launches.load(std::memory_order_relaxed); }
void ta_reset_launch_count(void) { ta::g_launches.store(0, std::memory_order_relaxed); }
}
