// Host build of the element arithmetic of csrc/dataprep_math.cuh (the same functions the CUDA kernels call), so that
// tests/test_dataprep_cpu.py can compare it with CPython / the reference on a machine without a GPU. TEST INFRASTRUCTURE:
// the product never loads this; it is compiled by the test with g++ -ffp-contract=off.
#include "../../trade-aid-multimodal-transformer_b200/csrc/dataprep_math.cuh"

using namespace ta::prep;

static Pow10 pow10_of(int nd) {
  u128 p = 1;
  double d = 1.0;
  for (int i = 0; i < nd; ++i) { p *= 10; d *= 10.0; }   // exact up to 10^22
  Pow10 P;
  P.lo = (u64)p;
  P.hi = (u64)(p >> 64);
  P.d = d;
  return P;
}

extern "C" {

void h_py_round(const double* x, long long n, int nd, double* out) {
  const Pow10 P = pow10_of(nd);
  for (long long i = 0; i < n; ++i) out[i] = py_round(x[i], P);
}

void h_floor_log10(const double* ax, long long n, const double* thresholds, int* out) {
  for (long long i = 0; i < n; ++i) out[i] = floor_log10(ax[i], thresholds);
}

void h_range(const double* x, long long n, int has_nwd, int decimals, double lower, double upper, double upper_clamp,
             const double* thresholds, const double* factors, double* out) {
  RangeParams rp;
  rp.has_whole_digits = has_nwd;
  rp.decimals = decimals;
  rp.P = pow10_of(decimals);
  rp.lower = lower;
  rp.upper = upper;
  rp.upper_clamp = upper_clamp;
  for (long long i = 0; i < n; ++i) out[i] = range_element(x[i], rp, thresholds, factors);
}

void h_percent(const double* x, long long n, int nd, double* out) {
  const Pow10 P = pow10_of(nd);
  out[0] = 0.0;
  for (long long i = 1; i < n; ++i) out[i] = x[i - 1] == 0.0 ? 0.0 : percent_change(x[i], x[i - 1], P);
}

void h_bin(const double* x, long long n, const double* pos, const double* neg, int G, int sorted, int* out) {
  for (long long i = 0; i < n; ++i) out[i] = bin_of(x[i], pos, neg, G, sorted != 0);
}

void h_order_key(const double* x, long long n, unsigned long long* keys, double* back) {
  for (long long i = 0; i < n; ++i) { keys[i] = order_key(x[i]); back[i] = order_value(keys[i]); }
}

}
