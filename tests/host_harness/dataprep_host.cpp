// Host build of the element arithmetic of csrc/dataprep_math.cuh (the same functions the CUDA kernels call), so that
// tests/test_dataprep_cpu.py can compare it with CPython / the reference on a machine without a GPU. TEST INFRASTRUCTURE:
// the product never loads this; it is compiled by the test with g++ -ffp-contract=off.
#include "../../trade-aid-multimodal-transformer_b200/csrc/dataprep_math.cuh"

using namespace ta::prep;

static Pow10 pow10_of(int nd) { return make_pow10(nd); }

extern "C" {

void h_py_round(const double* x, long long n, int nd, double* out) {
  const Pow10 P = pow10_of(nd);
  for (long long i = 0; i < n; ++i) out[i] = py_round(x[i], P);
}

// number of k in a pseudo-random sample (plus the edges) for which the division-free quotient differs from IEEE k / 10^nd
long long h_div_pow10_mismatches(int nd, long long samples, unsigned long long seed) {
  const Pow10 P = pow10_of(nd);
  long long bad = 0;
  unsigned long long st = seed;
  for (long long i = 0; i < samples; ++i) {
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    const int bits = 1 + (int)((st >> 58) % 53);
    unsigned long long k = (hash64(st) >> (64 - bits)) | (1ull << (bits - 1));
    if (i < 4096) k = (unsigned long long)i + 1;
    else if (i < 8192) k = (1ull << 53) - (unsigned long long)(i - 4096);
    const double a = (double)k;
    if (bits_of(div_pow10(a, P)) != bits_of(a / P.d)) ++bad;
  }
  return bad;
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

void h_repr_decimals(const double* x, long long n, int* out) {
  Pow10 P[AUTO_MAX_DECIMALS + 1];
  for (int d = 0; d <= AUTO_MAX_DECIMALS; ++d) P[d] = pow10_of(d);
  for (long long i = 0; i < n; ++i) out[i] = repr_decimals(x[i], P);
}

// out[i] = NaN and flags[i] = 1 where the element is left to the host
void h_range_auto(const double* x, const unsigned char* is_int, long long n, int has_nwd, double lower, double upper,
                  const double* upper_clamps, const double* thresholds, const double* factors, double* out, int* flags) {
  RangeAutoParams rp;
  rp.has_whole_digits = has_nwd;
  rp.lower = lower;
  rp.upper = upper;
  for (int d = 0; d <= AUTO_MAX_DECIMALS; ++d) { rp.P[d] = pow10_of(d); rp.upper_clamp[d] = upper_clamps[d]; }
  for (long long i = 0; i < n; ++i) {
    flags[i] = 0;
    out[i] = range_element_auto(x[i], is_int && is_int[i], rp, thresholds, factors, &flags[i]);
  }
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
