// Exact element arithmetic of the one-shot data preparation (row f-4: /root/reference/data_utils.py:212-225, 361-662,
// /root/reference/file_cache.py:329-385). Everything here is host+device: the kernels in dataprep.cu call it on the GPU,
// tests/host_harness/dataprep_host.cpp compiles the same functions with g++ so that the arithmetic can be checked
// against CPython on a machine without a GPU.
//
// The reference computes in Python floats (IEEE doubles) with three library calls whose results must be reproduced to
// the bit, because equal values become one vocabulary entry and the vocabulary index IS the token id:
//   * round(x, ndigits)   — CPython float.__round__ (Objects/floatobject.c double_round): the EXACT binary value is
//     rounded half-to-even to `ndigits` decimals (dtoa mode 3) and the decimal string converted back with
//     correctly-rounded strtod. Restated in integer arithmetic: x = m * 2^e2 exactly, k = RHE(m * 10^nd / 2^-e2) in
//     128-bit integers, result = RN(k / 10^nd) (one IEEE division when k < 2^53, a long division otherwise).
//   * floor(math.log10(|x|)) — glibc's log10 rounds up to an integer for arguments a few ulps below a power of ten,
//     so floor() of it is NOT the exact decade. The host builds the table of thresholds t_p = smallest double with
//     floor(log10(t)) >= p with the box's own libm (tradeaid data_utils._log10_thresholds); the device only compares.
//   * 10**k — int for k >= 0 (converted to double when multiplied with a float), float pow for k < 0: the host builds
//     the per-decade factor table with Python's own expression.
// All products/quotients use the *_rn intrinsics on the device: no FMA contraction may change a rounding.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define TA_HD __host__ __device__ __forceinline__
#else
#define TA_HD inline
#endif

namespace ta {
namespace prep {

typedef unsigned long long u64;
typedef unsigned __int128 u128;

constexpr int P10_MIN = -324;   // floor(log10(5e-324))
constexpr int P10_MAX = 308;
constexpr int P10_COUNT = P10_MAX - P10_MIN + 1;   // 633 decades
constexpr int MAX_DECIMALS = 22;                   // 10^22 is the largest power of ten that is an exact double

TA_HD u64 bits_of(double x) {
#if defined(__CUDA_ARCH__)
  return (u64)__double_as_longlong(x);
#else
  u64 b;
  memcpy(&b, &x, 8);
  return b;
#endif
}
TA_HD double double_of(u64 b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double x;
  memcpy(&x, &b, 8);
  return x;
#endif
}
TA_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
TA_HD double sub_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
TA_HD double div_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
TA_HD bool is_finite(double x) { return ((bits_of(x) >> 52) & 0x7ff) != 0x7ff; }
TA_HD double abs_of(double x) { return double_of(bits_of(x) & 0x7fffffffffffffffull); }

TA_HD double fma_rn(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}

// 10^nd as an exact 128-bit integer (nd <= 22: < 2^74), as the (exact) double, and its correctly rounded reciprocal.
struct Pow10 {
  u64 lo, hi;
  double d;
  double inv;
};

TA_HD Pow10 make_pow10(int nd) {
  u128 p = 1;
  double d = 1.0;
  for (int i = 0; i < nd; ++i) { p *= 10; d *= 10.0; }   // exact up to 10^22
  Pow10 P;
  P.lo = (u64)p;
  P.hi = (u64)(p >> 64);
  P.d = d;
  P.inv = div_rn(1.0, d);
  return P;
}

// RN(k / 10^nd) for an integer-valued 0 < k <= 2^53 without a division: Markstein's correction. With y = RN(1/b) and a
// faithful quotient q, r = a - b*q is exact in one FMA and RN(q + r*y) is the correctly rounded a/b (Markstein 1990; the
// one exception, a divisor whose significand is all ones, is not a power of ten). The first q = RN(a*y) can be 2 ulps off,
// so the correction runs twice. tests/test_dataprep_cpu.py holds this to IEEE division on 10^7 random k per nd.
TA_HD double div_pow10(double k, const Pow10& P) {
  double q = mul_rn(k, P.inv);
  q = fma_rn(fma_rn(-q, P.d, k), P.inv, q);
  return fma_rn(fma_rn(-q, P.d, k), P.inv, q);
}

// m * 2^e (m < 2^54, result normal): exact scaling by a power of two without ldexp.
TA_HD double scale_pow2(u64 m, int e) { return mul_rn((double)m, double_of((u64)(e + 1023) << 52)); }

TA_HD u64 mul_hi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (u64)(((u128)a * b) >> 64);
#endif
}

#if defined(__CUDACC__)
#define TA_HD_COLD __host__ __device__ __noinline__
#else
#define TA_HD_COLD inline
#endif

// General form of py_round: any finite or non-finite x, any 10^nd <= 10^22. Kept out of line on the device (128-bit
// shifts and divisions): the kernels' loops only carry the 64-bit form below.
TA_HD_COLD double py_round_wide(double x, Pow10 P) {
  const u64 b = bits_of(x);
  const int be = (int)((b >> 52) & 0x7ff);
  const u64 frac = b & ((1ull << 52) - 1);
  if (be == 0x7ff) return x;                 // inf / nan round to themselves
  if (be >= 1075) return x;                  // |x| >= 2^52: an integer
  u64 m;
  int s;                                     // |x| = m / 2^s exactly
  if (be == 0) { m = frac; s = 1074; } else { m = frac | (1ull << 52); s = 1075 - be; }
  if (m == 0) return x;                      // +-0.0
  const u64 sign = b & 0x8000000000000000ull;
  const u128 p = ((u128)P.hi << 64) | P.lo;
  const u128 prod = (u128)m * p;             // < 2^53 * 2^74
  u128 k = 0;
  if (s < 128) {                             // s >= 128: prod < 2^127 <= 2^(s-1), the value is below one half
    k = prod >> s;
    const u128 rem = prod - (k << s);
    const u128 half = (u128)1 << (s - 1);
    if (rem > half || (rem == half && ((u64)k & 1))) k += 1;
  }
  double r;
  if (k == 0) {
    r = 0.0;
  } else if (k < ((u128)1 << 53)) {
    r = div_rn((double)(u64)k, P.d);         // both exact doubles: IEEE division = correctly rounded k / 10^nd
  } else {
    // correctly rounded k / 10^nd for k >= 2^53 (only when |x| * 10^nd >= 2^53): binary long division to 55 bits
    u128 q = k / p, rem = k % p;             // q <= 2^53
    int t = 0;
    while (q < ((u128)1 << 54)) {
      rem <<= 1;
      q <<= 1;
      if (rem >= p) { rem -= p; q |= 1; }
      ++t;
    }
    const u64 q64 = (u64)q;                  // 55 bits: 53 + guard + first sticky
    u64 mant = q64 >> 2;
    const bool guard = (q64 >> 1) & 1, sticky = (q64 & 1) || rem != 0;
    if (guard && (sticky || (mant & 1))) mant += 1;
    r = scale_pow2(mant, 2 - t);
  }
  return double_of(bits_of(r) | sign);
}

// CPython round(x, nd), 0 <= nd <= 22. Inline form: 2^-11 <= |x| < 2^42 (so that 11 <= s <= 63: the scaled value
// m*10^nd / 2^s needs no 128-bit shift and k <= 2^53) with m * 10^nd < 2^64 (always for nd <= 3, |x| < 2^42); everything
// else - zeros, tiny, huge and non-finite values, many decimals - takes the general form.
TA_HD double py_round(double x, const Pow10& P) {
  const u64 b = bits_of(x);
  const unsigned be = (unsigned)(b >> 52) & 0x7ffu;
  if ((b << 1) == 0) return x;               // +-0.0 (unchanged prices: common in percent data)
  if (be - 1012u <= 52u && P.hi == 0) {
    const u64 m = (b & ((1ull << 52) - 1)) | (1ull << 52);
    if (mul_hi64(m, P.lo) == 0) {
      const int s = 1075 - (int)be;
      const u64 prod = m * P.lo;
      u64 k = prod >> s;
      const u64 rem = prod & ((1ull << s) - 1), half = 1ull << (s - 1);
      k += (rem > half) | ((rem == half) & (k & 1));
      const double r = k ? div_pow10((double)k, P) : 0.0;
      return double_of(bits_of(r) | (b & 0x8000000000000000ull));
    }
  }
  return py_round_wide(x, P);
}

// floor(math.log10(ax)) for a finite ax > 0 through the host-built threshold table (ascending, P10_COUNT entries;
// thresholds[k] = smallest double whose floor(log10) is >= k + P10_MIN). A binade [2^e, 2^(e+1)) meets at most two
// decades, floor(e * log10(2)) and the next one, so the binary exponent gives the candidate and two comparisons with the
// table settle it (the table, not the arithmetic, carries libm's rounding near the powers of ten).
TA_HD int floor_log10(double ax, const double* __restrict__ thresholds) {
  const u64 b = bits_of(ax);
  int e = (int)(b >> 52) - 1023;
  if (e == -1023) {                          // subnormal: exponent of the leading fraction bit
    int lead = 51;
    while (!((b >> lead) & 1)) --lead;
    e = lead - 1074;
  }
  int p = (e * 78913) >> 18;                 // floor(e * log10(2)) for |e| <= 1650
  if (p < P10_MAX && ax >= thresholds[p + 1 - P10_MIN]) ++p;
  else if (ax < thresholds[p - P10_MIN]) --p;
  return p;
}

struct RangeParams {
  int has_whole_digits;   // num_whole_digits is not None
  int decimals;           // apply_dec_places (decimal_places is not None on this path)
  Pow10 P;                // 10^decimals
  double lower;           // 10**(nwd-1)
  double upper;           // 10**nwd
  double upper_clamp;     // 10**nwd - 10**(-decimals) if decimals > 0 else 10**nwd - 1
};

// One element of range_numeric_data (data_utils.py:421-457). x finite. factors[p - P10_MIN] = float(10**(nwd-1-p)), +inf
// where Python cannot convert the int power to a float: the result is then +inf (never a legal ranged value).
TA_HD double range_element(double x, const RangeParams& rp, const double* __restrict__ thresholds,
                           const double* __restrict__ factors) {
  double f = 1.0;
  if (rp.has_whole_digits) {
    const int p = (x == 0.0) ? 0 : floor_log10(abs_of(x), thresholds);
    f = factors[p - P10_MIN];
    if (!is_finite(f)) return f;             // 10**k beyond DBL_MAX: Python raises OverflowError (the caller reports the row)
  }
  double scaled = (f != 0.0) ? py_round(mul_rn(x, f), rp.P) : 0.0;
  if (rp.has_whole_digits) {
    double a = abs_of(scaled);
    if (a < rp.lower && a > 0.0) a = rp.lower;
    if (a >= rp.upper) a = rp.upper_clamp;
    scaled = (x < 0.0) ? -a : a;
  }
  return scaled;
}

// ---- range_numeric_data with decimal_places = None (data_utils.py:431): every element is rounded to the number of characters
// after the '.' of ITS OWN str(). For a float in repr's positional range (1e-4 <= |x| < 1e16, or zero) that is max(1, d) with
// d the smallest number of decimals at which CPython's round() reproduces x: repr() prints the shortest digit string that
// reads back as x, and the decimal nearest to x with d places reads back as x exactly when some d-place decimal does
// (tests/test_dataprep_cpu.py holds this to str() on millions of values). A Python int prints without a '.', i.e. 0.
// Values repr() prints in exponent notation are left to the caller (returns -1): "1.5e-05".split('.')[-1] is "5e-05".
constexpr int AUTO_MAX_DECIMALS = 22;
struct RangeAutoParams {
  int has_whole_digits;
  double lower, upper;
  Pow10 P[AUTO_MAX_DECIMALS + 1];             // 10^d
  double upper_clamp[AUTO_MAX_DECIMALS + 1];  // 10**nwd - 10**(-d) (d > 0), 10**nwd - 1 (d = 0): Python's own expressions
};

TA_HD int repr_decimals(double x, const Pow10* P) {
  const double ax = abs_of(x);
  if (ax == 0.0) return 1;                                  // "0.0", "-0.0"
  if (!(ax >= 1e-4 && ax < 1e16)) return -1;                 // exponent notation (or not finite): the caller's business
  for (int d = 0; d <= AUTO_MAX_DECIMALS; ++d)
    if (bits_of(py_round(x, P[d])) == bits_of(x)) return d < 1 ? 1 : d;
  return -1;
}

// One element with decimal_places = None. Returns NaN (and *flag = 1) where the element needs the host's str().
TA_HD double range_element_auto(double x, bool is_int, const RangeAutoParams& rp, const double* __restrict__ thresholds,
                                const double* __restrict__ factors, int* flag) {
  const int d = is_int ? 0 : repr_decimals(x, rp.P);
  if (d < 0) { *flag = 1; return double_of(0x7ff8000000000000ull); }
  double f = 1.0;
  if (rp.has_whole_digits) {
    const int p = (x == 0.0) ? 0 : floor_log10(abs_of(x), thresholds);
    f = factors[p - P10_MIN];
    if (!is_finite(f)) return f;
  }
  double scaled = (f != 0.0) ? py_round(mul_rn(x, f), rp.P[d]) : 0.0;
  if (rp.has_whole_digits) {
    double a = abs_of(scaled);
    if (a < rp.lower && a > 0.0) a = rp.lower;
    if (a >= rp.upper) a = rp.upper_clamp[d];
    scaled = (x < 0.0) ? -a : a;
  }
  return scaled;
}

// One element of convert_to_percent_changes (data_utils.py:649-657 / file_cache.py:362-377); prev != 0.
// An unchanged value is common (prices quoted in cents) and a zero numerator sends the device's division down its slow
// path for the whole warp (ncu: 59 of 177 instructions per row), so it is answered directly: (0.0 / prev) * 100 is a zero
// of prev's sign, and round() keeps it.
TA_HD double percent_change(double cur, double prev, const Pow10& P) {
  const double d = sub_rn(cur, prev);
  if (d == 0.0) return double_of(bits_of(prev) & 0x8000000000000000ull);
  return py_round(mul_rn(div_rn(d, prev), 100.0), P);
}

// Group of one value in bin_numeric_data (data_utils.py:527-557): pos = [0, b_1 .. b_G], neg = [-b_G .. -b_1, 0].
// The reference's linear search for the first j with bound[j] <= v < bound[j+1], with its for/else edge groups.
TA_HD int bin_of_linear(double v, const double* pos, const double* neg, int G) {
  if (v > 0.0) {
    for (int j = 0; j < G; ++j)
      if (v >= pos[j] && v < pos[j + 1]) return j + 1;
    return G;
  }
  if (v == 0.0) return 0;
  for (int j = 0; j < G; ++j)
    if (v >= neg[j] && v < neg[j + 1]) return -(G - j);
  return -G;
}

// count of bounds[0..n) that are <= v (bounds non-decreasing)
TA_HD int count_le(const double* bounds, int n, double v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (bounds[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// The same group when the boundaries are non-decreasing (the host checks; they are whenever pow() is monotone): the
// first interval that holds v is the one after the last boundary <= v, so the search is a bisection.
TA_HD int bin_of_sorted(double v, const double* pos, const double* neg, int G) {
  if (v > 0.0) {
    const int j = count_le(pos + 1, G, v);          // v > 0 = pos[0]
    return j < G ? j + 1 : G;
  }
  if (v == 0.0) return 0;
  if (!(v >= neg[0])) return -G;                      // below the most negative boundary (or NaN)
  const int j = count_le(neg, G + 1, v) - 1;          // v < 0 = neg[G]  =>  j <= G - 1
  return -(G - j);
}

TA_HD int bin_of(double v, const double* pos, const double* neg, int G, bool sorted) {
  return sorted ? bin_of_sorted(v, pos, neg, G) : bin_of_linear(v, pos, neg, G);
}

// Order-preserving map of finite doubles to unsigned integers (for the radix select) and back.
TA_HD u64 order_key(double x) {
  const u64 b = bits_of(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
TA_HD double order_value(u64 k) { return double_of((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k); }

// Vocabulary key: Python's set() holds one element for 0.0 and -0.0 (equal, same hash).
TA_HD u64 vocab_key(double x) { return x == 0.0 ? 0ull : bits_of(x); }
TA_HD u64 hash64(u64 k) {   // splitmix64 finaliser
  k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ull;
  k ^= k >> 27; k *= 0x94d049bb133111ebull;
  k ^= k >> 31;
  return k;
}

}  // namespace prep
}  // namespace ta
