// Montgomery multiplication on the FP64 pipe: 52-bit limbs, every 52x52 -> 104-bit limb
// product obtained exactly from two DFMAs (round-toward-zero) and accumulated as raw bit
// patterns in 64-bit integer columns.
//
// Why: on B200 the integer path is bound by IMAD.WIDE issue (32 lane-products/clk/SM,
// profiles/r01_microbench_pipes.txt) and the 32-bit k_msm_accumulate runs that pipe at 86 %.
// DFMA issues at 64 lanes/clk/SM on a DIFFERENT pipe and yields 52x52-bit products: ~2.6x
// the product bits per issue slot.  Technique after Emmart, Zheng, Weems (2018), re-derived
// here; radix 2^52, R52 = 2^(52*L).
//
//   hi = fma_rz(a, b, 2^104)               = 2^104 + floor(ab / 2^52) * 2^52     (exact)
//   lo = fma_rz(a, b, (2^104 + 2^52) - hi) = 2^52 + (ab mod 2^52)               (exact)
// so bits(hi) = bits(2^104) + floor(ab/2^52) and bits(lo) = bits(2^52) + (ab mod 2^52);
// the constant exponent patterns are pre-subtracted from the column accumulators.
//
// Values are kept LAZILY reduced: with >= 6 spare bits (52*L - BITS) a product of inputs
// < 8p is < 2p (>= 5 spare bits: inputs < 4p), so additions / subtractions need no modular
// reduction, only carry normalisation.  Limbs entering a product must be normalised
// (0 <= limb < 2^52).  Element storage between operations: L signed 64-bit limbs.
#pragma once
#include <cstdint>

#include "ptx.cuh"

#ifndef __CUDA_ARCH__
#include <cmath>
#endif

namespace gb200 {

namespace f52 {
constexpr uint64_t B52 = 0x4330000000000000ull;   // bits(2^52)
constexpr uint64_t B104 = 0x4670000000000000ull;  // bits(2^104)
constexpr uint64_t M52 = (1ull << 52) - 1;
constexpr double TWO52 = 4503599627370496.0;
constexpr double TWO104 = 20282409603651670423947251286016.0;

HD double fma_rz(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rz(a, b, c);
#else
  return std::fma(a, b, c);  // host tests run under fesetround(FE_TOWARDZERO)
#endif
}
HD uint64_t bits(double x) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; __builtin_memcpy(&u, &x, 8); return u;
#endif
}
HD double from_bits(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double x; __builtin_memcpy(&x, &u, 8); return x;
#endif
}
// exact conversion of an integer 0 <= v < 2^52 to double without the (slow) I2F unit
HD double to_double(uint64_t v) { return from_bits(v | B52) - TWO52; }
// exact conversion of an integer-valued double 0 <= d < 2^52 to an integer without F2I
HD int64_t to_int(double d) { return (int64_t)(bits(d + TWO52) & M52); }
}  // namespace f52

// Element: L normalised limbs (int64, each in [0, 2^52)), value possibly >= p (lazy).
template <class P52>
struct F52 {
  static constexpr int L = P52::L;
  int64_t l[L];
  HD static F52 zero() { F52 r;
#pragma unroll
    for (int i = 0; i < L; i++) r.l[i] = 0; return r; }
  HD static F52 one() { F52 r;
#pragma unroll
    for (int i = 0; i < L; i++) r.l[i] = (int64_t)P52::one52(i); return r; }
  HD bool limbs_all_zero() const { int64_t t = 0;
#pragma unroll
    for (int i = 0; i < L; i++) t |= l[i]; return t == 0; }
};

// operands of a product: the same limbs as exact doubles
template <class P52>
struct D52 {
  static constexpr int L = P52::L;
  double d[L];
  HD explicit D52(const F52<P52>& x) {
#pragma unroll
    for (int i = 0; i < L; i++) d[i] = f52::to_double((uint64_t)x.l[i]); }
  HD D52() {}
};

// r = a * b / R52 (mod p): inputs < 8p (spare >= 6 bits) normalised, output < 2p normalised
template <class P52>
HD F52<P52> mul52(const D52<P52>& a, const D52<P52>& b) {
  constexpr int L = P52::L;
  constexpr double C1 = f52::TWO104;
  constexpr double C2 = f52::TWO104 + f52::TWO52;
  uint64_t acc[2 * L + 1];
  // pre-subtract the exponent patterns of every product that will be added (a*b and m*p have
  // the same column pattern): column k gets nlo(k) lo-terms and nhi(k) hi-terms, twice
#pragma unroll
  for (int k = 0; k <= 2 * L; k++) {
    int nlo = 0, nhi = 0;
    for (int i = 0; i < L; i++)
      for (int j = 0; j < L; j++) {
        if (i + j == k) nlo++;
        if (i + j + 1 == k) nhi++;
      }
    acc[k] = 0ull - 2ull * ((uint64_t)nlo * f52::B52 + (uint64_t)nhi * f52::B104);
  }
#pragma unroll
  for (int i = 0; i < L; i++)
#pragma unroll
    for (int j = 0; j < L; j++) {
      const double hi = f52::fma_rz(a.d[i], b.d[j], C1);
      const double lo = f52::fma_rz(a.d[i], b.d[j], C2 - hi);
      acc[i + j + 1] += f52::bits(hi);
      acc[i + j] += f52::bits(lo);
    }
  // Montgomery reduction, one 52-bit digit per step
  const double pinv = P52::pinv52();
#pragma unroll
  for (int i = 0; i < L; i++) {
    // when column i is read, the only product still owed to it is step i's own (j = 0) lo-term
    const uint64_t q = (acc[i] + f52::B52) & f52::M52;
    const double dq = f52::to_double(q);
    const double mh = f52::fma_rz(dq, pinv, C1);
    const double ml = f52::fma_rz(dq, pinv, C2 - mh);
    const double dm = ml - f52::TWO52;  // m = q * pinv mod 2^52 as an exact double
#pragma unroll
    for (int j = 0; j < L; j++) {
      const double pj = P52::mod52(j);
      const double hi = f52::fma_rz(dm, pj, C1);
      const double lo = f52::fma_rz(dm, pj, C2 - hi);
      acc[i + j + 1] += f52::bits(hi);
      acc[i + j] += f52::bits(lo);
    }
    acc[i + 1] += acc[i] >> 52;  // column i is complete and divisible by 2^52
  }
  F52<P52> r;
  uint64_t carry = 0;
#pragma unroll
  for (int k = 0; k < L; k++) {
    const uint64_t v = acc[L + k] + carry;
    r.l[k] = (int64_t)(v & f52::M52);
    carry = v >> 52;
  }
  return r;
}

template <class P52>
HD F52<P52> mul52(const F52<P52>& a, const F52<P52>& b) { return mul52<P52>(D52<P52>(a), D52<P52>(b)); }

// carry-normalise signed limb sums: value must be in [0, 2^(52L))
template <class P52>
HD void normalize52(F52<P52>& x) {
  constexpr int L = P52::L;
  int64_t carry = 0;
#pragma unroll
  for (int k = 0; k < L; k++) {
    const int64_t v = x.l[k] + carry;
    x.l[k] = v & (int64_t)f52::M52;
    carry = v >> 52;  // arithmetic shift: floor
  }
}

// r = a + b ; r = a - b + k*p   (limb-wise, then normalised; caller guarantees 0 <= result < 2^(52L))
template <class P52>
HD F52<P52> add52(const F52<P52>& a, const F52<P52>& b) {
  F52<P52> r;
#pragma unroll
  for (int i = 0; i < P52::L; i++) r.l[i] = a.l[i] + b.l[i];
  normalize52<P52>(r);
  return r;
}
template <class P52, int K>
HD F52<P52> sub52(const F52<P52>& a, const F52<P52>& b) {
  F52<P52> r;
#pragma unroll
  for (int i = 0; i < P52::L; i++) r.l[i] = a.l[i] - b.l[i] + (int64_t)K * (int64_t)P52::mod52(i);
  normalize52<P52>(r);
  return r;
}
// x < 2^k p  ->  x < ~ 2^(k-1) p : subtract 2^(k-1) p when the top limb says x >= 2^(k-1) p (never underflows)
template <class P52, int HALF_MULT>
HD void partial_reduce52(F52<P52>& x) {
  constexpr int L = P52::L;
  // top limb of HALF_MULT * p, rounded up: x.top > thr  =>  x > HALF_MULT*p
  const int64_t thr = (int64_t)((double)HALF_MULT * P52::mod52(L - 1)) + HALF_MULT;
  if (x.l[L - 1] > thr) {
#pragma unroll
    for (int i = 0; i < L; i++) x.l[i] -= (int64_t)HALF_MULT * (int64_t)P52::mod52(i);
    normalize52<P52>(x);
  }
}
// exact test "x == 0 (mod p)" for a normalised x < 2p
template <class P52>
HD bool is_zero_mod_p_lt2p(const F52<P52>& x) {
  int64_t z = 0, e = 0;
#pragma unroll
  for (int i = 0; i < P52::L; i++) { z |= x.l[i]; e |= x.l[i] ^ (int64_t)P52::mod52(i); }
  return z == 0 || e == 0;
}
// fully reduce a normalised x < 8p to [0, p)
template <class P52>
HD void canonical52(F52<P52>& x) {
  constexpr int L = P52::L;
#pragma unroll
  for (int round = 0; round < 3; round++) {
    const int mult = 4 >> round;  // 4p, 2p, p
    F52<P52> t;
#pragma unroll
    for (int i = 0; i < L; i++) t.l[i] = x.l[i] - (int64_t)mult * (int64_t)P52::mod52(i);
    int64_t carry = 0;
#pragma unroll
    for (int k = 0; k < L; k++) {
      const int64_t v = t.l[k] + carry;
      t.l[k] = v & (int64_t)f52::M52;
      carry = v >> 52;
    }
    if (carry >= 0) x = t;  // no borrow out of the top limb: x >= mult*p
  }
}

// ---- conversion between gnark's layout (N32 x 32-bit limbs, Montgomery R32 = 2^(32 N32)) and F52
// (Montgomery R52): bit repacking + one product by a constant
template <class P52>
HD F52<P52> repack_from_u32(const uint32_t* w) {  // integer value, no Montgomery change
  constexpr int L = P52::L, N = P52::N32;
  F52<P52> r;
#pragma unroll
  for (int k = 0; k < L; k++) {
    const int bit = 52 * k;
    uint64_t v = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {  // a 52-bit window spans at most 3 words
      const int wi = (bit >> 5) + j;
      const int sh = 32 * wi - bit;  // position of word wi relative to the window start
      if (wi < N && sh < 52) {
        v |= sh >= 0 ? ((uint64_t)w[wi] << sh) : ((uint64_t)w[wi] >> (-sh));
      }
    }
    r.l[k] = (int64_t)(v & f52::M52);
  }
  return r;
}
template <class P52>
HD void repack_to_u32(const F52<P52>& x, uint32_t* w) {  // x canonical (< p < 2^(32 N32))
  constexpr int L = P52::L, N = P52::N32;
#pragma unroll
  for (int wi = 0; wi < N; wi++) {
    const int bit = 32 * wi;
    uint64_t v = 0;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int k = bit / 52 + j;
      if (k < L) {
        const int sh = 52 * k - bit;
        v |= sh >= 0 ? ((uint64_t)x.l[k] << sh) : ((uint64_t)x.l[k] >> (-sh));
      }
    }
    w[wi] = (uint32_t)v;
  }
}
template <class P52>
HD F52<P52> const52(double (*f)(int)) { F52<P52> r;
#pragma unroll
  for (int i = 0; i < P52::L; i++) r.l[i] = (int64_t)f(i); return r; }

// x*R32 (gnark memory) -> x*R52, < 2p
template <class P52>
HD F52<P52> from_mont32(const uint32_t* w) {
  F52<P52> k;
#pragma unroll
  for (int i = 0; i < P52::L; i++) k.l[i] = (int64_t)P52::from_r32(i);
  return mul52<P52>(repack_from_u32<P52>(w), k);
}
// x*R52 (< 8p) -> x*R32 canonical, gnark memory
template <class P52>
HD void to_mont32(const F52<P52>& x, uint32_t* w) {
  F52<P52> k;
#pragma unroll
  for (int i = 0; i < P52::L; i++) k.l[i] = (int64_t)P52::to_r32(i);
  F52<P52> t = mul52<P52>(x, k);
  canonical52<P52>(t);
  repack_to_u32<P52>(t, w);
}

}  // namespace gb200
