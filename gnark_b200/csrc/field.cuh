// Prime-field arithmetic in Montgomery form, 32-bit limbs, for sm_100a.
//
// Memory layout is gnark-crypto's fp.Element / fr.Element bit for bit
// ([Limbs]uint64 little-endian, value x*R mod q, R = 2^(64*Limbs); SURVEY.md
// Appendix A, evidence backend/accelerated/icicle/groth16/bn254/icicle.go:119-130),
// so proving-key tables and witness vectors are used as uploaded: the
// reference's FromMontgomery / AffineFromMontgomery passes (icicle.go:121,322,350,
// 1008,1028,1452,1480) have no counterpart here.
//
// Multiplication: operand-scanning Montgomery product with TWO interleaved
// accumulators so that every 32x32 partial product (lo word, hi word) lands in a
// carry chain that visits each accumulator word exactly once:
//   row i adds a*b[i] and m_i*p;  products whose start position i+j is even go to
//   accumulator E, odd to O; both are indexed by ABSOLUTE word position, and the
//   word at position i of the "other" accumulator is folded in before m_i is
//   computed.  4N^2+N multiplier ops, no carry ripple, result < 2p, one
//   conditional subtraction.  (No tensor cores: integer path.)
#pragma once
#include "ptx.cuh"
#include "params_gen.cuh"

namespace gb200 {

// ---------------------------------------------------------------------------
// raw limb routines
// ---------------------------------------------------------------------------

// acc[base + j], acc[base + j + 1] += lo/hi(s[j] * m) for j = J0, J0+2, ... < N, one carry
// chain; the carry out of the chain is added to the next (fresh or carry-only) word.
template <int N, int J0, bool CARRY_IN>
HD void mad_chain(uint32_t* acc, int base, const uint32_t* s, uint32_t m) {
#pragma unroll
  for (int j = J0; j < N; j += 2) {
    if (j == J0 && !CARRY_IN)
      acc[base + j] = ptx::mad_lo_cc(s[j], m, acc[base + j]);
    else
      acc[base + j] = ptx::madc_lo_cc(s[j], m, acc[base + j]);
    acc[base + j + 1] = ptx::madc_hi_cc(s[j], m, acc[base + j + 1]);
  }
  constexpr int JL = J0 + 2 * ((N - 1 - J0) / 2);  // last j visited
  acc[base + JL + 2] = ptx::addc(acc[base + JL + 2], 0);
}

// r = a * b * R^-1 mod p, inputs < p (or < 2p with p having >= 2 spare bits), output < p
template <class P>
HD void mont_mul_raw(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  static_assert(N % 2 == 0, "even limb count expected");
  uint32_t acc[2][2 * N + 3];
  uint32_t p[N];
#pragma unroll
  for (int k = 0; k < N; k++) p[k] = P::mod(k);
#pragma unroll
  for (int k = 0; k < 2 * N + 3; k++) { acc[0][k] = 0; acc[1][k] = 0; }

#pragma unroll
  for (int i = 0; i < N; i++) {
    uint32_t* X = acc[i & 1];        // chain starting at absolute position i
    uint32_t* Y = acc[(i & 1) ^ 1];  // chain starting at absolute position i+1
    const uint32_t bi = b[i];
    if (i > 0) {
      X[i] = ptx::add_cc(X[i], Y[i]);             // fold; carry goes into the Y chain
      mad_chain<N, 1, true>(Y, i, a, bi);         // odd limbs of a: positions i+1 .. i+N
    } else {
      mad_chain<N, 1, false>(Y, i, a, bi);
    }
    mad_chain<N, 0, false>(X, i, a, bi);          // even limbs of a: positions i .. i+N-1
    const uint32_t m = X[i] * P::INV;
    mad_chain<N, 0, false>(X, i, p, m);           // X[i] becomes 0
    mad_chain<N, 1, false>(Y, i, p, m);
  }
  // result words are positions N .. 2N-1 of E + O (value < 2p < 2^(32N): higher words cancel)
  uint32_t s[N];
  s[0] = ptx::add_cc(acc[0][N], acc[1][N]);
#pragma unroll
  for (int k = 1; k < N; k++) s[k] = ptx::addc_cc(acc[0][N + k], acc[1][N + k]);
  // conditional subtract
  uint32_t d[N];
  d[0] = ptx::sub_cc(s[0], p[0]);
#pragma unroll
  for (int k = 1; k < N; k++) d[k] = ptx::subc_cc(s[k], p[k]);
  const uint32_t borrow = ptx::subc(0, 0);  // 0xffffffff if s < p
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = borrow ? s[k] : d[k];
}

// ---------------------------------------------------------------------------
// Wide (unreduced) products and a separate Montgomery reduction: the building blocks of the dedicated squaring
// (GB200_MONT_SQR) and of the lazily reduced Fp2 product (GB200_FP2_LAZY).  Compile-time options, off by default;
// bit-exact against the fused product (tests/test_emulation.py::test_wide_arithmetic and the *_opt emulation runs).
// ---------------------------------------------------------------------------

// t[0..2N) = a * b for ANY N-limb a, b (no reduction).  Same even/odd carry chains as mont_mul_raw without the m*p rows.
template <int N>
HD void wide_mul_raw(uint32_t* t, const uint32_t* a, const uint32_t* b) {
  uint32_t acc[2][2 * N + 3];
#pragma unroll
  for (int k = 0; k < 2 * N + 3; k++) { acc[0][k] = 0; acc[1][k] = 0; }
#pragma unroll
  for (int i = 0; i < N; i++) {
    mad_chain<N, 0, false>(acc[i & 1], i, a, b[i]);          // products starting at even+i positions
    mad_chain<N, 1, false>(acc[(i & 1) ^ 1], i, a, b[i]);
  }
  t[0] = ptx::add_cc(acc[0][0], acc[1][0]);
#pragma unroll
  for (int k = 1; k < 2 * N; k++) t[k] = ptx::addc_cc(acc[0][k], acc[1][k]);
  (void)ptx::addc(0, 0);
}

// Karatsuba on top of wide_mul_raw (GB200_MONT_KARATSUBA, large fields): one level for N = 12 (3 x 6-limb products:
// 108 multiplier operations instead of 144), two levels for N = 24 (9 x 6-limb: 324 instead of 576).
// -DGB200_KARATSUBA_MIN_LIMBS=8 extends it to the 8-limb fields (BN254: 3 x 4-limb = 48 instead of 64, so a Montgomery
// product is 48 + 72 = 120 multiplier operations instead of 136; the 12 / 24-limb decomposition is unchanged).
#ifndef GB200_KARATSUBA_MIN_LIMBS
#define GB200_KARATSUBA_MIN_LIMBS 12
#endif
//   a = aL + aH B, b = bL + bH B (B = 2^(32 N/2)):  a b = z0 + (z1 - z0 - z2) B + z2 B^2,
//   z0 = aL bL, z2 = aH bH, z1 = (aL + aH)(bL + bH) with the two carry bits of the sums handled apart.
template <int N>
HD void wide_mul_karatsuba(uint32_t* t, const uint32_t* a, const uint32_t* b) {
  if constexpr (N < GB200_KARATSUBA_MIN_LIMBS || (N % 2) != 0) {
    wide_mul_raw<N>(t, a, b);
  } else {
    constexpr int H = N / 2;
    uint32_t z0[2 * H], z2[2 * H], z1[2 * H + 2], sa[H], sb[H];
    wide_mul_karatsuba<H>(z0, a, b);
    wide_mul_karatsuba<H>(z2, a + H, b + H);
    // sums with their carry bits
    sa[0] = ptx::add_cc(a[0], a[H]);
#pragma unroll
    for (int k = 1; k < H; k++) sa[k] = ptx::addc_cc(a[k], a[H + k]);
    const uint32_t ca = ptx::addc(0, 0);
    sb[0] = ptx::add_cc(b[0], b[H]);
#pragma unroll
    for (int k = 1; k < H; k++) sb[k] = ptx::addc_cc(b[k], b[H + k]);
    const uint32_t cb = ptx::addc(0, 0);
    wide_mul_karatsuba<H>(z1, sa, sb);
    z1[2 * H] = 0; z1[2 * H + 1] = 0;
    // + ca * sb * 2^(32H) + cb * sa * 2^(32H) + ca cb 2^(64H)   (masks, no multiplications)
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    z1[H] = ptx::add_cc(z1[H], sb[0] & ma);
#pragma unroll
    for (int k = 1; k < H; k++) z1[H + k] = ptx::addc_cc(z1[H + k], sb[k] & ma);
    z1[2 * H] = ptx::addc(z1[2 * H], 0);
    z1[H] = ptx::add_cc(z1[H], sa[0] & mb);
#pragma unroll
    for (int k = 1; k < H; k++) z1[H + k] = ptx::addc_cc(z1[H + k], sa[k] & mb);
    z1[2 * H] = ptx::addc_cc(z1[2 * H], ca & cb);
    z1[2 * H + 1] = ptx::addc(z1[2 * H + 1], 0);
    // z1 -= z0 + z2   (the true middle term fits 2H + 1 limbs)
    z1[0] = ptx::sub_cc(z1[0], z0[0]);
#pragma unroll
    for (int k = 1; k < 2 * H; k++) z1[k] = ptx::subc_cc(z1[k], z0[k]);
    z1[2 * H] = ptx::subc_cc(z1[2 * H], 0);
    z1[2 * H + 1] = ptx::subc(z1[2 * H + 1], 0);
    z1[0] = ptx::sub_cc(z1[0], z2[0]);
#pragma unroll
    for (int k = 1; k < 2 * H; k++) z1[k] = ptx::subc_cc(z1[k], z2[k]);
    z1[2 * H] = ptx::subc_cc(z1[2 * H], 0);
    z1[2 * H + 1] = ptx::subc(z1[2 * H + 1], 0);
    // assemble: t = z0 + z1 B + z2 B^2
#pragma unroll
    for (int k = 0; k < 2 * H; k++) { t[k] = z0[k]; t[2 * H + k] = z2[k]; }
    t[H] = ptx::add_cc(t[H], z1[0]);
#pragma unroll
    for (int k = 1; k < 2 * H + 2 && H + k < 2 * N; k++) t[H + k] = ptx::addc_cc(t[H + k], z1[k]);
#pragma unroll
    for (int k = 3 * H + 2; k < 2 * N; k++) t[k] = ptx::addc_cc(t[k], 0);
    (void)ptx::addc(0, 0);
  }
}

// cross terms of a square: chains of a_i * a_j, j > i, all j of one parity
template <int N, int I>
struct SqrCross {
  HD static void run(uint32_t (*acc)[2 * N + 3], const uint32_t* a) {
    if constexpr (I + 1 < N) mad_chain<N, I + 1, false>(acc[1], I, a, a[I]);   // positions 2I+1, 2I+3, ... (odd)
    if constexpr (I + 2 < N) mad_chain<N, I + 2, false>(acc[0], I, a, a[I]);   // positions 2I+2, 2I+4, ... (even)
    if constexpr (I + 2 < N) SqrCross<N, I + 1>::run(acc, a);
  }
};

// t[0..2N) = a * a:  2 * sum_{i<j} a_i a_j 2^(32(i+j)) + sum_i a_i^2 2^(64 i)   (N(N-1)/2 + N products instead of N^2)
template <int N>
HD void wide_sqr_raw(uint32_t* t, const uint32_t* a) {
  uint32_t acc[2][2 * N + 3];
#pragma unroll
  for (int k = 0; k < 2 * N + 3; k++) { acc[0][k] = 0; acc[1][k] = 0; }
  SqrCross<N, 0>::run(acc, a);
  uint32_t c[2 * N];
  c[0] = ptx::add_cc(acc[0][0], acc[1][0]);
#pragma unroll
  for (int k = 1; k < 2 * N; k++) c[k] = ptx::addc_cc(acc[0][k], acc[1][k]);
  (void)ptx::addc(0, 0);
  // double (the cross sum is < 2^(64N - 1))
#pragma unroll
  for (int k = 2 * N - 1; k > 0; k--) c[k] = (c[k] << 1) | (c[k - 1] >> 31);
  c[0] <<= 1;
  // add the diagonal a_i^2 at words 2i, 2i+1
  t[0] = ptx::mad_lo_cc(a[0], a[0], c[0]);
  t[1] = ptx::madc_hi_cc(a[0], a[0], c[1]);
#pragma unroll
  for (int i = 1; i < N; i++) {
    t[2 * i] = ptx::madc_lo_cc(a[i], a[i], c[2 * i]);
    t[2 * i + 1] = ptx::madc_hi_cc(a[i], a[i], c[2 * i + 1]);
  }
  (void)ptx::addc(0, 0);
}

// r = T * R^-1 mod p for a 2N-limb T < p * R; output < p.  The m*p rows of mont_mul_raw on their own.
template <class P>
HD void mont_reduce_wide(uint32_t* r, const uint32_t* T) {
  constexpr int N = P::N;
  uint32_t acc[2][2 * N + 3];
  uint32_t p[N];
#pragma unroll
  for (int k = 0; k < N; k++) p[k] = P::mod(k);
#pragma unroll
  for (int k = 0; k < 2 * N + 3; k++) { acc[0][k] = k < 2 * N ? T[k] : 0; acc[1][k] = 0; }
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint32_t* X = acc[i & 1];
    uint32_t* Y = acc[(i & 1) ^ 1];
    // m from the folded low word, computed before the fold so that no carry is live across plain code
    const uint32_t m = (X[i] + Y[i]) * P::INV;
    if (i > 0) {
      X[i] = ptx::add_cc(X[i], Y[i]);             // fold; the carry goes into the Y chain
      mad_chain<N, 1, true>(Y, i, p, m);
    } else {
      mad_chain<N, 1, false>(Y, i, p, m);
    }
    mad_chain<N, 0, false>(X, i, p, m);           // X[i] becomes 0
  }
  uint32_t s[N];
  s[0] = ptx::add_cc(acc[0][N], acc[1][N]);
#pragma unroll
  for (int k = 1; k < N; k++) s[k] = ptx::addc_cc(acc[0][N + k], acc[1][N + k]);
  uint32_t d[N];
  d[0] = ptx::sub_cc(s[0], p[0]);
#pragma unroll
  for (int k = 1; k < N; k++) d[k] = ptx::subc_cc(s[k], p[k]);
  const uint32_t borrow = ptx::subc(0, 0);
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = borrow ? s[k] : d[k];
}

// multi-limb helpers without reduction (K limbs)
template <int K>
HD void limbs_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = ptx::add_cc(a[0], b[0]);
#pragma unroll
  for (int k = 1; k < K; k++) r[k] = ptx::addc_cc(a[k], b[k]);
  (void)ptx::addc(0, 0);
}
template <int K>
HD void limbs_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = ptx::sub_cc(a[0], b[0]);
#pragma unroll
  for (int k = 1; k < K; k++) r[k] = ptx::subc_cc(a[k], b[k]);
  (void)ptx::subc(0, 0);
}

template <class P>
HD void mod_add_raw(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  uint32_t s[N], d[N];
  s[0] = ptx::add_cc(a[0], b[0]);
#pragma unroll
  for (int k = 1; k < N; k++) s[k] = ptx::addc_cc(a[k], b[k]);
  d[0] = ptx::sub_cc(s[0], P::mod(0));
#pragma unroll
  for (int k = 1; k < N; k++) d[k] = ptx::subc_cc(s[k], P::mod(k));
  const uint32_t borrow = ptx::subc(0, 0);
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = borrow ? s[k] : d[k];
}

template <class P>
HD void mod_sub_raw(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  uint32_t d[N];
  d[0] = ptx::sub_cc(a[0], b[0]);
#pragma unroll
  for (int k = 1; k < N; k++) d[k] = ptx::subc_cc(a[k], b[k]);
  const uint32_t borrow = ptx::subc(0, 0);  // all ones if a < b
  r[0] = ptx::add_cc(d[0], P::mod(0) & borrow);
#pragma unroll
  for (int k = 1; k < N; k++) r[k] = ptx::addc_cc(d[k], P::mod(k) & borrow);
}

// ---------------------------------------------------------------------------
// Fp<P>: value type
// ---------------------------------------------------------------------------
template <class P>
struct alignas(16) Fp {
  static constexpr int N = P::N;
  static constexpr int DEGREE = 1;
  using Params = P;
  using Base = Fp<P>;
  uint32_t l[N];

  HD static Fp zero() { Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
  HD static Fp one() { Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = P::r1(i); return r; }
  HD static Fp r2() { Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = P::r2(i); return r; }

  HD bool is_zero() const { uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= l[i]; return t == 0; }
  HD bool operator==(const Fp& o) const { uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= (l[i] ^ o.l[i]); return t == 0; }
  HD bool operator!=(const Fp& o) const { return !(*this == o); }

  HD friend Fp operator+(const Fp& a, const Fp& b) { Fp r; mod_add_raw<P>(r.l, a.l, b.l); return r; }
  HD friend Fp operator-(const Fp& a, const Fp& b) { Fp r; mod_sub_raw<P>(r.l, a.l, b.l); return r; }
#if defined(GB200_MONT_KARATSUBA)
  // large fields: Karatsuba product + separate reduction (fewer multiplier operations than the fused product)
  HD static void mul_dispatch(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    if constexpr (N >= GB200_KARATSUBA_MIN_LIMBS) { uint32_t t[2 * N]; wide_mul_karatsuba<N>(t, a, b); mont_reduce_wide<P>(r, t); }
    else mont_mul_raw<P>(r, a, b);
  }
#else
  HD static void mul_dispatch(uint32_t* r, const uint32_t* a, const uint32_t* b) { mont_mul_raw<P>(r, a, b); }
#endif
#ifndef GB200_INLINE_LIMBS
#define GB200_INLINE_LIMBS 8      // fields of up to this many 32-bit limbs get the product inlined at every use
#endif
#if defined(__CUDA_ARCH__)
  // large fields: keep one copy of the unrolled product per kernel (I-cache, compile time)
#if defined(GB200_CALL_BYVAL)
  // operands and result of the out-of-line product travel in registers: a reference parameter forces the caller to
  // spill both operands to its stack frame and the callee to load them back (checked in SASS: 0 LDL/STL by value)
  static __device__ __noinline__ Fp mul_ni(Fp a, Fp b) { Fp r; mul_dispatch(r.l, a.l, b.l); return r; }
#else
  static __device__ __noinline__ Fp mul_ni(const Fp& a, const Fp& b) { Fp r; mul_dispatch(r.l, a.l, b.l); return r; }
#endif
  HD friend Fp operator*(const Fp& a, const Fp& b) {
    if (N > GB200_INLINE_LIMBS) return mul_ni(a, b);
    Fp r; mul_dispatch(r.l, a.l, b.l); return r;
  }
#else
  HD friend Fp operator*(const Fp& a, const Fp& b) { Fp r; mul_dispatch(r.l, a.l, b.l); return r; }
#endif
#if defined(GB200_MONT_SQR)
  HD Fp sqr() const {
    uint32_t t[2 * N];
    wide_sqr_raw<N>(t, l);
    Fp r; mont_reduce_wide<P>(r.l, t); return r;
  }
#else
  HD Fp sqr() const { return (*this) * (*this); }
#endif
  // a*b - c*d with ONE Montgomery reduction: a b + (p^2 - c d) < 2 p^2 < p R (top bit of p clear), reduced once.
  // Used by the XYZZ mixed addition under GB200_XYZZ_LAZY (Y3 = R (Q - X3) - Y1 PPP): one reduction (N^2 + N
  // multiplier operations) less per addition.  Bit-exact with a*b - c*d (emulation: test_wide_arithmetic).
  HD static Fp mul_sub(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
    uint32_t t0[2 * N], t1[2 * N], q[2 * N];
#if defined(GB200_MONT_KARATSUBA)
    wide_mul_karatsuba<N>(t0, a.l, b.l);
    wide_mul_karatsuba<N>(t1, c.l, d.l);
#else
    wide_mul_raw<N>(t0, a.l, b.l);
    wide_mul_raw<N>(t1, c.l, d.l);
#endif
#pragma unroll
    for (int k = 0; k < 2 * N; k++) q[k] = P::psq(k);
    limbs_sub<2 * N>(q, q, t1);
    limbs_add<2 * N>(t0, t0, q);
    Fp r; mont_reduce_wide<P>(r.l, t0); return r;
  }
  HD Fp neg() const { return is_zero() ? *this : (zero() - *this); }
  HD Fp dbl() const { return *this + *this; }
  // Montgomery -> canonical (multiply by 1) and back
  HD Fp from_mont() const { Fp o = zero(); o.l[0] = 1; return (*this) * o; }
  HD Fp to_mont() const { return (*this) * r2(); }

  // a^(p-2) (only used a handful of times per MSM / for batch inversion seeds)
  HD Fp inverse() const {
    Fp result = one();
    Fp base = *this;
    for (int w = 0; w < N; w++) {
      uint32_t e = P::pm2(w);
      // constant-time-ish square and multiply, LSB first
      for (int bit = 0; bit < 32; bit++) {
        if ((e >> bit) & 1) result = result * base;
        base = base.sqr();
      }
    }
    return result;
  }
  // Same result as inverse() by the binary extended Euclidean algorithm (Kaliski's "almost Montgomery inverse"):
  // shifts, additions and subtractions only, on the ALU pipe instead of ~1.5*BITS Montgomery products on the
  // multiplier pipe.  Phase 1 gives x^-1 * 2^k mod p (BITS <= k <= 2*BITS) for the stored value x = a*R;
  // a^-1 * R = x^-1 * R^2 = x^-1 * 2^(64N), so phase 2 is 64N - k modular doublings.  Needs the top bit of the top
  // limb of p clear (r, s < 2p must fit N limbs): true for every modulus here.
  // Kaliski's single-bit cases are merged so that a warp does not diverge: with u, v both odd and (u, r) kept as
  // the larger pair (conditional swap), one iteration is  u -= v; r += s; t = ctz(u); u >>= t; s <<= t; k += t
  // (his "u > v" step followed by t-1 "u even" steps), about 0.8*BITS iterations in all.
  HD Fp inverse_gcd() const {
    if (is_zero()) return zero();
    uint32_t u[N], v[N], r[N], s[N];
#pragma unroll
    for (int i = 0; i < N; i++) { u[i] = P::mod(i); v[i] = l[i]; r[i] = 0; s[i] = 0; }
    s[0] = 1;
    auto ctz32 = [](uint32_t x) -> int {
#if defined(__CUDA_ARCH__)
      return __ffs((int)x) - 1;
#else
      return __builtin_ctz(x);
#endif
    };
    auto shr = [](uint32_t* a, int t) {       // 0 < t < 32
#pragma unroll
      for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> t) | (a[i + 1] << (32 - t));
      a[N - 1] >>= t;
    };
    auto shl = [](uint32_t* a, int t) {       // 0 < t < 32
#pragma unroll
      for (int i = N - 1; i > 0; i--) a[i] = (a[i] << t) | (a[i - 1] >> (32 - t));
      a[0] <<= t;
    };
    auto shr_limb = [](uint32_t* a) {
#pragma unroll
      for (int i = 0; i < N - 1; i++) a[i] = a[i + 1];
      a[N - 1] = 0;
    };
    auto shl_limb = [](uint32_t* a) {
#pragma unroll
      for (int i = N - 1; i > 0; i--) a[i] = a[i - 1];
      a[0] = 0;
    };
    auto add = [](uint32_t* a, const uint32_t* b) {
      uint64_t c = 0;
#pragma unroll
      for (int i = 0; i < N; i++) { c += (uint64_t)a[i] + b[i]; a[i] = (uint32_t)c; c >>= 32; }
    };
    auto sub = [](uint32_t* a, const uint32_t* b) {
      uint64_t br = 0;
#pragma unroll
      for (int i = 0; i < N; i++) { const uint64_t d = (uint64_t)a[i] - b[i] - br; a[i] = (uint32_t)d; br = (d >> 32) & 1; }
    };
    // -1 / 0 / +1 for a < b / a == b / a > b, without data-dependent branches
    auto cmp = [](const uint32_t* a, const uint32_t* b) -> int {
      int res = 0;
#pragma unroll
      for (int i = 0; i < N; i++) res = a[i] != b[i] ? (a[i] > b[i] ? 1 : -1) : res;
      return res;
    };
    int k = 0;
    // x is non-zero: strip its trailing zeros (Kaliski's "v even" steps; r = 0 so r <<= t is a no-op)
    while (v[0] == 0) { shr_limb(v); k += 32; }
    { const int t = ctz32(v[0]); if (t) { shr(v, t); k += t; } }
    bool swapped = false;                    // false: (u, r) / (v, s) are Kaliski's pairs; true: exchanged
    for (;;) {
      const int c = cmp(u, v);
      if (c == 0) break;
      if (c < 0) {                           // keep u > v
#pragma unroll
        for (int i = 0; i < N; i++) {
          const uint32_t tu = u[i]; u[i] = v[i]; v[i] = tu;
          const uint32_t tr = r[i]; r[i] = s[i]; s[i] = tr;
        }
        swapped = !swapped;
      }
      sub(u, v);
      add(r, s);
      while (u[0] == 0) { shr_limb(u); shl_limb(s); k += 32; }
      const int t = ctz32(u[0]);             // u - v is even and non-zero: t >= 1 unless whole limbs were stripped
      if (t) { shr(u, t); shl(s, t); k += t; }
    }
    // u == v (== gcd = 1): Kaliski's last step  v = 0, s += r, r = 2r, k++  ->  result r
    uint32_t* rr = swapped ? s : r;
    shl(rr, 1);
    k++;
    uint32_t pm[N];
#pragma unroll
    for (int i = 0; i < N; i++) pm[i] = P::mod(i);
    if (cmp(rr, pm) >= 0) sub(rr, pm);
    sub(pm, rr);                             // p - r = x^-1 * 2^k mod p
    Fp out;
#pragma unroll
    for (int i = 0; i < N; i++) out.l[i] = pm[i];
    for (; k < 64 * N; k++) out = out.dbl();
    return out;
  }
  // multiply by a small unsigned constant via additions
  HD Fp mul_small(unsigned k) const {
    Fp acc = zero();
    Fp cur = *this;
    while (k) {
      if (k & 1) acc = acc + cur;
      cur = cur.dbl();
      k >>= 1;
    }
    return acc;
  }
};

// runtime-index pm2 lookup needs a real array on device; provide via switch-free table
// (P::pm2(w) with runtime w compiles to a local constant array; fine for the rare inverse)

// ---------------------------------------------------------------------------
// Fp2<F, BETA>: F[u]/(u^2 + BETA), i.e. u^2 = -BETA  (BETA = 1: BN254, BLS12-381; 5: BLS12-377)
// rule restated from std/algebra/emulated/fields_bn254/e2.go:203-213 and
// std/algebra/native/fields_bls12377/e2.go:134.  Memory = gnark E2{A0, A1}.
// ---------------------------------------------------------------------------
// Fp2 product / square: out of line by default (one copy of the 3-multiplication body per kernel);
// -DGB200_INLINE_FP2 inlines them at every use (A/B knob: no call, no stack round trip, larger code)
#if defined(GB200_INLINE_FP2)
#define HD_FP2 HD
#else
#define HD_FP2 HDNI
#endif

template <class T> struct is_device_fp { static constexpr bool value = false; };
template <class P> struct is_device_fp<Fp<P>> { static constexpr bool value = true; };

template <class F, unsigned BETA>
struct alignas(16) Fp2 {
  static constexpr int DEGREE = 2;
  using Base = F;
  F a0, a1;
  HD static Fp2 zero() { Fp2 r; r.a0 = F::zero(); r.a1 = F::zero(); return r; }
  HD static Fp2 one() { Fp2 r; r.a0 = F::one(); r.a1 = F::zero(); return r; }
  HD bool is_zero() const { return a0.is_zero() && a1.is_zero(); }
  HD bool operator==(const Fp2& o) const { return a0 == o.a0 && a1 == o.a1; }
  HD bool operator!=(const Fp2& o) const { return !(*this == o); }
  HD friend Fp2 operator+(const Fp2& x, const Fp2& y) { Fp2 r; r.a0 = x.a0 + y.a0; r.a1 = x.a1 + y.a1; return r; }
  HD friend Fp2 operator-(const Fp2& x, const Fp2& y) { Fp2 r; r.a0 = x.a0 - y.a0; r.a1 = x.a1 - y.a1; return r; }
  HD static F mul_beta(const F& t) { return BETA == 1 ? t : t.mul_small(BETA); }
  // not inlined on device: one copy of the 3-multiplication body per kernel
  // Karatsuba on UNREDUCED double-width products: 3 wide products and 2 Montgomery reductions instead of 3 of each.
  //   c1 = (a0+a1)(b0+b1) - a0b0 - a1b1   (>= 0, < 2p^2)
  //   c0 = a0b0 + BETA (p^2 - a1b1)       (>= 0, < (1+BETA) p^2 < pR for every modulus here)
  template <class FF = F>
  HD static Fp2 mul_lazy(const Fp2& x, const Fp2& y) {
    constexpr int N = FF::N;
    using P = typename FF::Params;
    uint32_t t0[2 * N], t1[2 * N], t2[2 * N], sa[N], sb[N];
    wide_mul_raw<N>(t0, x.a0.l, y.a0.l);
    wide_mul_raw<N>(t1, x.a1.l, y.a1.l);
    limbs_add<N>(sa, x.a0.l, x.a1.l);           // < 2p: fits N limbs (top bit of p clear)
    limbs_add<N>(sb, y.a0.l, y.a1.l);
    wide_mul_raw<N>(t2, sa, sb);
    limbs_sub<2 * N>(t2, t2, t0);
    limbs_sub<2 * N>(t2, t2, t1);
    uint32_t q[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) q[k] = P::psq(k);
    limbs_sub<2 * N>(q, q, t1);                 // p^2 - a1 b1
#pragma unroll
    for (unsigned k = 0; k < BETA; k++) limbs_add<2 * N>(t0, t0, q);
    Fp2 r;
    mont_reduce_wide<P>(r.a0.l, t0);
    mont_reduce_wide<P>(r.a1.l, t2);
    return r;
  }
#if defined(GB200_CALL_BYVAL) && defined(__CUDA_ARCH__)
  HD_FP2 static Fp2 mul(Fp2 x, Fp2 y) {
#else
  HD_FP2 static Fp2 mul(const Fp2& x, const Fp2& y) {
#endif
#if defined(GB200_FP2_LAZY)
    if constexpr (is_device_fp<F>::value) return mul_lazy<F>(x, y);
#endif
    // Karatsuba: 3 base multiplications
    F v0 = x.a0 * y.a0;
    F v1 = x.a1 * y.a1;
    F s = (x.a0 + x.a1) * (y.a0 + y.a1);
    Fp2 r;
    r.a0 = v0 - mul_beta(v1);
    r.a1 = s - v0 - v1;
    return r;
  }
  HD friend Fp2 operator*(const Fp2& x, const Fp2& y) { return mul(x, y); }
#if defined(GB200_CALL_BYVAL) && defined(__CUDA_ARCH__)
  HDNI static Fp2 sqr_ni(Fp2 x) {
    if (BETA == 1) {
      Fp2 r;
      F t = x.a0 * x.a1;
      r.a0 = (x.a0 + x.a1) * (x.a0 - x.a1);
      r.a1 = t + t;
      return r;
    }
    return mul(x, x);
  }
  HD Fp2 sqr() const { return sqr_ni(*this); }
#else
  HD_FP2 Fp2 sqr() const {
    if (BETA == 1) {  // (a0+a1)(a0-a1), 2 a0 a1
      Fp2 r;
      F t = a0 * a1;
      r.a0 = (a0 + a1) * (a0 - a1);
      r.a1 = t + t;
      return r;
    }
    return mul(*this, *this);
  }
#endif
  HD Fp2 neg() const { Fp2 r; r.a0 = a0.neg(); r.a1 = a1.neg(); return r; }
  HD Fp2 dbl() const { Fp2 r; r.a0 = a0.dbl(); r.a1 = a1.dbl(); return r; }
  HD Fp2 inverse() const {
    F n = a0.sqr() + mul_beta(a1.sqr());
    F ni = n.inverse();
    Fp2 r; r.a0 = a0 * ni; r.a1 = (a1 * ni).neg();
    return r;
  }
  HD Fp2 inverse_gcd() const {
    F n = a0.sqr() + mul_beta(a1.sqr());
    F ni = n.inverse_gcd();
    Fp2 r; r.a0 = a0 * ni; r.a1 = (a1 * ni).neg();
    return r;
  }
};

// 52-bit-limb (FP64 pipe) parameter pack of a base field, when one exists (field52.cuh)
template <class F> struct F52Traits { static constexpr bool ok = false; using P52 = void; };
template <> struct F52Traits<Fp<bn254_fp_params>> { static constexpr bool ok = true; using P52 = bn254_fp_params52; };
template <> struct F52Traits<Fp<bls12_381_fp_params>> { static constexpr bool ok = true; using P52 = bls12_381_fp_params52; };
template <> struct F52Traits<Fp<bls12_377_fp_params>> { static constexpr bool ok = true; using P52 = bls12_377_fp_params52; };
template <> struct F52Traits<Fp<bw6_761_fp_params>> { static constexpr bool ok = true; using P52 = bw6_761_fp_params52; };

// concrete fields
using bn254_fp = Fp<bn254_fp_params>;
using bn254_fr = Fp<bn254_fr_params>;
using bls12_381_fp = Fp<bls12_381_fp_params>;
using bls12_381_fr = Fp<bls12_381_fr_params>;
using bls12_377_fp = Fp<bls12_377_fp_params>;
using bls12_377_fr = Fp<bls12_377_fr_params>;
using bw6_761_fp = Fp<bw6_761_fp_params>;
using bw6_761_fr = Fp<bw6_761_fr_params>;
using bn254_fp2 = Fp2<bn254_fp, 1>;
using bls12_381_fp2 = Fp2<bls12_381_fp, 1>;
using bls12_377_fp2 = Fp2<bls12_377_fp, 5>;

}  // namespace gb200
