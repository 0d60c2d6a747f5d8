// Host (CPU) Montgomery field with 64-bit limbs, same interface and memory layout
// as the device Fp<P>, so curve.cuh's group law can be instantiated on the host
// for Groth16 proof assembly - the few scalar multiplications / additions the
// reference also performs on the CPU, in Go (backend/groth16/bn254/prove.go:185,
// 199-200,212-214,241-269,287-292).  Serial, latency-bound work: a CPU core does a
// 254-bit scalar multiplication ~20x faster than one GPU thread.
#pragma once
#include <cstdint>
#include "params_gen.cuh"

namespace gb200 {

template <class P>
struct alignas(16) HFp {
  static constexpr int N = P::N;       // 32-bit limbs (layout compatible with Fp<P>)
  static constexpr int M = P::N / 2;   // 64-bit limbs
  static constexpr int DEGREE = 1;
  using Params = P;
  using Base = HFp<P>;
  uint64_t l[M];

  static uint64_t modl(int i) { return (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32); }
  static uint64_t inv64() {
    // -p^-1 mod 2^64 by Newton iteration from the 32-bit constant
    uint64_t p0 = modl(0);
    uint64_t x = (uint64_t)(0u - P::INV);  // p^-1 mod 2^32 (P::INV = -p^-1)
    x *= 2 - p0 * x;                        // now mod 2^64
    return 0 - x;
  }
  static HFp zero() { HFp r; for (int i = 0; i < M; i++) r.l[i] = 0; return r; }
  static HFp one() { HFp r; for (int i = 0; i < M; i++) r.l[i] = (uint64_t)P::r1(2 * i) | ((uint64_t)P::r1(2 * i + 1) << 32); return r; }
  static HFp r2() { HFp r; for (int i = 0; i < M; i++) r.l[i] = (uint64_t)P::r2(2 * i) | ((uint64_t)P::r2(2 * i + 1) << 32); return r; }
  bool is_zero() const { uint64_t t = 0; for (int i = 0; i < M; i++) t |= l[i]; return t == 0; }
  bool operator==(const HFp& o) const { uint64_t t = 0; for (int i = 0; i < M; i++) t |= l[i] ^ o.l[i]; return t == 0; }
  bool operator!=(const HFp& o) const { return !(*this == o); }

  static bool geq_mod(const uint64_t* a) {
    for (int i = M - 1; i >= 0; i--) {
      uint64_t m = modl(i);
      if (a[i] > m) return true;
      if (a[i] < m) return false;
    }
    return true;
  }
  static void sub_mod(uint64_t* a) {
    unsigned __int128 br = 0;
    for (int i = 0; i < M; i++) {
      unsigned __int128 t = (unsigned __int128)a[i] - modl(i) - br;
      a[i] = (uint64_t)t;
      br = (t >> 64) & 1;
    }
  }
  friend HFp operator+(const HFp& a, const HFp& b) {
    HFp r; unsigned __int128 c = 0;
    for (int i = 0; i < M; i++) { c += (unsigned __int128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (c || geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  friend HFp operator-(const HFp& a, const HFp& b) {
    HFp r; unsigned __int128 br = 0;
    for (int i = 0; i < M; i++) {
      unsigned __int128 t = (unsigned __int128)a.l[i] - b.l[i] - br;
      r.l[i] = (uint64_t)t; br = (t >> 64) & 1;
    }
    if (br) {
      unsigned __int128 c = 0;
      for (int i = 0; i < M; i++) { c += (unsigned __int128)r.l[i] + modl(i); r.l[i] = (uint64_t)c; c >>= 64; }
    }
    return r;
  }
  friend HFp operator*(const HFp& a, const HFp& b) {
    // CIOS
    static const uint64_t ninv = inv64();
    uint64_t t[M + 2];
    for (int i = 0; i < M + 2; i++) t[i] = 0;
    for (int i = 0; i < M; i++) {
      unsigned __int128 c = 0;
      for (int j = 0; j < M; j++) {
        c += (unsigned __int128)a.l[j] * b.l[i] + t[j];
        t[j] = (uint64_t)c; c >>= 64;
      }
      c += t[M]; t[M] = (uint64_t)c; t[M + 1] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * ninv;
      c = (unsigned __int128)m * modl(0) + t[0];
      c >>= 64;
      for (int j = 1; j < M; j++) {
        c += (unsigned __int128)m * modl(j) + t[j];
        t[j - 1] = (uint64_t)c; c >>= 64;
      }
      c += t[M]; t[M - 1] = (uint64_t)c; c >>= 64;
      t[M] = t[M + 1] + (uint64_t)c;
    }
    HFp r;
    for (int i = 0; i < M; i++) r.l[i] = t[i];
    if (t[M] || geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  HFp sqr() const { return (*this) * (*this); }
  HFp neg() const { return is_zero() ? *this : (zero() - *this); }
  HFp dbl() const { return *this + *this; }
  HFp from_mont() const { HFp o = zero(); o.l[0] = 1; return (*this) * o; }
  HFp to_mont() const { return (*this) * r2(); }
  HFp inverse() const {
    HFp result = one(), base = *this;
    for (int w = 0; w < N; w++) {
      uint32_t e = P::pm2(w);
      for (int bit = 0; bit < 32; bit++) {
        if ((e >> bit) & 1) result = result * base;
        base = base.sqr();
      }
    }
    return result;
  }
  HFp mul_small(unsigned k) const {
    HFp acc = zero(), cur = *this;
    while (k) { if (k & 1) acc = acc + cur; cur = cur.dbl(); k >>= 1; }
    return acc;
  }
};

}  // namespace gb200
