// Scalar-field arithmetic on the HOST with a run-time limb count, for the handful of field operations a prover
// does between device stages (PLONK: linearised-polynomial coefficients, blinding patches, support of the
// permutation - backend/plonk/bn254/prove.go:1366-1487, :1211-1220, setup.go:377-392).  Values are fr.Elements in
// gnark's memory layout: little-endian 64-bit limbs, Montgomery form, so everything read from or written to the
// device / the caller is used as is.  Plain C++ (CIOS Montgomery product on unsigned __int128); not a hot path.
#pragma once
#include <cstdint>
#include <cstring>

namespace gb200 {

constexpr int HOSTFR_MAX_LIMBS = 6;   // BW6-761 Fr: 377 bits

struct HostFr {
  uint64_t v[HOSTFR_MAX_LIMBS];
};

struct HostFrCtx {
  int L = 0;                          // 64-bit limbs
  uint64_t mod[HOSTFR_MAX_LIMBS] = {0}, one[HOSTFR_MAX_LIMBS] = {0}, r2[HOSTFR_MAX_LIMBS] = {0};
  uint64_t ninv = 0;                  // -mod^-1 mod 2^64
  int two_adicity = 0;
  HostFr root_of_unity;               // of order 2^two_adicity (Montgomery)
  HostFr mult_gen;                    // fft.Domain.FrMultiplicativeGen (Montgomery)

  // P: one of the generated <curve>_fr_params (params_gen.cuh)
  template <class P>
  static HostFrCtx make() {
    HostFrCtx c;
    c.L = P::N / 2;
    for (int i = 0; i < c.L; i++) {
      c.mod[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
      c.one[i] = (uint64_t)P::r1(2 * i) | ((uint64_t)P::r1(2 * i + 1) << 32);
      c.r2[i] = (uint64_t)P::r2(2 * i) | ((uint64_t)P::r2(2 * i + 1) << 32);
      c.root_of_unity.v[i] = (uint64_t)P::root_of_unity_mont(2 * i) | ((uint64_t)P::root_of_unity_mont(2 * i + 1) << 32);
      c.mult_gen.v[i] = (uint64_t)P::mult_gen_mont(2 * i) | ((uint64_t)P::mult_gen_mont(2 * i + 1) << 32);
    }
    for (int i = c.L; i < HOSTFR_MAX_LIMBS; i++) c.root_of_unity.v[i] = c.mult_gen.v[i] = 0;
    uint64_t x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - c.mod[0] * x;
    c.ninv = 0 - x;
    c.two_adicity = P::TWO_ADICITY;
    return c;
  }

  size_t nbytes() const { return (L > 0 && L <= HOSTFR_MAX_LIMBS) ? (size_t)L * 8 : 0; }
  HostFr zero() const { HostFr r; memset(r.v, 0, sizeof(r.v)); return r; }
  HostFr one_() const { HostFr r = zero(); memcpy(r.v, one, nbytes()); return r; }
  HostFr load(const void* p) const { HostFr r = zero(); memcpy(r.v, p, nbytes()); return r; }
  void store(void* p, const HostFr& a) const { memcpy(p, a.v, nbytes()); }
  bool is_zero(const HostFr& a) const { uint64_t t = 0; for (int i = 0; i < L; i++) t |= a.v[i]; return t == 0; }
  bool eq(const HostFr& a, const HostFr& b) const { return memcmp(a.v, b.v, nbytes()) == 0; }

  bool geq_mod(const uint64_t* a) const {
    for (int i = L - 1; i >= 0; i--) { if (a[i] > mod[i]) return true; if (a[i] < mod[i]) return false; }
    return true;
  }
  void sub_mod(uint64_t* a) const {
    unsigned __int128 br = 0;
    for (int i = 0; i < L; i++) { unsigned __int128 d = (unsigned __int128)a[i] - mod[i] - br; a[i] = (uint64_t)d; br = (d >> 64) & 1; }
  }
  HostFr add(const HostFr& a, const HostFr& b) const {
    HostFr r = zero(); unsigned __int128 c = 0;
    for (int i = 0; i < L; i++) { c += (unsigned __int128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    if (c || geq_mod(r.v)) sub_mod(r.v);
    return r;
  }
  HostFr sub(const HostFr& a, const HostFr& b) const {
    HostFr r = zero(); unsigned __int128 br = 0;
    for (int i = 0; i < L; i++) { unsigned __int128 d = (unsigned __int128)a.v[i] - b.v[i] - br; r.v[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) { unsigned __int128 c = 0; for (int i = 0; i < L; i++) { c += (unsigned __int128)r.v[i] + mod[i]; r.v[i] = (uint64_t)c; c >>= 64; } }
    return r;
  }
  HostFr neg(const HostFr& a) const { return is_zero(a) ? a : sub(zero(), a); }
  // Montgomery product (CIOS)
  HostFr mul(const HostFr& a, const HostFr& b) const {
    uint64_t t[HOSTFR_MAX_LIMBS + 2];
    memset(t, 0, sizeof(t));
    for (int i = 0; i < L; i++) {
      unsigned __int128 c = 0;
      for (int j = 0; j < L; j++) { c += (unsigned __int128)a.v[j] * b.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
      c += t[L]; t[L] = (uint64_t)c; t[L + 1] = (uint64_t)(c >> 64);
      const uint64_t m = t[0] * ninv;
      c = ((unsigned __int128)m * mod[0] + t[0]) >> 64;
      for (int j = 1; j < L; j++) { c += (unsigned __int128)m * mod[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
      c += t[L]; t[L - 1] = (uint64_t)c; t[L] = t[L + 1] + (uint64_t)(c >> 64);
    }
    HostFr r = zero();
    memcpy(r.v, t, nbytes());
    if (t[L] || geq_mod(r.v)) sub_mod(r.v);
    return r;
  }
  HostFr sqr(const HostFr& a) const { return mul(a, a); }
  HostFr from_u64(uint64_t x) const { HostFr r = zero(); r.v[0] = x; HostFr R2 = zero(); memcpy(R2.v, r2, nbytes()); return mul(r, R2); }
  HostFr pow_u64(HostFr base, uint64_t e) const {
    HostFr r = one_();
    while (e) { if (e & 1) r = mul(r, base); base = sqr(base); e >>= 1; }
    return r;
  }
  HostFr pow2k(HostFr a, int k) const { for (int i = 0; i < k; i++) a = sqr(a); return a; }   // a^(2^k)
  HostFr inv(const HostFr& a) const {   // a^(mod-2)
    uint64_t e[HOSTFR_MAX_LIMBS];
    memcpy(e, mod, sizeof(e));
    unsigned __int128 br = 2;
    for (int i = 0; i < L && br; i++) { unsigned __int128 d = (unsigned __int128)e[i] - br; e[i] = (uint64_t)d; br = (d >> 64) & 1; }
    HostFr r = one_(), b = a;
    for (int w = 0; w < L; w++)
      for (int k = 0; k < 64; k++) { if ((e[w] >> k) & 1) r = mul(r, b); b = sqr(b); }
    return r;
  }
  // generator of the subgroup of order 2^logn (fft.NewDomain's Generator)
  HostFr domain_generator(int logn) const { return pow2k(root_of_unity, two_adicity - logn); }
};

}  // namespace gb200
