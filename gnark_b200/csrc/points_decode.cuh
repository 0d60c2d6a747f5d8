// Point slices in gnark-crypto's SERIALISED encodings -> device tables (second half: the PLONK proving
// key).  backend/plonk/bn254/marshal.go:96-129 writes pk.Kzg and pk.KzgLagrange with gnark-crypto's encoder: a uint32
// length and then the points, compressed (WriteTo, the default) or uncompressed (WriteRawTo); backend/groth16 keys use the
// same encoder outside their dump format (marshal.go:136-214).  Decoding is the expensive part of loading such a key
// on the CPU - one square root in Fp per compressed point - and is embarrassingly parallel: one thread per point here.
//
// Encoding (gnark-crypto <curve>/marshal.go; the BLS12-381 case is the ZCash convention): coordinates big-endian,
// canonical (NOT Montgomery); metadata in the most significant bits of the first byte:
//     BN254 (2 bits):  00 uncompressed, 10 compressed / y smallest, 11 compressed / y largest, 01 compressed infinity
//     others (3 bits): 000 uncompressed, 010 uncompressed infinity, 100 / 101 compressed smallest / largest,
//                      110 compressed infinity
// uncompressed: X || Y; compressed: X only, y = sqrt(x^3 + b) with the root picked by the "lexicographically largest"
// flag (y > (p-1)/2 on canonical values; for an Fp2 element the A1 coordinate decides unless it is zero, then A0).
// Fp2 coordinates (G2) are written A1 || A0.  Infinity = (0, 0) in memory.
// Square roots: exponentiation by (p+1)/4 where p = 3 mod 4 (BN254, BLS12-381, BW6-761), Tonelli-Shanks otherwise
// (BLS12-377: p - 1 = 2^46 t); in Fp2 = Fp[u]/(u^2 + beta) by the norm: x0^2 = (a0 +- sqrt(a0^2 + beta a1^2)) / 2,
// x1 = a1 / (2 x0).  Twist coefficients: BN254 3/(9+u), BLS12-381 4(1+u), BLS12-377 1/u, BW6-761 (over Fp) 4.
// Pinned by reference-held data: the compressed generators in gnark's serialised verifying keys (BN254, BLS12-381) and
// the 8192 compressed BLS12-381 G1 points and the 65 compressed G2 points of the Ethereum KZG ceremony file, the compressed
// G2 generators of gnark's verifying keys (tests/test_golden_kzg.py, test_emulation.py, test_gpu_round2.py).
#pragma once
#include "curve.cuh"

namespace gb200 {

enum { POINTS_RAW = 1, POINTS_COMPRESSED = 2 };
enum { DECODE_OK = 0, DECODE_BAD_FLAGS = 1, DECODE_NOT_REDUCED = 2, DECODE_NOT_ON_CURVE = 3 };

template <class FB>   // FB: the base prime field Fp<P>
struct DecodeConsts {
  FB b;                       // curve coefficient of this group's curve over Fp (G1; BW6-761 G2) - Montgomery
  FB b2[2];                   // twist coefficient b' = b2[0] + b2[1] u (G2 over Fp2)
  FB inv2;                    // 1 / 2
  uint32_t sqrt_exp[FB::N];   // (p + 1) / 4, valid when p = 3 mod 4
  uint32_t half[FB::N];       // (p - 1) / 2
  int flag_bits;              // 2 (BN254) or 3
  int sqrt_ok;                // p = 3 mod 4
  // Tonelli-Shanks, p = 1 mod 4: p - 1 = 2^ts_s * t, ts_e = (t - 1) / 2, ts_c = g^t for a quadratic non-residue g
  int ts_s;
  uint32_t ts_e[FB::N];
  FB ts_c;
};

template <class FB>
HD FB decode_pow(const FB& base, const uint32_t* e);

// per curve id (include/gnark_b200.h: 0 BN254, 1 BLS12-381, 2 BLS12-377, 3 BW6-761): G1 coefficient, and the twist as
// b * xi^(+-1) with xi = xi0 + xi1 u (G2 over Fp2) or a small integer (BW6-761 G2, over Fp)
struct CurveBSpec { int b1; int b2; int xi0, xi1, invert; };
inline CurveBSpec decode_curve_spec(int curve) {
  switch (curve) {
    case 0: return {3, 3, 9, 1, 1};      // y^2 = x^3 + 3 ; twist 3 / (9 + u)
    case 1: return {4, 4, 1, 1, 0};      // y^2 = x^3 + 4 ; twist 4 (1 + u)
    case 2: return {1, 1, 0, 1, 1};      // y^2 = x^3 + 1 ; twist 1 / u
    default: return {-1, 4, 0, 0, 0};    // BW6-761: y^2 = x^3 - 1 ; G2 (over Fp) y^2 = x^3 + 4
  }
}

template <class FB>
inline FB decode_small(int v) {
  FB r = FB::one().mul_small(v < 0 ? (unsigned)(-v) : (unsigned)v);
  return v < 0 ? r.neg() : r;
}

// host side: the constants of one base field for one curve and group.  F: the coordinate field of the group
template <class F, class FB>
inline DecodeConsts<FB> decode_make_consts(int curve, int group) {
  constexpr int N = FB::N;
  const CurveBSpec spec = decode_curve_spec(curve);
  DecodeConsts<FB> k;
  k.b = decode_small<FB>(group == 2 ? spec.b2 : spec.b1);
  k.b2[0] = FB::zero(); k.b2[1] = FB::zero();
  if constexpr (F::DEGREE == 2) {
    F xi; xi.a0 = decode_small<FB>(spec.xi0); xi.a1 = decode_small<FB>(spec.xi1);
    F bb; bb.a0 = decode_small<FB>(spec.b2); bb.a1 = FB::zero();
    const F t = spec.invert ? bb * xi.inverse() : bb * xi;
    k.b2[0] = t.a0; k.b2[1] = t.a1;
  }
  k.inv2 = (FB::one() + FB::one()).inverse();
  uint32_t pl[N], p1[N];
  for (int i = 0; i < N; i++) pl[i] = FB::Params::mod(i);
  k.sqrt_ok = (pl[0] & 3u) == 3u;
  uint64_t carry = 1;     // (p + 1) / 4 and (p - 1) / 2 = p >> 1
  for (int i = 0; i < N; i++) { const uint64_t v = (uint64_t)pl[i] + carry; p1[i] = (uint32_t)v; carry = v >> 32; }
  for (int i = 0; i < N; i++) {
    k.sqrt_exp[i] = (p1[i] >> 2) | (i + 1 < N ? p1[i + 1] << 30 : (uint32_t)(carry << 30));
    k.half[i] = (pl[i] >> 1) | (i + 1 < N ? pl[i + 1] << 31 : 0u);
  }
  k.flag_bits = FB::Params::BITS == 254 ? 2 : 3;      // BN254: two spare bits in the first byte
  // Tonelli-Shanks constants (only used when p = 1 mod 4)
  k.ts_s = 0;
  for (int i = 0; i < N; i++) k.ts_e[i] = 0;
  k.ts_c = FB::one();
  if (!k.sqrt_ok) {
    uint32_t t[N];
    for (int i = 0; i < N; i++) t[i] = pl[i];
    t[0] &= ~1u;                                       // p - 1
    auto shr1 = [&](uint32_t* v) { for (int i = 0; i < N; i++) v[i] = (v[i] >> 1) | (i + 1 < N ? v[i + 1] << 31 : 0u); };
    while (!(t[0] & 1u)) { shr1(t); k.ts_s++; }        // t odd, p - 1 = 2^s t
    for (int i = 0; i < N; i++) k.ts_e[i] = t[i];
    shr1(k.ts_e);                                      // (t - 1) / 2
    for (unsigned g = 2;; g++) {                       // smallest non-residue: g^((p-1)/2) = -1
      const FB gg = FB::one().mul_small(g);
      if (decode_pow(gg, k.half) == FB::one().neg()) { k.ts_c = decode_pow(gg, t); break; }
    }
  }
  return k;
}

// big-endian canonical bytes -> limbs; `mask_bits` most significant bits of the first byte are cleared
template <class FB>
HD FB decode_read_fp(const uint8_t* p, int mask_bits) {
  constexpr int N = FB::N;
  FB v;
#pragma unroll
  for (int j = 0; j < N; j++) {
    const uint8_t* q = p + 4 * (N - 1 - j);
    uint32_t w = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    if (j == N - 1 && mask_bits) w &= 0xffffffffu >> mask_bits;
    v.l[j] = w;
  }
  return v;
}
template <class FB>
HD bool decode_lt_mod(const FB& v) {     // canonical value < p
  for (int j = FB::N - 1; j >= 0; j--) {
    const uint32_t m = FB::Params::mod(j);
    if (v.l[j] < m) return true;
    if (v.l[j] > m) return false;
  }
  return false;
}
template <class FB>
HD bool decode_gt(const FB& canonical, const uint32_t* half) {   // canonical > half
  for (int j = FB::N - 1; j >= 0; j--) {
    if (canonical.l[j] > half[j]) return true;
    if (canonical.l[j] < half[j]) return false;
  }
  return false;
}
template <class FB>
HD FB decode_pow(const FB& base, const uint32_t* e) {
  FB r = FB::one();
  bool started = false;
  for (int w = FB::N - 1; w >= 0; w--)
    for (int bit = 31; bit >= 0; bit--) {
      if (started) r = r.sqr();
      if ((e[w] >> bit) & 1) { r = started ? r * base : base; started = true; }
    }
  return r;
}

// square root in Fp; false when a is not a residue
template <class FB>
HD bool decode_sqrt_fp(const FB& a, const DecodeConsts<FB>& k, FB& r) {
  if (a.is_zero()) { r = FB::zero(); return true; }
  if (k.sqrt_ok) {
    r = decode_pow(a, k.sqrt_exp);
    return r.sqr() == a;
  }
  // Tonelli-Shanks: w = a^((t-1)/2), x = a w, b = a^t; invariant x^2 = a b, b of order dividing 2^(v-1)
  const FB one = FB::one();
  const FB w = decode_pow(a, k.ts_e);
  FB x = a * w;
  FB b = x * w;
  FB c = k.ts_c;
  int v = k.ts_s;
  while (!(b == one)) {
    int m = 0;
    FB t2 = b;
    while (!(t2 == one)) { t2 = t2.sqr(); m++; if (m >= v) return false; }
    FB cc = c;
    for (int i = 0; i < v - m - 1; i++) cc = cc.sqr();
    x = x * cc;
    c = cc.sqr();
    b = b * c;
    v = m;
  }
  r = x;
  return true;
}

// square root in Fp2 = Fp[u] / (u^2 + BETA); false when a is not a square
template <class F, class FB>
HD bool decode_sqrt_fp2(const F& a, const DecodeConsts<FB>& k, F& r) {
  constexpr unsigned BETA = F::BETA_VALUE;
  if (a.a1.is_zero()) {
    FB s;
    if (decode_sqrt_fp(a.a0, k, s)) { r.a0 = s; r.a1 = FB::zero(); return true; }
    // a0 = (y1 u)^2 = -BETA y1^2
    const FB q = a.a0.neg() * FB::one().mul_small(BETA).inverse();
    if (decode_sqrt_fp(q, k, s)) { r.a0 = FB::zero(); r.a1 = s; return true; }
    return false;
  }
  const FB norm = a.a0.sqr() + a.a1.sqr().mul_small(BETA);
  FB al;
  if (!decode_sqrt_fp(norm, k, al)) return false;
  for (int sign = 0; sign < 2; sign++) {
    const FB d = (sign ? a.a0 - al : a.a0 + al) * k.inv2;
    FB x0;
    if (!decode_sqrt_fp(d, k, x0) || x0.is_zero()) continue;
    r.a0 = x0;
    r.a1 = a.a1 * x0.dbl().inverse();
    if (r * r == a) return true;
  }
  return false;
}

// one point; returns a DECODE_* status.  F: coordinate field (FB or Fp2 over FB)
template <class F, class FB>
HD int decode_point(const uint8_t* p, int compressed, const DecodeConsts<FB>& k, Affine<F>& out) {
  constexpr int DEG = F::DEGREE;
  constexpr size_t FPB = (size_t)FB::N * 4;
  const int fb = k.flag_bits;
  const uint32_t flags = p[0] >> (8 - fb);
  bool inf, largest = false;
  if (fb == 2) {          // BN254: 00 raw, 01 compressed infinity, 10 / 11 compressed smallest / largest
    if (compressed) { if (flags == 0) return DECODE_BAD_FLAGS; inf = flags == 1; largest = flags == 3; }
    else { if (flags != 0) return DECODE_BAD_FLAGS; inf = false; }
  } else {                // 000 raw, 010 raw infinity, 100 / 101 compressed smallest / largest, 110 compressed infinity
    if (compressed) { if (flags != 4 && flags != 5 && flags != 6) return DECODE_BAD_FLAGS; inf = flags == 6; largest = flags == 5; }
    else { if (flags != 0 && flags != 2) return DECODE_BAD_FLAGS; inf = flags == 2; }
  }
  if (inf) { out = Affine<F>::inf(); return DECODE_OK; }
  // coordinates (Fp2: A1 || A0)
  FB c[2][DEG];
  const int ncoord = compressed ? 1 : 2;
  for (int q = 0; q < ncoord; q++)
    for (int d = 0; d < DEG; d++) {
      const size_t off = ((size_t)q * DEG + d) * FPB;
      FB v = decode_read_fp<FB>(p + off, off == 0 ? fb : 0);
      if (!decode_lt_mod(v)) return DECODE_NOT_REDUCED;
      c[q][DEG - 1 - d] = v.to_mont();
    }
  if constexpr (DEG == 1) {
    out.x = c[0][0];
    if (!compressed) {
      out.y = c[1][0];
      if (out.x.is_zero() && out.y.is_zero()) return DECODE_OK;            // BN254 writes infinity as all zeros
      return (out.y.sqr() == out.x.sqr() * out.x + k.b) ? DECODE_OK : DECODE_NOT_ON_CURVE;
    }
    const FB y2 = out.x.sqr() * out.x + k.b;
    FB y;
    if (!decode_sqrt_fp(y2, k, y)) return DECODE_NOT_ON_CURVE;
    if (decode_gt(y.from_mont(), k.half) != largest) y = y.neg();
    out.y = y;
    return DECODE_OK;
  } else {
    out.x.a0 = c[0][0]; out.x.a1 = c[0][1];
    if (!compressed) {
      out.y.a0 = c[1][0]; out.y.a1 = c[1][1];
      return DECODE_OK;    // raw G2: membership of the twist is the caller's subgroup check, as with UnsafeReadFrom
    }
    F bt; bt.a0 = k.b2[0]; bt.a1 = k.b2[1];
    const F y2 = out.x * out.x * out.x + bt;
    F y;
    if (!decode_sqrt_fp2<F, FB>(y2, k, y)) return DECODE_NOT_ON_CURVE;
    const bool big = !y.a1.is_zero() ? decode_gt(y.a1.from_mont(), k.half) : decode_gt(y.a0.from_mont(), k.half);
    if (big != largest) y = y.neg();
    out.y = y;
    return DECODE_OK;
  }
}

}  // namespace gb200
