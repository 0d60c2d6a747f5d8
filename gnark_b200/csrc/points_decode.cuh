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
// flag (y > (p-1)/2 on canonical values).  Fp2 coordinates (G2) are written A1 || A0.  Infinity = (0, 0) in memory.
// Pinned by reference-held data: the compressed generators in gnark's serialised verifying keys (BN254, BLS12-381) and
// the 8192 compressed BLS12-381 points of the Ethereum KZG ceremony file (tests/test_golden_kzg.py, test_gpu_round2.py).
#pragma once
#include "curve.cuh"

namespace gb200 {

enum { POINTS_RAW = 1, POINTS_COMPRESSED = 2 };
enum { DECODE_OK = 0, DECODE_BAD_FLAGS = 1, DECODE_NOT_REDUCED = 2, DECODE_NOT_ON_CURVE = 3 };

template <class FB>   // FB: the base prime field Fp<P>
struct DecodeConsts {
  FB b;                       // curve coefficient (Montgomery)
  uint32_t sqrt_exp[FB::N];   // (p + 1) / 4, valid when p = 3 mod 4
  uint32_t half[FB::N];       // (p - 1) / 2
  int flag_bits;              // 2 (BN254) or 3
  int sqrt_ok;                // p = 3 mod 4
};

// host side: the constants of one base field; b_small = the curve coefficient as a small signed integer
// (BN254 3, BLS12-381 4, BLS12-377 1, BW6-761 -1)
template <class FB>
inline DecodeConsts<FB> decode_make_consts(int b_small) {
  constexpr int N = FB::N;
  DecodeConsts<FB> k;
  k.b = FB::one().mul_small(b_small < 0 ? (unsigned)(-b_small) : (unsigned)b_small);
  if (b_small < 0) k.b = k.b.neg();
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
    FB y = decode_pow(y2, k.sqrt_exp);
    if (y.sqr() != y2) return DECODE_NOT_ON_CURVE;
    if (decode_gt(y.from_mont(), k.half) != largest) y = y.neg();
    out.y = y;
    return DECODE_OK;
  } else {
    // G2: uncompressed only (the PLONK / Groth16 proving keys hold no compressed G2 slices worth a kernel)
    out.x.a0 = c[0][0]; out.x.a1 = c[0][1];
    out.y.a0 = c[1][0]; out.y.a1 = c[1][1];
    return DECODE_OK;      // membership of the twist is the caller's subgroup check, as with UnsafeReadFrom
  }
}

}  // namespace gb200
