// Short-Weierstrass a = 0 group law, generic over the coordinate field F
// (Fp for G1 and BW6-761 G2, Fp2 for the other G2s).  The curve coefficient b is
// never needed by MSM (a = 0 addition/doubling formulas do not use it).
//
// Bucket accumulators are XYZZ ("extended Jacobian": x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2) - the representation gnark-crypto's MultiExp uses for its buckets
// (SURVEY.md §8d op model: 8M+2S mixed add).  Formulas: EFD madd-2008-s,
// add-2008-s, dbl-2008-s-1, mdbl-2008-s (public formulas, written from the
// algebra).
//
// Memory layouts (gnark):
//   Affine<F>   = {X, Y}        infinity = (0, 0)
//   Jacobian<F> = {X, Y, Z}     x = X/Z^2, y = Y/Z^3, infinity Z = 0
#pragma once
#include "field.cuh"

namespace gb200 {

template <class F>
struct alignas(16) Affine {
  F x, y;
  HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  HD static Affine inf() { Affine p; p.x = F::zero(); p.y = F::zero(); return p; }
  HD Affine neg() const { Affine p; p.x = x; p.y = y.neg(); return p; }
};

template <class F>
struct alignas(16) Jacobian {
  F x, y, z;
};

template <class F>
struct alignas(16) XYZZ {
  F x, y, zz, zzz;

  HD static XYZZ inf() { XYZZ p; p.x = F::one(); p.y = F::one(); p.zz = F::zero(); p.zzz = F::zero(); return p; }
  HD bool is_inf() const { return zz.is_zero(); }
  HD static XYZZ from_affine(const Affine<F>& a) {
    if (a.is_inf()) return inf();
    XYZZ p; p.x = a.x; p.y = a.y; p.zz = F::one(); p.zzz = F::one(); return p;
  }
  HD XYZZ neg() const { XYZZ p = *this; p.y = y.neg(); return p; }

  // this = 2 * a (a affine, not infinity)             mdbl-2008-s
  HDNI void set_double_affine(const Affine<F>& a) {
    if (a.y.is_zero()) { *this = inf(); return; }
    F U = a.y.dbl();
    F V = U.sqr();
    F W = U * V;
    F S = a.x * V;
    F xx = a.x.sqr();
    F M = xx.dbl() + xx;
    x = M.sqr() - S.dbl();
    y = M * (S - x) - W * a.y;
    zz = V;
    zzz = W;
  }

  // this *= 2                                           dbl-2008-s-1
  HDNI void dbl() {
    if (is_inf()) return;
    if (y.is_zero()) { *this = inf(); return; }
    F U = y.dbl();
    F V = U.sqr();
    F W = U * V;
    F S = x * V;
    F xx = x.sqr();
    F M = xx.dbl() + xx;
    F X3 = M.sqr() - S.dbl();
    F Y3 = M * (S - X3) - W * y;
    x = X3;
    y = Y3;
    zz = V * zz;
    zzz = W * zzz;
  }

  // this += a (affine; handles a = inf, this = inf, a = +-this)   madd-2008-s
  HD void add_mixed(const Affine<F>& a) {
    if (a.is_inf()) return;
    if (is_inf()) { x = a.x; y = a.y; zz = F::one(); zzz = F::one(); return; }
    F U2 = a.x * zz;
    F S2 = a.y * zzz;
    F P = U2 - x;
    F R = S2 - y;
    if (P.is_zero()) {
      if (R.is_zero()) set_double_affine(a);
      else *this = inf();
      return;
    }
    F PP = P.sqr();
    F PPP = P * PP;
    F Q = x * PP;
    F X3 = R.sqr() - PPP - Q.dbl();
    y = R * (Q - X3) - y * PPP;
    x = X3;
    zz = zz * PP;
    zzz = zzz * PPP;
  }

  // this += q                                           add-2008-s
  HDNI void add(const XYZZ& q) {
    if (q.is_inf()) return;
    if (is_inf()) { *this = q; return; }
    F U1 = x * q.zz;
    F U2 = q.x * zz;
    F S1 = y * q.zzz;
    F S2 = q.y * zzz;
    F P = U2 - U1;
    F R = S2 - S1;
    if (P.is_zero()) {
      if (R.is_zero()) dbl();
      else *this = inf();
      return;
    }
    F PP = P.sqr();
    F PPP = P * PP;
    F Q = U1 * PP;
    F X3 = R.sqr() - PPP - Q.dbl();
    y = R * (Q - X3) - S1 * PPP;
    x = X3;
    zz = zz * q.zz * PP;
    zzz = zzz * q.zzz * PPP;
  }

  // inversion-free conversion to gnark's Jacobian: Z = ZZZ (= z^3) gives
  // X' = X*ZZ^2 (x = X'/Z^2), Y' = Y*ZZZ^2 (y = Y'/Z^3).
  HD Jacobian<F> to_jacobian() const {
    Jacobian<F> j;
    if (is_inf()) { j.x = F::one(); j.y = F::one(); j.z = F::zero(); return j; }
    j.x = x * zz.sqr();
    j.y = y * zzz.sqr();
    j.z = zzz;
    return j;
  }
  HD static XYZZ from_jacobian(const Jacobian<F>& j) {
    if (j.z.is_zero()) return inf();
    XYZZ p; p.x = j.x; p.y = j.y; p.zz = j.z.sqr(); p.zzz = p.zz * j.z; return p;
  }
  HD Affine<F> to_affine() const {
    if (is_inf()) return Affine<F>::inf();
    Affine<F> a;
    F zi = zzz.inverse();           // 1/z^3
    F zi2 = (zi * zz).sqr();        // (z^2/z^3)^2 = 1/z^2
    a.x = x * zi2;
    a.y = y * zi;
    return a;
  }
};

// ---------------------------------------------------------------------------
// Lane-cooperative point arithmetic for the LATENCY-bound reduction tail of the MSM.
//
// One thread's XYZZ addition is 14 dependent-ish field multiplications of ~136 carry-chained IMAD.WIDE each: 4-8 us on a
// B200 when only a few warps run, and the tail (combine, chunk sums, set sum, Horner) is ~60 of them in a row.  The
// formulas have more parallelism than that: an addition is 4 levels of at most 4 INDEPENDENT products, a doubling 4 levels
// of at most 4.  A QUAD of 4 adjacent lanes therefore holds the SAME operands (replicated), each lane computes ONE product
// of the level, and the four results are exchanged with warp shuffles - the dependent multiplication depth drops from
// 14 to 4 (9 to 4 for a doubling).  ("warp-shuffle point additions" of the north star, applied where latency, not
// throughput, is the limit; the throughput-bound accumulate kernel keeps one thread per addition.)
//
// The arithmetic is written once against a Quad policy: QuadDev exchanges through __shfl_sync, QuadHost computes all four
// products itself - which is how the formulas and their special cases are unit-tested without a GPU (tests/test_emulation.py).
// ---------------------------------------------------------------------------
struct QuadHost {
  template <class F>
  HD void mul4(const F& a0, const F& b0, const F& a1, const F& b1, const F& a2, const F& b2, const F& a3, const F& b3,
               F& o0, F& o1, F& o2, F& o3) const {
    o0 = a0 * b0; o1 = a1 * b1; o2 = a2 * b2; o3 = a3 * b3;
  }
};
#if defined(__CUDACC__)
struct QuadDev {
  unsigned mask;   // the four lanes of this quad
  int sub;         // this lane's index in the quad
  __device__ static QuadDev make() {
    const unsigned lane = threadIdx.x & 31u;
    return QuadDev{0xFu << (lane & 28u), (int)(lane & 3u)};
  }
  // lane `sub` multiplies pair number `sub`; every lane of the quad receives all four products
  template <class F>
  __device__ void mul4(const F& a0, const F& b0, const F& a1, const F& b1, const F& a2, const F& b2, const F& a3, const F& b3,
                       F& o0, F& o1, F& o2, F& o3) const {
    constexpr int NW = sizeof(F) / 4;
    F a, b;
    {
      const uint32_t *p0 = reinterpret_cast<const uint32_t*>(&a0), *p1 = reinterpret_cast<const uint32_t*>(&a1),
                     *p2 = reinterpret_cast<const uint32_t*>(&a2), *p3 = reinterpret_cast<const uint32_t*>(&a3);
      const uint32_t *q0 = reinterpret_cast<const uint32_t*>(&b0), *q1 = reinterpret_cast<const uint32_t*>(&b1),
                     *q2 = reinterpret_cast<const uint32_t*>(&b2), *q3 = reinterpret_cast<const uint32_t*>(&b3);
      uint32_t* pa = reinterpret_cast<uint32_t*>(&a);
      uint32_t* pb = reinterpret_cast<uint32_t*>(&b);
#pragma unroll
      for (int w = 0; w < NW; w++) {
        pa[w] = sub == 0 ? p0[w] : (sub == 1 ? p1[w] : (sub == 2 ? p2[w] : p3[w]));
        pb[w] = sub == 0 ? q0[w] : (sub == 1 ? q1[w] : (sub == 2 ? q2[w] : q3[w]));
      }
    }
    const F m = a * b;
    const uint32_t* pm = reinterpret_cast<const uint32_t*>(&m);
    uint32_t *r0 = reinterpret_cast<uint32_t*>(&o0), *r1 = reinterpret_cast<uint32_t*>(&o1),
             *r2 = reinterpret_cast<uint32_t*>(&o2), *r3 = reinterpret_cast<uint32_t*>(&o3);
    const int base = (int)(threadIdx.x & 28u);
#pragma unroll
    for (int w = 0; w < NW; w++) {
      r0[w] = __shfl_sync(mask, pm[w], base + 0);
      r1[w] = __shfl_sync(mask, pm[w], base + 1);
      r2[w] = __shfl_sync(mask, pm[w], base + 2);
      r3[w] = __shfl_sync(mask, pm[w], base + 3);
    }
  }
};
#endif

// p *= 2 (all lanes of the quad hold the same p and receive the same result)            dbl-2008-s-1, 4 levels
template <class F, class Q>
HD void xyzz_dbl_coop(XYZZ<F>& p, const Q& quad) {
  if (p.is_inf()) return;
  if (p.y.is_zero()) { p = XYZZ<F>::inf(); return; }
  const F U = p.y.dbl();
  F V, XX, d0, d1;
  quad.mul4(U, U, p.x, p.x, U, U, p.x, p.x, V, XX, d0, d1);                 // level 1: V = U^2, XX = X^2
  const F M = XX.dbl() + XX;
  F W, S, ZZ3, MM;
  quad.mul4(U, V, p.x, V, V, p.zz, M, M, W, S, ZZ3, MM);                    // level 2: W = U V, S = X V, ZZ3 = V ZZ, M^2
  const F X3 = MM - S.dbl();
  F WY, ZZZ3, T;
  quad.mul4(W, p.y, W, p.zzz, M, S - X3, W, p.y, WY, ZZZ3, T, d0);          // level 3: W Y, ZZZ3 = W ZZZ, M (S - X3)
  p.x = X3;
  p.y = T - WY;
  p.zz = ZZ3;
  p.zzz = ZZZ3;
}

// p += q (all lanes of the quad hold the same p, q and receive the same result)          add-2008-s, 4 levels
template <class F, class Q>
HD void xyzz_add_coop(XYZZ<F>& p, const XYZZ<F>& q, const Q& quad) {
  if (q.is_inf()) return;
  if (p.is_inf()) { p = q; return; }
  F U1, U2, S1, S2;
  quad.mul4(p.x, q.zz, q.x, p.zz, p.y, q.zzz, q.y, p.zzz, U1, U2, S1, S2);   // level 1
  const F P = U2 - U1;
  const F R = S2 - S1;
  if (P.is_zero()) {
    if (R.is_zero()) xyzz_dbl_coop<F, Q>(p, quad);
    else p = XYZZ<F>::inf();
    return;
  }
  F PP, RR, ZZa, ZZZa;
  quad.mul4(P, P, R, R, p.zz, q.zz, p.zzz, q.zzz, PP, RR, ZZa, ZZZa);        // level 2
  F PPP, Qv, ZZ3, d0;
  quad.mul4(P, PP, U1, PP, ZZa, PP, P, PP, PPP, Qv, ZZ3, d0);                // level 3
  const F X3 = RR - PPP - Qv.dbl();
  F T1, T2, ZZZ3;
  quad.mul4(R, Qv - X3, S1, PPP, ZZZa, PPP, S1, PPP, T1, T2, ZZZ3, d0);      // level 4
  p.x = X3;
  p.y = T1 - T2;
  p.zz = ZZ3;
  p.zzz = ZZZ3;
}

// k * p by double-and-add, cooperative
template <class F, class Q>
HD XYZZ<F> xyzz_mul_small_coop(const XYZZ<F>& p, uint32_t k, const Q& quad) {
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int bit = 31; bit >= 0; bit--) {
    xyzz_dbl_coop<F, Q>(acc, quad);
    if ((k >> bit) & 1) xyzz_add_coop<F, Q>(acc, p, quad);
  }
  return acc;
}

// k * p by double-and-add (k small: window offsets, Horner steps)
template <class F>
HD XYZZ<F> xyzz_mul_small(XYZZ<F> p, uint32_t k) {
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int bit = 31; bit >= 0; bit--) {
    acc.dbl();
    if ((k >> bit) & 1) acc.add(p);
  }
  return acc;
}

}  // namespace gb200
