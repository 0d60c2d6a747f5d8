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
