// Bucket accumulation on the FP64 pipe: XYZZ mixed addition (madd-2008-s, as curve.cuh)
// over the lazily reduced 52-bit-limb field of field52.cuh.
//
// Bounds (p < 2^BITS, spare = 52 L - BITS >= 6; a product accepts inputs < 8p, returns < 2p):
//   accumulator invariant:  X < ~4p, Y < 4p, ZZ < 2p, ZZZ < 2p   (inf: ZZ limbs all zero)
//   U2 = x2*ZZ, S2 = y2*ZZZ              < 2p      (table coordinates are canonical, < p)
//   P  = U2 - X + 4p                      in (0, 6p)     R = S2 - Y + 4p in (0, 6p)
//   PP = P^2, PPP = P*PP, Q = X*PP        < 2p
//   X3 = R^2 - PPP - 2Q + 6p              in (0, 8p)  -> partial reduction to < ~4p
//   T  = Q - X3 + 4p'                     in (0, 6p)
//   Y3 = R*T - Y*PPP + 2p                 in (0, 4p)
//   ZZ3 = ZZ*PP, ZZZ3 = ZZZ*PPP           < 2p
// P == 0 (mod p) is detected exactly on PP (< 2p): PP in {0, p}.
#pragma once
#include "curve.cuh"
#include "field52.cuh"

namespace gb200 {

template <class P52>
struct alignas(16) Affine52 {   // table entry: canonical Montgomery-R52 coordinates as exact doubles
  double x[P52::L], y[P52::L];
};

template <class P52>
struct XYZZ52 {
  using E = F52<P52>;
  E x, y, zz, zzz;
  HD static XYZZ52 inf() { XYZZ52 r; r.x = E::one(); r.y = E::one(); r.zz = E::zero(); r.zzz = E::zero(); return r; }
  HD bool is_inf() const { return zz.limbs_all_zero(); }

  // this = 2 * (ax, ay)   (affine, not infinity; coordinates canonical)        mdbl-2008-s
  HDNI void set_double_affine(const E& ax, const E& ay) {
    if (ay.limbs_all_zero()) { *this = inf(); return; }
    E U = add52<P52>(ay, ay);                       // < 2p
    E V = mul52<P52>(U, U);                         // < 2p
    E W = mul52<P52>(U, V);
    E S = mul52<P52>(ax, V);
    E xx = mul52<P52>(ax, ax);
    E M = add52<P52>(add52<P52>(xx, xx), xx);       // < 6p
    E MM = mul52<P52>(M, M);
    E X3 = sub52<P52, 4>(MM, add52<P52>(S, S));     // MM - 2S + 4p  in (0, 6p)
    partial_reduce52<P52, 4>(X3);                   // < ~4p
    E T = sub52<P52, 5>(S, X3);                     // in (0, 7p)
    E Y3 = sub52<P52, 2>(mul52<P52>(M, T), mul52<P52>(W, ay));   // in (0, 4p)
    x = X3; y = Y3; zz = V; zzz = W;
  }

  // this += (+-)(ax, ay); (ax, ay) = canonical table coordinates as doubles.  The sign is folded
  // into the subtraction that forms R, so negating costs nothing.
  HD void add_mixed(const D52<P52>& dax, const D52<P52>& day, bool negate) {
    constexpr int L = P52::L;
    // infinity of the affine operand: all limbs of x and y zero (canonical storage)
    double t0 = 0.0;
#pragma unroll
    for (int i = 0; i < L; i++) t0 += dax.d[i] + day.d[i];
    if (t0 == 0.0) return;
    if (is_inf()) {
      E ay;
#pragma unroll
      for (int i = 0; i < L; i++) { x.l[i] = f52::to_int(dax.d[i]); ay.l[i] = f52::to_int(day.d[i]); }
      if (negate) { E z = E::zero(); ay = sub52<P52, 1>(z, ay); }   // p - y  (y != 0: no 2-torsion)
      y = ay;
      zz = E::one(); zzz = E::one();
      return;
    }
    const D52<P52> dzz(zz), dzzz(zzz);
    E U2 = mul52<P52>(dax, dzz);
    E S2 = mul52<P52>(day, dzzz);
    E P = sub52<P52, 5>(U2, x);                     // (0, 7p)
    E R;
    if (negate) {                                   // -S2 - Y + 6p  in (0, 6p)
#pragma unroll
      for (int i = 0; i < L; i++) R.l[i] = 6 * (int64_t)P52::mod52(i) - S2.l[i] - y.l[i];
      normalize52<P52>(R);
    } else {
      R = sub52<P52, 4>(S2, y);                     // (0, 6p)
    }
    const D52<P52> dP(P);
    E PP = mul52<P52>(dP, dP);
    if (is_zero_mod_p_lt2p<P52>(PP)) {
      // same x: doubling when R == 0 (mod p), otherwise P + (-P) = infinity
      E r = R;
      canonical52<P52>(r);
      if (r.limbs_all_zero()) {
        E ax, ay;
#pragma unroll
        for (int i = 0; i < L; i++) { ax.l[i] = f52::to_int(dax.d[i]); ay.l[i] = f52::to_int(day.d[i]); }
        if (negate) { E z = E::zero(); ay = sub52<P52, 1>(z, ay); }
        set_double_affine(ax, ay);
      } else {
        *this = inf();
      }
      return;
    }
    const D52<P52> dPP(PP);
    E PPP = mul52<P52>(dP, dPP);
    E Q = mul52<P52>(D52<P52>(x), dPP);
    const D52<P52> dR(R), dPPP(PPP);
    E RR = mul52<P52>(dR, dR);
    // X3 = RR - PPP - 2Q + 6p  in (0, 8p)
    E X3;
#pragma unroll
    for (int i = 0; i < L; i++) X3.l[i] = RR.l[i] - PPP.l[i] - 2 * Q.l[i] + 6 * (int64_t)P52::mod52(i);
    normalize52<P52>(X3);
    partial_reduce52<P52, 4>(X3);                   // < 4p (1 + 1e-13)
    E T = sub52<P52, 5>(Q, X3);                     // (0, 7p)
    E A = mul52<P52>(dR, D52<P52>(T));
    E B = mul52<P52>(D52<P52>(y), dPPP);
    y = sub52<P52, 2>(A, B);                        // (0, 4p)
    x = X3;
    zz = mul52<P52>(dzz, dPP);
    zzz = mul52<P52>(dzzz, dPPP);
  }

  // to the 32-bit representation shared with the reduction kernels (canonical Montgomery R32)
  template <class F>
  HD XYZZ<F> to_xyzz32() const {
    XYZZ<F> r;
    if (is_inf()) return XYZZ<F>::inf();
    to_mont32<P52>(x, r.x.l);
    to_mont32<P52>(y, r.y.l);
    to_mont32<P52>(zz, r.zz.l);
    to_mont32<P52>(zzz, r.zzz.l);
    return r;
  }
};

// gnark affine point (Montgomery R32) -> table entry (Montgomery R52, canonical, doubles)
template <class P52, class F>
HD Affine52<P52> affine_to_52(const Affine<F>& a) {
  Affine52<P52> r;
  F52<P52> x = from_mont32<P52>(a.x.l), y = from_mont32<P52>(a.y.l);
  canonical52<P52>(x);
  canonical52<P52>(y);
#pragma unroll
  for (int i = 0; i < P52::L; i++) { r.x[i] = (double)x.l[i]; r.y[i] = (double)y.l[i]; }
  return r;
}

}  // namespace gb200
