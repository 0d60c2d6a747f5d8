// PLONK quotient building blocks: the fused constraint
// evaluation of computeNumerator (backend/plonk/bn254/prove.go:841-1123, closures
// gateConstraint :871-889, orderingConstraint :907-931, localConstraint :933-941,
// allConstraints :961-988), the 1/(X^n - 1) scaling of divideByZH (:1287-1324) and a
// parallel batch inversion (batchInvert :1130-1143).  No BSB22 commitment gates yet.
#pragma once
#include <cuda_runtime.h>

#include "ntt.cuh"

namespace gb200 {

constexpr int PLONK_MAX_BLIND = 4;  // blinding polynomials have degree <= 2 (order 1 or 2) in the reference

template <class Fr>
struct PlonkCosetArgs {
  // device vectors, n elements each, Lagrange basis ON THE CURRENT COSET, regular (natural) layout
  const Fr *l, *r, *o, *z, *s1, *s2, *s3, *ql, *qr, *qm, *qo, *qk;
  const Fr* tw;          // w^j, j < n/2 (domain0 twiddles)
  const Fr* den_inv;     // 1 / (coset*w^j - 1), j < n   (precomputedDenominators, :1002-1007)
  Fr* out;               // rho*n elements, LagrangeCoset on the big domain, BIT-REVERSED layout
  Fr alpha, beta, gamma;
  Fr coset;              // current coset shift  (g, g*w4, g*w4^2, ...)
  Fr cs, css;            // domain1.FrMultiplicativeGen and its square (:891-893)
  Fr coset_n_minus_one;  // coset^n - 1
  Fr lone_scale;         // (coset^n - 1) / n
  Fr bl[PLONK_MAX_BLIND], br[PLONK_MAX_BLIND], bo[PLONK_MAX_BLIND], bz[PLONK_MAX_BLIND];
  int nbl, nbr, nbo, nbz;  // number of blinding coefficients (original, unscaled)
  uint32_t n, logn, rho, log_rho, coset_index;
};

template <class Fr>
HD Fr plonk_horner(const Fr* c, int nc, const Fr& x) {
  Fr acc = Fr::zero();
  for (int k = nc - 1; k >= 0; k--) acc = acc * x + c[k];
  return acc;
}

// One point of one coset.  wj = w^j, wj1 = w^(j+1 mod n).
template <class Fr>
HD Fr plonk_all_constraints(const PlonkCosetArgs<Fr>& a, uint32_t j, const Fr& wj, const Fr& wj1) {
  const uint32_t j1 = (j + 1 == a.n) ? 0 : j + 1;
  const Fr x = a.coset * wj;     // evaluation point
  const Fr x1 = a.coset * wj1;
  // blinded L, R, O, Z, ZS: p + (X^n - 1) b(X) on the coset (:967-981; the reference pre-scales b's
  // coefficients so that evaluating at w^j yields exactly this value)
  Fr L = a.l[j] + a.coset_n_minus_one * plonk_horner(a.bl, a.nbl, x);
  Fr R = a.r[j] + a.coset_n_minus_one * plonk_horner(a.br, a.nbr, x);
  Fr O = a.o[j] + a.coset_n_minus_one * plonk_horner(a.bo, a.nbo, x);
  Fr Z = a.z[j] + a.coset_n_minus_one * plonk_horner(a.bz, a.nbz, x);
  Fr ZS = a.z[j1] + a.coset_n_minus_one * plonk_horner(a.bz, a.nbz, x1);
  // gate (:871-889)
  Fr gate = a.ql[j] * L + a.qr[j] * R + a.qm[j] * L * R + a.qo[j] * O + a.qk[j];
  // ordering (:907-931)
  Fr id = x * a.beta;
  Fr t0 = a.gamma + L + id;
  Fr t1 = id * a.cs + R + a.gamma;
  Fr t2 = id * a.css + O + a.gamma;
  Fr rr = t0 * t1 * t2 * Z;
  t0 = a.s1[j] * a.beta + L + a.gamma;
  t1 = a.s2[j] * a.beta + R + a.gamma;
  t2 = a.s3[j] * a.beta + O + a.gamma;
  Fr ordering = t0 * t1 * t2 * ZS - rr;
  // local (:933-941): (Z - 1) * L1(x), L1(x) = (x^n - 1) / (n (x - 1))
  Fr local = (Z - Fr::one()) * (a.lone_scale * a.den_inv[j]);
  return (local * a.alpha + ordering) * a.alpha + gate;   // :985
}

#ifdef __CUDACC__
template <class Fr>
__global__ void __launch_bounds__(256) k_plonk_constraints(PlonkCosetArgs<Fr> a) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n) return;
  const uint32_t half = a.n >> 1;
  // w^j for j >= n/2 is -w^(j - n/2)
  auto w_at = [&](uint32_t k) -> Fr {
    if (a.n == 1) return Fr::one();
    return k < half ? a.tw[k] : a.tw[k - half].neg();
  };
  const uint32_t j1 = (j + 1 == a.n) ? 0 : j + 1;
  const Fr v = plonk_all_constraints<Fr>(a, j, w_at(j), w_at(j1));
  // cres[bitrev_{rho n}(rho*j + i)] (:1070-1076)
  const uint32_t m = a.rho * j + a.coset_index;
  a.out[ntt_bitrev(m, (int)(a.logn + a.log_rho))] = v;
}

// r[i] *= tab[bitrev(i) % rho]   (divideByZH :1312-1317)
template <class Fr>
__global__ void __launch_bounds__(256) k_plonk_zh_scale(Fr* __restrict__ r, uint32_t logm, uint32_t rho,
                                                        const Fr* __restrict__ tab) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << logm)) return;
  r[i] = r[i] * tab[ntt_bitrev(i, (int)logm) % rho];
}

// in-place batch inversion, one thread per chunk of BI_CHUNK elements (Montgomery's trick);
// zeros are left as zeros.
constexpr int BI_CHUNK = 32;
template <class Fr>
__global__ void __launch_bounds__(128) k_batch_invert(Fr* __restrict__ v, size_t n) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t lo = t * BI_CHUNK;
  if (lo >= n) return;
  const size_t hi = lo + BI_CHUNK < n ? lo + BI_CHUNK : n;
  Fr pre[BI_CHUNK];
  Fr acc = Fr::one();
  for (size_t i = lo; i < hi; i++) {
    pre[i - lo] = acc;
    const Fr x = v[i];
    if (!x.is_zero()) acc = acc * x;
  }
  Fr inv = acc.inverse();
  for (size_t i = hi; i-- > lo;) {
    const Fr x = v[i];
    if (x.is_zero()) continue;
    v[i] = inv * pre[i - lo];
    inv = inv * x;
  }
}

// d[j] = coset * w^j - 1   (then batch-inverted: precomputedDenominators)
template <class Fr>
__global__ void __launch_bounds__(256) k_plonk_denominators(Fr* __restrict__ d, uint32_t n, const Fr* __restrict__ tw,
                                                            Fr coset) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t half = n >> 1;
  Fr w = n == 1 ? Fr::one() : (j < half ? tw[j] : tw[j - half].neg());
  d[j] = coset * w - Fr::one();
}
#endif

}  // namespace gb200
