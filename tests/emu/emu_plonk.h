// Host walk of k_plonk_constraints (plonk.cuh) for one coset, shared by the emulation library (hostemu.cpp) and the
// mock C ABI of the PLONK orchestration test (tests/mock/mock_capi.cpp).  TEST SUPPORT, not a product path.
#pragma once
#include <vector>

#include "ntt.cuh"
#include "plonk.cuh"

namespace gb200 {

// the per-point logic of k_plonk_constraints (plonk.cuh) walked sequentially for one coset
template <class Fr>
int plonk_coset_emu(const void* const* polys, const void* abg, const void* const* blind, const int* nblind, uint32_t logn,
                    uint32_t coset_index, uint32_t rho, void* out_) {
  NttDomainHost<Fr> dom0;
  dom0.init((int)logn, nullptr, nullptr, false);
  const uint32_t n = 1u << logn;
  uint32_t log_rho = 0;
  while ((1u << log_rho) < rho) log_rho++;
  NttDomainHost<Fr> dom1;   // only for its generator
  Fr w4 = NttDomainHost<Fr>::default_generator((int)(logn + log_rho));
  const Fr g = NttDomainHost<Fr>::default_coset();
  PlonkCosetArgs<Fr> a;
  const Fr* const* P = reinterpret_cast<const Fr* const*>(polys);
  a.l = P[0]; a.r = P[1]; a.o = P[2]; a.z = P[3]; a.s1 = P[4]; a.s2 = P[5]; a.s3 = P[6];
  a.ql = P[7]; a.qr = P[8]; a.qm = P[9]; a.qo = P[10]; a.qk = P[11];
  const Fr* c3 = reinterpret_cast<const Fr*>(abg);
  a.alpha = c3[0]; a.beta = c3[1]; a.gamma = c3[2];
  Fr coset = g;
  for (uint32_t k = 0; k < coset_index; k++) coset = coset * w4;
  a.coset = coset; a.cs = g; a.css = g.sqr();
  Fr cn = coset;
  for (uint32_t k = 0; k < logn; k++) cn = cn.sqr();
  a.coset_n_minus_one = cn - Fr::one();
  a.lone_scale = a.coset_n_minus_one * dom0.ninv;
  Fr* bd[4] = {a.bl, a.br, a.bo, a.bz};
  for (int q = 0; q < 4; q++) plonk_set_blinding<Fr>(a, bd[q], reinterpret_cast<const Fr*>(blind[q]), nblind[q]);
  a.nbl = nblind[0]; a.nbr = nblind[1]; a.nbo = nblind[2]; a.nbz = nblind[3];
  a.n = n; a.logn = logn; a.rho = rho; a.log_rho = log_rho; a.coset_index = coset_index;
  std::vector<Fr> wp(n), den(n);
  wp[0] = Fr::one();
  for (uint32_t j = 1; j < n; j++) wp[j] = wp[j - 1] * dom0.gen;
  for (uint32_t j = 0; j < n; j++) den[j] = (coset * wp[j] - Fr::one()).inverse();
  a.den_inv = den.data();
  a.tw = nullptr;
  Fr* out = reinterpret_cast<Fr*>(out_);
  for (uint32_t j = 0; j < n; j++) {
    const Fr v = plonk_all_constraints<Fr>(a, j, wp[j], wp[(j + 1) % n]);
    out[ntt_bitrev(rho * j + coset_index, (int)(logn + log_rho))] = v;
  }
  return 0;
}

}  // namespace gb200
