// TEST INFRASTRUCTURE ONLY.  Host stand-ins for the device entry points that gnark_b200/csrc/plonk_host.cu calls,
// so that the C++ PLONK orchestration (stage order, layouts, blinding patches, linearised-polynomial algebra,
// slice arithmetic) can be exercised on a machine without a GPU: plonk_host.cu is compiled as plain C++ and linked
// with this file into libgb200_plonkmock.so (tests/test_plonk_host_cpu.py).  "Device" memory is host memory, the
// stream is null, every stub computes the documented result of its entry point with the host code paths of the
// same field / curve templates (BN254 and BLS12-381: 4-limb Fr with 4- and 6-limb Fp).  The product library never contains any of this.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "capi_common.h"
#include "curve.cuh"
#include "emu_plonk.h"
#include "ntt.cuh"
#include "params_gen.cuh"

using namespace gb200;

namespace {
std::string g_err;
DeviceCtx g_ctx;
}

namespace gb200 {
int32_t set_error(const std::string& msg) { g_err = msg; return 1; }
int32_t cuda_fail(const char* what, cudaError_t) { g_err = std::string("cuda: ") + what; return 2; }
int32_t device_ctx(int, DeviceCtx** out) { g_ctx.ready = true; *out = &g_ctx; return 0; }
const NttOps* get_ntt_ops(int curve) {
  static NttOps o[2]{};
  if (curve == 0) {
    o[0].fr_bytes = sizeof(Fp<bn254_fr_params>);
    o[0].two_adicity = bn254_fr_params::TWO_ADICITY;
    o[0].host_fr = []() -> const HostFrCtx* { static const HostFrCtx c = HostFrCtx::make<bn254_fr_params>(); return &c; };
    return &o[0];
  }
  if (curve == 1) {
    o[1].fr_bytes = sizeof(Fp<bls12_381_fr_params>);
    o[1].two_adicity = bls12_381_fr_params::TWO_ADICITY;
    o[1].host_fr = []() -> const HostFrCtx* { static const HostFrCtx c = HostFrCtx::make<bls12_381_fr_params>(); return &c; };
    return &o[1];
  }
  return nullptr;
}
const MsmOps* get_msm_ops(int curve, int group) {
  static MsmOps o[2]{};
  if (group != 1 || curve < 0 || curve > 1) return nullptr;
  if (curve == 0) { o[0].scalar_bits = 254; o[0].fr_bytes = 32; o[0].affine_bytes = sizeof(Affine<Fp<bn254_fp_params>>); o[0].jac_bytes = sizeof(Jacobian<Fp<bn254_fp_params>>); }
  else { o[1].scalar_bits = 255; o[1].fr_bytes = 32; o[1].affine_bytes = sizeof(Affine<Fp<bls12_381_fp_params>>); o[1].jac_bytes = sizeof(Jacobian<Fp<bls12_381_fp_params>>); }
  return &o[curve];
}
int32_t scratch_alloc(int, size_t bytes, void** out) { *out = malloc(bytes ? bytes : 1); return *out ? 0 : 1; }
int32_t scratch_free(int, void* p) { free(p); return 0; }
}  // namespace gb200

extern "C" {

// the two CUDA runtime calls plonk_host.cu makes directly
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t count, enum cudaMemcpyKind, cudaStream_t) {
  memmove(dst, src, count);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t count, cudaStream_t) { memset(p, v, count); return cudaSuccess; }
// ~b200_table_s (capi_common.h) releases its device buffers; the mock's tables null theirs before delete
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }

const char* b200_last_error(void) { return g_err.c_str(); }
int32_t b200_alloc(int32_t, size_t bytes, void** out) { *out = malloc(bytes ? bytes : 1); return *out ? 0 : 1; }
int32_t b200_free(int32_t, void* p) { free(p); return 0; }
int32_t b200_h2d(int32_t, void* d, const void* s, size_t n) { memcpy(d, s, n); return 0; }
int32_t b200_d2h(int32_t, void* d, const void* s, size_t n) { memcpy(d, s, n); return 0; }
int32_t b200_sync(int32_t) { return 0; }

}  // extern "C"

template <class Fr, class Fq>
struct Mock {
static int32_t ntt_domain_new(int32_t dev, int32_t curve, uint32_t log2n, const void* gen, const void* coset, b200_domain_t* out) {
    NttDomainHost<Fr>* h = new NttDomainHost<Fr>();
  h->init((int)log2n, (const Fr*)gen, (const Fr*)coset);
  b200_domain_s* d = new b200_domain_s();
  d->dev = dev; d->curve = curve; d->logn = (int)log2n; d->ops = nullptr; d->impl = h;
  *out = d;
  return 0;
}
static int32_t ntt_domain_free(b200_domain_t d) { if (d) { delete (NttDomainHost<Fr>*)d->impl; delete d; } return 0; }
static int32_t ntt_async(b200_domain_t d, void* data, int32_t inverse, int32_t decimation, int32_t on_coset) {
  ((NttDomainHost<Fr>*)d->impl)->transform((Fr*)data, inverse != 0, decimation, on_coset != 0);
  return 0;
}
static int32_t vec_bit_reverse(int32_t, int32_t, void* data, uint32_t log2n) {
  Fr* a = (Fr*)data;
  const uint32_t n = 1u << log2n;
  for (uint32_t i = 0; i < n; i++) { const uint32_t j = ntt_bitrev(i, (int)log2n); if (i < j) { Fr t = a[i]; a[i] = a[j]; a[j] = t; } }
  return 0;
}
static int32_t vec_axpy(int32_t, int32_t, void* y, const void* a, const void* x, size_t n) {
  Fr* Y = (Fr*)y; const Fr* X = (const Fr*)x; const Fr s = *(const Fr*)a;
  for (size_t i = 0; i < n; i++) Y[i] = Y[i] + s * X[i];
  return 0;
}
// iop.BuildRatioCopyConstraint as documented in include/gnark_b200.h
static int32_t plonk_build_z(b200_domain_t d0, const void* l, const void* r, const void* o, const int64_t* perm,
                           const void* beta_, const void* gamma_, void* z_) {
  NttDomainHost<Fr>* dom = (NttDomainHost<Fr>*)d0->impl;
  const uint32_t n = dom->n;
  const Fr beta = *(const Fr*)beta_, gamma = *(const Fr*)gamma_, g = dom->coset;
  std::vector<Fr> supp(3 * (size_t)n);
  Fr wp = Fr::one();
  for (uint32_t i = 0; i < n; i++) { supp[i] = wp; supp[n + i] = g * wp; supp[2 * (size_t)n + i] = g * g * wp; wp = wp * dom->gen; }
  const Fr* f[3] = {(const Fr*)l, (const Fr*)r, (const Fr*)o};
  Fr* z = (Fr*)z_;
  z[0] = Fr::one();
  for (uint32_t i = 0; i + 1 < n; i++) {
    Fr num = Fr::one(), den = Fr::one();
    for (int j = 0; j < 3; j++) {
      num = num * (f[j][i] + beta * supp[(size_t)j * n + i] + gamma);
      den = den * (f[j][i] + beta * supp[(size_t)perm[(size_t)j * n + i]] + gamma);
    }
    z[i + 1] = z[i] * num * den.inverse();
  }
  return 0;
}
static int32_t plonk_constraints_coset(b200_domain_t d0, const void* big_coset_gen, const void* big_gen,
                                     const b200_plonk_coset_args* a) {
  NttDomainHost<Fr>* dom = (NttDomainHost<Fr>*)d0->impl;
  // the handle must be domain0 with THIS coset's generator g * w4^i
  Fr coset = *(const Fr*)big_coset_gen;
  for (uint32_t k = 0; k < a->coset_index; k++) coset = coset * *(const Fr*)big_gen;
  if (!(coset == dom->coset)) return set_error("mock: constraint call on a domain handle of another coset");
  const void* polys[12] = {a->l, a->r, a->o, a->z, a->s1, a->s2, a->s3, a->ql, a->qr, a->qm, a->qo, a->qk};
  Fr abg[3] = {*(const Fr*)a->alpha, *(const Fr*)a->beta, *(const Fr*)a->gamma};
  const void* blind[4] = {a->bl, a->br, a->bo, a->bz};
  const int nb[4] = {a->nbl, a->nbr, a->nbo, a->nbz};
  return plonk_coset_emu<Fr>(polys, abg, blind, nb, (uint32_t)dom->logn, a->coset_index, a->rho, a->out);
}
static int32_t plonk_bsb22_coset(b200_domain_t d0, const void* qcp, const void* pi2, uint32_t coset_index, uint32_t rho, void* out) {
  NttDomainHost<Fr>* dom = (NttDomainHost<Fr>*)d0->impl;
  uint32_t log_rho = 0;
  while ((1u << log_rho) < rho) log_rho++;
  const Fr* Q = (const Fr*)qcp; const Fr* PI = (const Fr*)pi2; Fr* O = (Fr*)out;
  for (uint32_t j = 0; j < dom->n; j++) {
    const uint32_t k = ntt_bitrev(rho * j + coset_index, dom->logn + (int)log_rho);
    O[k] = O[k] + Q[j] * PI[j];
  }
  return 0;
}
// r[i] *= 1/(X^n - 1) on the big coset, then FFTInverse(DIT, OnCoset) on domain1
static int32_t plonk_divide_by_zh(b200_domain_t d1, uint32_t log_n0, void* data) {
  NttDomainHost<Fr>* dom = (NttDomainHost<Fr>*)d1->impl;
  const uint32_t m = dom->n, n = 1u << log_n0, rho = m / n;
  Fr gn = dom->coset, wn = dom->gen;
  for (uint32_t k = 0; k < log_n0; k++) { gn = gn.sqr(); wn = wn.sqr(); }
  std::vector<Fr> tab(rho);
  Fr cur = gn;
  for (uint32_t i = 0; i < rho; i++) { tab[i] = (cur - Fr::one()).inverse(); cur = cur * wn; }
  Fr* a = (Fr*)data;
  for (uint32_t i = 0; i < m; i++) a[i] = a[i] * tab[ntt_bitrev(i, dom->logn) % rho];
  dom->transform(a, true, B200_DIT, true);
  return 0;
}
static int32_t poly_eval(int32_t, int32_t, const void* c, size_t n, const void* x, void* out) {
  const Fr* C = (const Fr*)c; const Fr X = *(const Fr*)x;
  Fr acc = Fr::zero();
  for (size_t i = n; i-- > 0;) acc = acc * X + C[i];
  *(Fr*)out = acc;
  return 0;
}
static int32_t poly_div_by_linear(int32_t, int32_t, void* c, size_t n, const void* z, void* rem) {
  Fr* C = (Fr*)c; const Fr Z = *(const Fr*)z;
  Fr acc = Fr::zero();
  std::vector<Fr> q(n ? n : 1, Fr::zero());
  for (size_t i = n; i-- > 1;) { acc = C[i] + Z * acc; q[i - 1] = acc; }
  *(Fr*)rem = n ? C[0] + Z * acc : Fr::zero();
  for (size_t i = 0; i < n; i++) C[i] = q[i];          // quotient in [0, n-1), top slot cleared
  return 0;
}
static int32_t table_upload(int32_t dev, int32_t curve, int32_t group, const void* pts, size_t n, int32_t, b200_table_t* out) {
  if (group != 1) return set_error("mock: G1 only");
  b200_table_s* t = new b200_table_s();
  t->dev = dev; t->curve = curve; t->group = group; t->n = n; t->ops = nullptr;
  t->d_points = malloc(n * sizeof(Affine<Fq>));
  memcpy(t->d_points, pts, n * sizeof(Affine<Fq>));
  *out = t;
  return 0;
}
static int32_t table_free(b200_table_t t) { if (t) { free(t->d_points); t->d_points = nullptr; delete t; } return 0; }
// sum s_i P_i by double-and-add (small n)
static int32_t msm_g1(b200_table_t t, size_t off, size_t n, const void* scalars, int32_t, void* out) {
  if (off + n > t->n) return set_error("mock: msm range");
  const Affine<Fq>* P = (const Affine<Fq>*)t->d_points + off;
  const Fr* S = (const Fr*)scalars;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  for (size_t i = 0; i < n; i++) {
    const Fr s = S[i].from_mont();
    XYZZ<Fq> q = XYZZ<Fq>::inf();
    for (int w = Fr::N - 1; w >= 0; w--)
      for (int b = 31; b >= 0; b--) { q.dbl(); if ((s.l[w] >> b) & 1) q.add_mixed(P[i]); }
    acc.add(q);
  }
  *(Jacobian<Fq>*)out = acc.to_jacobian();
  return 0;
}
};

#define DISPATCH(curve, CALL)                                                        \
  do {                                                                               \
    if ((curve) == 0) return Mock<Fp<bn254_fr_params>, Fp<bn254_fp_params>>::CALL;   \
    if ((curve) == 1) return Mock<Fp<bls12_381_fr_params>, Fp<bls12_381_fp_params>>::CALL; \
    return set_error("mock: BN254 / BLS12-381 only");                                \
  } while (0)

extern "C" {
int32_t b200_ntt_domain_new(int32_t dev, int32_t curve, uint32_t log2n, const void* gen, const void* coset, b200_domain_t* out) { DISPATCH(curve, ntt_domain_new(dev, curve, log2n, gen, coset, out)); }
int32_t b200_ntt_domain_free(b200_domain_t d) { if (!d) return 0; DISPATCH(d->curve, ntt_domain_free(d)); }
int32_t b200_ntt_async(b200_domain_t d, void* data, int32_t inverse, int32_t decimation, int32_t on_coset) { DISPATCH(d->curve, ntt_async(d, data, inverse, decimation, on_coset)); }
int32_t b200_vec_bit_reverse(int32_t dev, int32_t curve, void* data, uint32_t log2n) { DISPATCH(curve, vec_bit_reverse(dev, curve, data, log2n)); }
int32_t b200_vec_axpy(int32_t dev, int32_t curve, void* y, const void* a, const void* x, size_t n) { DISPATCH(curve, vec_axpy(dev, curve, y, a, x, n)); }
int32_t b200_plonk_build_z(b200_domain_t d0, const void* l, const void* r, const void* o, const int64_t* perm, const void* beta, const void* gamma, void* z) { DISPATCH(d0->curve, plonk_build_z(d0, l, r, o, perm, beta, gamma, z)); }
int32_t b200_plonk_constraints_coset(b200_domain_t d0, const void* g, const void* w4, const b200_plonk_coset_args* a) { DISPATCH(d0->curve, plonk_constraints_coset(d0, g, w4, a)); }
int32_t b200_plonk_bsb22_coset(b200_domain_t d0, const void* qcp, const void* pi2, uint32_t ci, uint32_t rho, void* out) { DISPATCH(d0->curve, plonk_bsb22_coset(d0, qcp, pi2, ci, rho, out)); }
int32_t b200_plonk_divide_by_zh(b200_domain_t d1, uint32_t log_n0, void* data) { DISPATCH(d1->curve, plonk_divide_by_zh(d1, log_n0, data)); }
int32_t b200_poly_eval(int32_t dev, int32_t curve, const void* c, size_t n, const void* x, void* out) { DISPATCH(curve, poly_eval(dev, curve, c, n, x, out)); }
int32_t b200_poly_div_by_linear(int32_t dev, int32_t curve, void* c, size_t n, const void* z, void* rem) { DISPATCH(curve, poly_div_by_linear(dev, curve, c, n, z, rem)); }
int32_t b200_table_upload(int32_t dev, int32_t curve, int32_t group, const void* pts, size_t n, int32_t flags, b200_table_t* out) { DISPATCH(curve, table_upload(dev, curve, group, pts, n, flags, out)); }
int32_t b200_table_free(b200_table_t t) { if (!t) return 0; DISPATCH(t->curve, table_free(t)); }
int32_t b200_msm_g1(b200_table_t t, size_t off, size_t n, const void* scalars, int32_t on_dev, void* out) { DISPATCH(t->curve, msm_g1(t, off, n, scalars, on_dev, out)); }
// the pipelined MSM of the library (device scalars -> device result, joined later): here computed on the spot
int32_t b200_msm_pipelined(b200_table_t t, size_t off, size_t n, const void* d_scalars, void* d_out) { DISPATCH(t->curve, msm_g1(t, off, n, d_scalars, 1, d_out)); }
int32_t b200_msm_join(int32_t) { return 0; }
}  // extern "C"
