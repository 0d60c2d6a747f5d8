// Fixed-base batch scalar multiplication: out[i] = k_i * B for one base B and n scalars
// (gnark-crypto's curve.BatchScalarMultiplicationG1/G2; reference call sites: Groth16 Setup
// backend/groth16/bn254/setup.go:233,302, Lagrange/monomial SRS generation test/unsafekzg/kzgsrs.go:198,
// proof assembly prove.go:185 - SURVEY.md §8(f)-3).
//
//   table[w][d] = (d+1) * 2^(c*w) * B   affine, d < 2^(c-1), w < nwin = bits/c + 1
//   k = sum_w digit_w 2^(c*w), signed digits in [-2^(c-1), 2^(c-1))  (top window absorbs the carry)
//   out_i = sum_w sign(digit_w) * table[w][|digit_w| - 1]              nwin mixed additions, no doublings
//
// Both the table and the results go from XYZZ to affine through one batched inversion per FB_CHUNK points
// (Montgomery's trick), so an output point costs nwin mixed adds + ~6 multiplications instead of an inversion.
// The templates are HD: tests/test_emulation.py runs them on the CPU through hostemu.cpp.
#pragma once
#include "curve.cuh"

namespace gb200 {

constexpr int FB_CHUNK = 16;        // points per batched inversion
constexpr int FB_MAX_WINDOWS = 192; // c >= 2 and 377-bit scalars: 189 windows

struct FixedBasePlan {
  int c, nwin;
  uint32_t half;                    // table entries per window = 2^(c-1)
};
HD FixedBasePlan fixed_base_plan(int scalar_bits, int c) {
  FixedBasePlan p;
  p.c = c;
  p.nwin = scalar_bits / c + 1;
  p.half = 1u << (c - 1);
  return p;
}
// window size by batch size: table cost nwin * 2^(c-1) * ~1.5c additions against n * nwin for the batch
HD int fixed_base_window_for(size_t n) {
  int lg = 0;
  while (((size_t)1 << (lg + 1)) <= n) lg++;
  int c = lg - 6;
  return c < 2 ? 2 : (c > 16 ? 16 : c);
}

// 2^(c*w) * B
template <class F>
HD XYZZ<F> fixed_base_window_base(const Affine<F>& B, int c, int w) {
  XYZZ<F> q = XYZZ<F>::from_affine(B);
  for (int k = 0; k < c * w; k++) q.dbl();
  return q;
}
// (d+1) * Pw, Pw affine
template <class F>
HD XYZZ<F> fixed_base_entry(const Affine<F>& Pw, uint32_t d) {
  XYZZ<F> acc = XYZZ<F>::inf();
  const uint32_t k = d + 1;
  int top = 31;
  while (top > 0 && !((k >> top) & 1)) top--;
  for (int bit = top; bit >= 0; bit--) {
    acc.dbl();
    if ((k >> bit) & 1) acc.add_mixed(Pw);
  }
  return acc;
}

// k * B from the table (scalar in Montgomery form, as fr.Element)
template <class Fr, class F>
HD XYZZ<F> fixed_base_eval(const FixedBasePlan& pl, const Affine<F>* table, const Fr& k_mont) {
  constexpr int N = Fr::N;
  const Fr s = k_mont.from_mont();
  XYZZ<F> acc = XYZZ<F>::inf();
  uint32_t carry = 0;
  const uint32_t mask = (1u << pl.c) - 1u;
  for (int w = 0; w < pl.nwin; w++) {
    const int bit = w * pl.c;
    const int limb = bit >> 5, sh = bit & 31;
    uint32_t raw = 0;
    if (limb < N) {
      raw = s.l[limb] >> sh;
      if (sh + pl.c > 32 && limb + 1 < N) raw |= s.l[limb + 1] << (32 - sh);
    }
    raw &= mask;
    uint32_t d = raw + carry;
    bool neg = false;
    carry = 0;
    if (d >= pl.half && w != pl.nwin - 1) { d = (1u << pl.c) - d; neg = true; carry = 1; }
    if (d == 0) continue;
    // the top window holds bits - (nwin-1)*c < c bits plus a carry: d <= 2^(c-1) always fits the table
    Affine<F> p = table[(size_t)w * pl.half + (d - 1)];
    if (neg) p.y = p.y.neg();
    acc.add_mixed(p);
  }
  return acc;
}

// cnt <= FB_CHUNK points: XYZZ -> affine with ONE inversion.  x = X/ZZ, y = Y/ZZZ; with zi = 1/ZZZ:
// 1/ZZ = (zi * ZZ)^2 (ZZ^3 = ZZZ^2).  Points at infinity (ZZZ = 0) are skipped in the product.
template <class F>
HD void xyzz_batch_to_affine(const XYZZ<F>* in, Affine<F>* out, uint32_t cnt) {
  F pre[FB_CHUNK];
  F acc = F::one();
  for (uint32_t i = 0; i < cnt; i++) {
    pre[i] = acc;
    if (!in[i].is_inf()) acc = acc * in[i].zzz;
  }
  F inv = acc.inverse();
  for (uint32_t i = cnt; i-- > 0;) {
    if (in[i].is_inf()) { out[i] = Affine<F>::inf(); continue; }
    const F zi = inv * pre[i];
    inv = inv * in[i].zzz;
    const F zi2 = (zi * in[i].zz).sqr();
    Affine<F> a;
    a.x = in[i].x * zi2;
    a.y = in[i].y * zi;
    out[i] = a;
  }
}

#ifdef __CUDACC__
template <class F>
__global__ void __launch_bounds__(64) k_fb_window_bases(Affine<F> B, int c, int nwin, Affine<F>* __restrict__ Pw) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwin) return;
  Pw[w] = fixed_base_window_base<F>(B, c, w).to_affine();
}
template <class F>
__global__ void __launch_bounds__(128) k_fb_table(FixedBasePlan pl, const Affine<F>* __restrict__ Pw,
                                                  XYZZ<F>* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)pl.nwin * pl.half) return;
  const uint32_t w = (uint32_t)(t / pl.half), d = (uint32_t)(t % pl.half);
  out[t] = fixed_base_entry<F>(Pw[w], d);
}
template <class Fr, class F>
__global__ void __launch_bounds__(128) k_fb_eval(FixedBasePlan pl, const Affine<F>* __restrict__ table,
                                                 const Fr* __restrict__ scalars, size_t n, XYZZ<F>* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = fixed_base_eval<Fr, F>(pl, table, scalars[i]);
}
template <class F>
__global__ void __launch_bounds__(128) k_fb_to_affine(const XYZZ<F>* __restrict__ in, Affine<F>* __restrict__ out, size_t n) {
  const size_t first = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * FB_CHUNK;
  if (first >= n) return;
  const uint32_t cnt = (uint32_t)(n - first < (size_t)FB_CHUNK ? n - first : (size_t)FB_CHUNK);
  xyzz_batch_to_affine<F>(in + first, out + first, cnt);
}

// d_scalars: n fr.Elements (Montgomery) on the device; d_out: n affine points on the device
template <class Fr, class F>
cudaError_t fixed_base_enqueue(cudaStream_t st, const Affine<F>& base, const Fr* d_scalars, size_t n, int c,
                               Affine<F>* d_out) {
  if (n == 0) return cudaSuccess;
  const FixedBasePlan pl = fixed_base_plan(Fr::Params::BITS, c);
  if (pl.nwin > FB_MAX_WINDOWS) return cudaErrorInvalidValue;
  const size_t entries = (size_t)pl.nwin * pl.half;
  const size_t tmp_pts = entries > n ? entries : n;
  Affine<F>* Pw = nullptr;
  Affine<F>* table = nullptr;
  XYZZ<F>* tmp = nullptr;
  cudaError_t e;
  auto release = [&]() {                 // stream-ordered: safe after the kernels above it were enqueued
    if (tmp) cudaFreeAsync(tmp, st);
    if (table) cudaFreeAsync(table, st);
    if (Pw) cudaFreeAsync(Pw, st);
  };
  if ((e = cudaMallocAsync(&Pw, (size_t)pl.nwin * sizeof(Affine<F>), st)) != cudaSuccess) return e;
  if ((e = cudaMallocAsync(&table, entries * sizeof(Affine<F>), st)) != cudaSuccess) { release(); return e; }
  if ((e = cudaMallocAsync(&tmp, tmp_pts * sizeof(XYZZ<F>), st)) != cudaSuccess) { release(); return e; }
  k_fb_window_bases<F><<<(pl.nwin + 63) / 64, 64, 0, st>>>(base, pl.c, pl.nwin, Pw);
  k_fb_table<F><<<(unsigned)((entries + 127) / 128), 128, 0, st>>>(pl, Pw, tmp);
  const size_t g1 = (entries + FB_CHUNK - 1) / FB_CHUNK;
  k_fb_to_affine<F><<<(unsigned)((g1 + 127) / 128), 128, 0, st>>>(tmp, table, entries);
  k_fb_eval<Fr, F><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pl, table, d_scalars, n, tmp);
  const size_t g2 = (n + FB_CHUNK - 1) / FB_CHUNK;
  k_fb_to_affine<F><<<(unsigned)((g2 + 127) / 128), 128, 0, st>>>(tmp, d_out, n);
  e = cudaGetLastError();
  release();
  return e;
}
#endif  // __CUDACC__

}  // namespace gb200
