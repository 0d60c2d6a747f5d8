// Radix-2 NTT over Fr in gnark-crypto's fft.Domain conventions - shared
// host/device logic (pass planning, tile indexing, butterflies).
//
// Replaces domain.FFT / domain.FFTInverse at backend/groth16/bn254/prove.go:362-386,
// backend/plonk/bn254/prove.go:1038,1056,1320 and ICICLE's Ntt at
// backend/accelerated/icicle/groth16/bn254/icicle.go:1425,1428,1474.
// Semantics:
//   DIF: natural in -> bit-reversed out, (a,b) -> (a+b, (a-b)*w^(p*2^q)) at stage q
//        (recursion shape: backend/groth16/bn254/mpcsetup/lagrange.go:132-169)
//   DIT: bit-reversed in -> natural out
//   inverse: w^-1 and scale 1/n;  OnCoset: coefficient j scaled by g^j before a
//   forward transform / by g^-j after an inverse one (g = FrMultiplicativeGen).
//
// Work decomposition: the log2(n) stages are grouped into passes; a pass owns S
// consecutive index bits [lo_bit, lo_bit+S) and is executed on tiles of
// 2^(S+cb) elements staged in shared memory (cb extra low bits make every global
// access a 2^cb-element contiguous run).  One global read + one global write per
// pass; the twiddle table (w^k, k < n/2) is read through L2.
#pragma once
#include <vector>
#include "field.cuh"

namespace gb200 {

enum { NTT_DIF = 0, NTT_DIT = 1 };
constexpr int NTT_MAX_TILE_LOG = 11;   // 2048 elements per tile (the kernel's launch bound)
constexpr int NTT_DEFAULT_TILE_LOG = 8; // 256 elements per tile: see NttDomainDev::init
constexpr int NTT_MAX_PASSES = 8;

struct NttPass {
  int lo_bit;  // lowest index bit transformed by this pass
  int S;       // number of stages (bits) in this pass
  int cb;      // contiguous low bits carried along for coalescing (cb <= lo_bit)
};

struct NttPlan {
  int logn;
  int npasses;
  NttPass pass[NTT_MAX_PASSES];  // in increasing bit order (DIT runs them 0.., DIF runs them reversed)
};

// passes a tile size needs for 2^logn points (first pass: tile_log stages, then <= tile_log - 2 stages per pass)
inline int ntt_passes_needed(int logn, int tile_log) {
  if (logn <= tile_log) return 1;
  const int smax = tile_log - 2;
  return 1 + (logn - tile_log + smax - 1) / smax;
}

inline NttPlan ntt_make_plan(int logn, int tile_log = NTT_MAX_TILE_LOG) {
  // a tile too small for the size would need more than NTT_MAX_PASSES passes: use the smallest tile that fits
  // (found by the randomised emulation soak, tools/fuzz_emulation.py: tile 2^3 at 2^13 points overran pass[])
  if (tile_log < 3) tile_log = 3;
  while (tile_log < NTT_MAX_TILE_LOG && ntt_passes_needed(logn, tile_log) > NTT_MAX_PASSES) tile_log++;
  NttPlan pl;
  pl.logn = logn;
  pl.npasses = 0;
  int bit = 0;
  // lowest pass: contiguous tile
  int s0 = logn < tile_log ? logn : tile_log;
  pl.pass[pl.npasses++] = NttPass{0, s0, 0};
  bit = s0;
  while (bit < logn) {
    int smax = tile_log - 2;           // keep >= 4 consecutive elements per run
    int rem = logn - bit;
    // balance the remaining bits over the remaining passes
    int np = (rem + smax - 1) / smax;
    int s = (rem + np - 1) / np;
    int cb = tile_log - s;             // fill the tile with contiguous low bits
    if (cb > bit) cb = bit;
    pl.pass[pl.npasses++] = NttPass{bit, s, cb};
    bit += s;
  }
  return pl;
}

HD uint32_t ntt_bitrev(uint32_t i, int logn) {
#ifdef __CUDA_ARCH__
  return logn ? (__brev(i) >> (32 - logn)) : 0;
#else
  uint32_t r = 0;
  for (int k = 0; k < logn; k++) { r = (r << 1) | (i & 1); i >>= 1; }
  return r;
#endif
}

// global element index of local slot `loc` of tile `tile` for pass p.
//   i = H * 2^(lo_bit+S) + M * 2^lo_bit + Lh * 2^cb + Ll
//   loc = M * 2^cb + Ll ;  tile = H * 2^(lo_bit-cb) + Lh
HD uint32_t ntt_tile_index(const NttPass& p, uint32_t tile, uint32_t loc) {
  const uint32_t Ll = loc & ((1u << p.cb) - 1u);
  const uint32_t M = loc >> p.cb;
  const int lhbits = p.lo_bit - p.cb;
  const uint32_t Lh = tile & ((1u << lhbits) - 1u);
  const uint32_t H = tile >> lhbits;
  return (H << (p.lo_bit + p.S)) | (M << p.lo_bit) | (Lh << p.cb) | Ll;
}

// twiddle exponent (index into w^k, k < n/2) of the butterfly whose lower element
// has global index i, acting on index bit beta
HD uint32_t ntt_twiddle_index(int logn, uint32_t i, int beta) {
  const uint32_t p = i & ((1u << beta) - 1u);
  return p << (logn - 1 - beta);
}

template <class Fr>
HD void ntt_bfly_dif(Fr& a, Fr& b, const Fr& w) {
  Fr t = a - b;
  a = a + b;
  b = t * w;
}
template <class Fr>
HD void ntt_bfly_dit(Fr& a, Fr& b, const Fr& w) {
  Fr t = b * w;
  b = a - t;
  a = a + t;
}

// ---------------------------------------------------------------------------
// Host-side domain: builds the tables the device domain uploads, and (for the
// CPU emulation tests) runs the same pass/tile/stage walk sequentially.
// ---------------------------------------------------------------------------
template <class Fr>
struct NttDomainHost {
  int logn = 0;
  uint32_t n = 0;
  Fr gen, gen_inv, coset, coset_inv, ninv;
  std::vector<Fr> tw, itw;       // w^k, w^-k  (k < max(n/2,1))
  std::vector<Fr> cos, icos;     // g^j ; g^-j / n   (j < n)

  static Fr pow_u64(Fr b, uint64_t e) {
    Fr r = Fr::one();
    while (e) { if (e & 1) r = r * b; b = b.sqr(); e >>= 1; }
    return r;
  }
  static Fr default_generator(int logn) {
    Fr w;
    for (int i = 0; i < Fr::N; i++) w.l[i] = Fr::Params::root_of_unity_mont(i);
    for (int k = logn; k < Fr::Params::TWO_ADICITY; k++) w = w.sqr();
    return w;
  }
  static Fr default_coset() {
    Fr g;
    for (int i = 0; i < Fr::N; i++) g.l[i] = Fr::Params::mult_gen_mont(i);
    return g;
  }

  void init(int logn_, const Fr* gen_mont, const Fr* coset_mont, bool with_coset_tables = true) {
    logn = logn_;
    n = 1u << logn;
    gen = gen_mont ? *gen_mont : default_generator(logn);
    coset = coset_mont ? *coset_mont : default_coset();
    gen_inv = gen.inverse();
    coset_inv = coset.inverse();
    Fr nn = Fr::zero();
    {  // n as a field element: one() added n times is too slow; build from bits
      Fr acc = Fr::one();
      for (int k = 0; k < logn; k++) acc = acc.dbl();
      nn = acc;
    }
    ninv = nn.inverse();
    const uint32_t half = n > 1 ? n / 2 : 1;
    tw.resize(half); itw.resize(half);
    tw[0] = Fr::one(); itw[0] = Fr::one();
    for (uint32_t k = 1; k < half; k++) { tw[k] = tw[k - 1] * gen; itw[k] = itw[k - 1] * gen_inv; }
    if (with_coset_tables) {
      cos.resize(n); icos.resize(n);
      cos[0] = Fr::one(); icos[0] = ninv;
      for (uint32_t j = 1; j < n; j++) { cos[j] = cos[j - 1] * coset; icos[j] = icos[j - 1] * coset_inv; }
    }
  }

  // sequential walk of the kernel structure (emulation)
  int tile_log = NTT_MAX_TILE_LOG;   // emulation tests lower it to walk multi-pass plans at small sizes
  void transform(Fr* data, bool inverse, int decimation, bool on_coset) const {
    NttPlan plan = ntt_make_plan(logn, tile_log);
    const std::vector<Fr>& T = inverse ? itw : tw;
    // pre-scale (forward coset)
    if (!inverse && on_coset) {
      for (uint32_t i = 0; i < n; i++) {
        uint32_t j = decimation == NTT_DIT ? ntt_bitrev(i, logn) : i;
        data[i] = data[i] * cos[j];
      }
    }
    for (int pi = 0; pi < plan.npasses; pi++) {
      const NttPass& p = decimation == NTT_DIT ? plan.pass[pi] : plan.pass[plan.npasses - 1 - pi];
      const uint32_t tile_elems = 1u << (p.S + p.cb);
      const uint32_t ntiles = n >> (p.S + p.cb);
      std::vector<Fr> loc(tile_elems);
      for (uint32_t tile = 0; tile < ntiles; tile++) {
        for (uint32_t e = 0; e < tile_elems; e++) loc[e] = data[ntt_tile_index(p, tile, e)];
        for (int k = 0; k < p.S; k++) {
          const int s = decimation == NTT_DIT ? k : p.S - 1 - k;
          const int lb = p.cb + s;
          const int beta = p.lo_bit + s;
          for (uint32_t t = 0; t < tile_elems / 2; t++) {
            const uint32_t lo = ((t >> lb) << (lb + 1)) | (t & ((1u << lb) - 1u));
            const uint32_t hi = lo | (1u << lb);
            const uint32_t gi = ntt_tile_index(p, tile, lo);
            const Fr& w = T[ntt_twiddle_index(logn, gi, beta)];
            if (decimation == NTT_DIT) ntt_bfly_dit(loc[lo], loc[hi], w);
            else ntt_bfly_dif(loc[lo], loc[hi], w);
          }
        }
        for (uint32_t e = 0; e < tile_elems; e++) data[ntt_tile_index(p, tile, e)] = loc[e];
      }
    }
    if (inverse) {
      for (uint32_t i = 0; i < n; i++) {
        if (on_coset) {
          uint32_t j = decimation == NTT_DIF ? ntt_bitrev(i, logn) : i;
          data[i] = data[i] * icos[j];
        } else {
          data[i] = data[i] * ninv;
        }
      }
    }
  }
};

}  // namespace gb200
