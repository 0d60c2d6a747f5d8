// Batched-affine tree reduction of the sorted bucket entries (opt-in: GB200_MSM_BATCH_AFFINE=<levels>).
//
// The default accumulate kernel sums a bucket's entries serially into an XYZZ accumulator: 8M + 2S per entry,
// and it runs the integer-multiplier pipe at 86 % of its peak (profiles/r01_ncu_accumulate_summary.md), so the
// only way to go faster is fewer multiplications per entry.  gnark-crypto's MultiExp gets there on the CPU with
// AFFINE buckets and batched additions; this is the data-parallel form of the same idea:
//
//   level l:  inside every bucket, entries (2j, 2j+1) are added as AFFINE points -> entry j of level l+1
//             (an odd last entry passes through), so a bucket of k entries becomes ceil(k/2);
//   a thread owns MSM_BA_BATCH consecutive OUTPUT entries and shares ONE field inversion between their slopes
//   (Montgomery's trick: 3 multiplications per entry), the inversion itself is the shift-and-add binary GCD
//   (Fp::inverse_gcd) that runs on the ALU pipe, not on the multiplier;
//   cost per addition: 3 (batch) + 1 (lambda) + 1S + 1M = 5M + 1S instead of 8M + 2S.
//
// After `levels` levels the surviving ceil(k / 2^levels) entries per bucket go through the unchanged XYZZ
// pipeline (tasks -> combine -> reduce), which also keeps its heavy-bucket handling.  Every special case of an
// affine addition is handled: P = Q (tangent slope), P = -Q (infinity), P or Q at infinity.
// The templates are HD; tests/test_emulation.py walks them on the CPU (hostemu.cpp).
#pragma once
#include "msm.cuh"

namespace gb200 {

constexpr int MSM_BA_BATCH = 32;

// level-0 source: the base table gathered through the sorted entry values (index | sign << 31)
template <class F>
struct BaSrcTable {
  const Affine<F>* table;
  const uint32_t* svals;
  HD Affine<F> load(uint32_t i) const { return msm_load_point(table, svals[i]); }
};
// later levels: the previous level's output
template <class F>
struct BaSrcPoints {
  const Affine<F>* pts;
  HD Affine<F> load(uint32_t i) const { return pts[i]; }
};

// kinds of an output entry
enum { BA_COPY_P = 0, BA_COPY_Q = 1, BA_INF = 2, BA_ADD = 3 };

// slope of P + Q as num / den; returns the kind (den is only meaningful for BA_ADD)
template <class F>
HD int msm_ba_classify(const Affine<F>& P, const Affine<F>& Q, bool has_q, F& num, F& den) {
  if (!has_q || Q.is_inf()) return BA_COPY_P;
  if (P.is_inf()) return BA_COPY_Q;
  if (P.x == Q.x) {
    if (P.y == Q.y && !P.y.is_zero()) {      // tangent: 3 x^2 / 2 y
      const F xx = P.x.sqr();
      num = xx.dbl() + xx;
      den = P.y.dbl();
      return BA_ADD;
    }
    return BA_INF;                           // P = -Q (or a 2-torsion point, which these curves do not have)
  }
  num = Q.y - P.y;
  den = Q.x - P.x;
  return BA_ADD;
}

// One thread: output entries [o_begin, o_end) (at most MSM_BA_BATCH; msm_ba_batch_for) of a level.
//   off_in / off_out: bucket offsets of the input / output level (nb + 1 entries each)
template <class F, class SRC>
HD void msm_ba_level_thread(const SRC& src, const uint32_t* off_in, const uint32_t* off_out, uint32_t nb,
                            uint32_t o_begin, uint32_t o_end, Affine<F>* out) {
  F pre[MSM_BA_BATCH];
  // bucket of o_begin: largest b with off_out[b] <= o_begin  (empty buckets have off_out[b] == off_out[b+1])
  uint32_t lo = 0, hi = nb;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo + 1) >> 1);
    if (off_out[mid] <= o_begin) lo = mid; else hi = mid - 1;
  }
  const uint32_t b_first = lo;
  // pass 1: denominators, running product
  F acc = F::one();
  uint32_t b = b_first;
  for (uint32_t o = o_begin; o < o_end; o++) {
    while (o >= off_out[b + 1]) b++;
    const uint32_t j = o - off_out[b];
    const uint32_t i0 = off_in[b] + 2 * j;
    const bool has_q = i0 + 1 < off_in[b + 1];
    const Affine<F> P = src.load(i0);
    Affine<F> Q = P;
    if (has_q) Q = src.load(i0 + 1);
    F num, den;
    const int kind = msm_ba_classify<F>(P, Q, has_q, num, den);
    pre[o - o_begin] = acc;
    if (kind == BA_ADD) acc = acc * den;
  }
  F inv = acc.inverse_gcd();
  // pass 2 (backwards): slopes and results
  for (uint32_t o = o_end; o-- > o_begin;) {
    while (o < off_out[b]) b--;
    const uint32_t j = o - off_out[b];
    const uint32_t i0 = off_in[b] + 2 * j;
    const bool has_q = i0 + 1 < off_in[b + 1];
    const Affine<F> P = src.load(i0);
    Affine<F> Q = P;
    if (has_q) Q = src.load(i0 + 1);
    F num, den;
    const int kind = msm_ba_classify<F>(P, Q, has_q, num, den);
    Affine<F> R;
    if (kind == BA_COPY_P) R = P;
    else if (kind == BA_COPY_Q) R = Q;
    else if (kind == BA_INF) R = Affine<F>::inf();
    else {
      const F dinv = inv * pre[o - o_begin];
      inv = inv * den;
      const F lam = num * dinv;
      R.x = lam.sqr() - P.x - Q.x;
      R.y = lam * (P.x - R.x) - P.y;
    }
    out[o] = R;
  }
}

// outputs per thread (= additions sharing one inversion) for a level with at most `bound` outputs: the full
// batch while that still leaves >= ~150k threads (148 SMs x 8 blocks x 128), smaller batches for the small deep levels
HD uint32_t msm_ba_batch_for(size_t bound) {
  size_t b = bound / 150000;
  if (b > (size_t)MSM_BA_BATCH) b = MSM_BA_BATCH;
  if (b < 8) b = 8;
  return (uint32_t)b;
}

// next level's per-bucket count
HD uint32_t msm_ba_next_count(uint32_t k) { return (k + 1) >> 1; }

}  // namespace gb200
