// CPU oracle, C++ leg.  TEST INFRASTRUCTURE ONLY: imported by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs -
// never by the product (gnark_b200/).
//
// PARITY PARTLY PINNED (oracle/params.py header has the list): the reference's arithmetic
// (github.com/consensys/gnark-crypto v0.21.0, go.mod:9) is not in /root/reference and it
// holds no MSM / NTT result vectors as such; the external fixtures it does ship - the
// Ethereum KZG ceremony SRS (BLS12-381 G1 MSM 4096 + Fr NTT 2^12) and gnark's serialised
// verifying keys (BN254 / BLS12-381 constants) - are checked in tests/test_golden_kzg.py,
// this file included.  Everything else is property-anchored.  This file restates the PUBLISHED ALGORITHMS that module
// uses for the calls the reference makes, and is itself validated against the
// big-int Python oracle (oracle/ec.py, oracle/ntt.py) in tests/test_oracle_c.py:
//   * MultiExp (call sites backend/groth16/bn254/prove.go:194,207,227,237,283):
//     signed-digit windowed Pippenger, extended-Jacobian (XYZZ) buckets, one task
//     per window (+ point-range splitting when threads > windows), running-sum
//     bucket reduction, Horner over windows.
//   * fft.Domain.FFT/FFTInverse (prove.go:362-386): recursive radix-2 DIF/DIT in
//     the shape of backend/groth16/bn254/mpcsetup/lagrange.go:132-169.
//   * BatchScalarMultiplicationG1/G2 (backend/groth16/bn254/setup.go:233,302):
//     windowed fixed-base table + batch inversion.
// Field layout = gnark's ([Limbs]uint64 little-endian, Montgomery R = 2^(64*Limbs)).
// Moduli are passed in by the caller (oracle/params.py); nothing is hard-coded.
#include <algorithm>
#include <atomic>
#include <functional>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

typedef uint64_t u64;
typedef unsigned __int128 u128;

namespace {

template <int N>
struct Ctx {
  u64 p[N], one[N], r2[N], ninv;
};

template <int N>
void ctx_init(Ctx<N>& c, const u64* mod) {
  for (int i = 0; i < N; i++) c.p[i] = mod[i];
  if (mod[N - 1] >> 63) { fprintf(stderr, "oracle: modulus with the top bit set is not supported by mul()\n"); abort(); }
  u64 x = 1;
  for (int i = 0; i < 6; i++) x *= 2 - mod[0] * x;  // p^-1 mod 2^64
  c.ninv = 0 - x;
  // one = 2^(64N) mod p, r2 = 2^(128N) mod p by repeated doubling
  u64 t[N];
  for (int i = 0; i < N; i++) t[i] = 0;
  t[0] = 1;
  auto dbl = [&](u64* a) {
    u64 carry = 0;
    for (int i = 0; i < N; i++) { u64 nc = a[i] >> 63; a[i] = (a[i] << 1) | carry; carry = nc; }
    bool ge = carry != 0;
    if (!ge) {
      ge = true;
      for (int i = N - 1; i >= 0; i--) { if (a[i] > c.p[i]) break; if (a[i] < c.p[i]) { ge = false; break; } }
    }
    if (ge) { u128 br = 0; for (int i = 0; i < N; i++) { u128 d = (u128)a[i] - c.p[i] - br; a[i] = (u64)d; br = (d >> 64) & 1; } }
  };
  for (int k = 0; k < 64 * N; k++) dbl(t);
  for (int i = 0; i < N; i++) c.one[i] = t[i];
  for (int k = 0; k < 64 * N; k++) dbl(t);
  for (int i = 0; i < N; i++) c.r2[i] = t[i];
}

// prime field element; TAG separates the base field (0) from the scalar field (1)
template <int N, int TAG>
struct Fp {
  static Ctx<N> C;
  static const int DEG = 1;
  u64 v[N];
  static Fp zero() { Fp r; memset(r.v, 0, sizeof(r.v)); return r; }
  static Fp one() { Fp r; memcpy(r.v, C.one, sizeof(r.v)); return r; }
  bool is_zero() const { u64 t = 0; for (int i = 0; i < N; i++) t |= v[i]; return t == 0; }
  bool eq(const Fp& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
  static bool ge_p(const u64* a) {
    for (int i = N - 1; i >= 0; i--) { if (a[i] > C.p[i]) return true; if (a[i] < C.p[i]) return false; }
    return true;
  }
  static void sub_p(u64* a) { u128 br = 0; for (int i = 0; i < N; i++) { u128 d = (u128)a[i] - C.p[i] - br; a[i] = (u64)d; br = (d >> 64) & 1; } }
  Fp add(const Fp& o) const {
    Fp r; u128 c = 0;
    for (int i = 0; i < N; i++) { c += (u128)v[i] + o.v[i]; r.v[i] = (u64)c; c >>= 64; }
    if (c || ge_p(r.v)) sub_p(r.v);
    return r;
  }
  Fp sub(const Fp& o) const {
    Fp r; u128 br = 0;
    for (int i = 0; i < N; i++) { u128 d = (u128)v[i] - o.v[i] - br; r.v[i] = (u64)d; br = (d >> 64) & 1; }
    if (br) { u128 c = 0; for (int i = 0; i < N; i++) { c += (u128)r.v[i] + C.p[i]; r.v[i] = (u64)c; c >>= 64; } }
    return r;
  }
  Fp neg() const { return is_zero() ? *this : zero().sub(*this); }
  Fp dbl() const { return add(*this); }
  // CIOS Montgomery product; every modulus here leaves the top bit of the top word clear, so the
  // two carry words of the textbook CIOS collapse into one ("no-carry" variant: t stays < 2p)
  Fp mul(const Fp& o) const {
    u64 t[N];
    for (int i = 0; i < N; i++) {
      u128 A = (u128)v[0] * o.v[i] + (i ? t[0] : 0);
      const u64 t0 = (u64)A; A >>= 64;
      const u64 m = t0 * C.ninv;
      u128 Cc = ((u128)m * C.p[0] + t0) >> 64;
      for (int j = 1; j < N; j++) {
        A += (u128)v[j] * o.v[i] + (i ? t[j] : 0);
        Cc += (u128)m * C.p[j] + (u64)A; A >>= 64;
        t[j - 1] = (u64)Cc; Cc >>= 64;
      }
      t[N - 1] = (u64)Cc + (u64)A;
    }
    Fp r; memcpy(r.v, t, sizeof(r.v));
    if (ge_p(r.v)) sub_p(r.v);
    return r;
  }
  Fp sqr() const { return mul(*this); }
  Fp from_mont() const { Fp o = zero(); o.v[0] = 1; return mul(o); }
  Fp to_mont() const { Fp o; memcpy(o.v, C.r2, sizeof(o.v)); return mul(o); }
  Fp pow(const u64* e, int words) const {
    Fp r = one(), b = *this;
    for (int w = 0; w < words; w++)
      for (int k = 0; k < 64; k++) { if ((e[w] >> k) & 1) r = r.mul(b); b = b.sqr(); }
    return r;
  }
  Fp inv() const {
    u64 e[N]; memcpy(e, C.p, sizeof(e));
    u128 br = 2;  // p - 2
    for (int i = 0; i < N && br; i++) { u128 d = (u128)e[i] - br; e[i] = (u64)d; br = (d >> 64) & 1; }
    return pow(e, N);
  }
  Fp mul_small(int k) const {  // k may be negative
    Fp acc = zero(), cur = *this; int a = k < 0 ? -k : k;
    while (a) { if (a & 1) acc = acc.add(cur); cur = cur.dbl(); a >>= 1; }
    return k < 0 ? acc.neg() : acc;
  }
};
template <int N, int TAG> Ctx<N> Fp<N, TAG>::C;

// quadratic extension u^2 = BETA (small signed int set at run time)
template <class B>
struct Fp2 {
  static int BETA;
  static const int DEG = 2;
  B a0, a1;
  static Fp2 zero() { return Fp2{B::zero(), B::zero()}; }
  static Fp2 one() { return Fp2{B::one(), B::zero()}; }
  bool is_zero() const { return a0.is_zero() && a1.is_zero(); }
  bool eq(const Fp2& o) const { return a0.eq(o.a0) && a1.eq(o.a1); }
  Fp2 add(const Fp2& o) const { return Fp2{a0.add(o.a0), a1.add(o.a1)}; }
  Fp2 sub(const Fp2& o) const { return Fp2{a0.sub(o.a0), a1.sub(o.a1)}; }
  Fp2 neg() const { return Fp2{a0.neg(), a1.neg()}; }
  Fp2 dbl() const { return Fp2{a0.dbl(), a1.dbl()}; }
  Fp2 mul(const Fp2& o) const {
    B x = a0.mul(o.a0), y = a1.mul(o.a1);
    B z = a0.add(a1).mul(o.a0.add(o.a1));
    return Fp2{x.add(y.mul_small(BETA)), z.sub(x).sub(y)};
  }
  Fp2 sqr() const { return mul(*this); }
  Fp2 inv() const {
    B n = a0.sqr().sub(a1.sqr().mul_small(BETA));
    B ni = n.inv();
    return Fp2{a0.mul(ni), a1.mul(ni).neg()};
  }
};
template <class B> int Fp2<B>::BETA = -1;

// ---- group law (a = 0), extended Jacobian -----------------------------------
template <class K> struct Aff { K x, y; bool is_inf() const { return x.is_zero() && y.is_zero(); } };
template <class K> struct Jac { K x, y, z; };
template <class K>
struct Ext {
  K x, y, zz, zzz;
  static Ext inf() { return Ext{K::one(), K::one(), K::zero(), K::zero()}; }
  bool is_inf() const { return zz.is_zero(); }
  void dbl_self() {
    if (is_inf()) return;
    if (y.is_zero()) { *this = inf(); return; }
    K U = y.dbl(), V = U.sqr(), W = U.mul(V), S = x.mul(V), xx = x.sqr();
    K M = xx.dbl().add(xx);
    K X3 = M.sqr().sub(S.dbl());
    K Y3 = M.mul(S.sub(X3)).sub(W.mul(y));
    x = X3; y = Y3; zz = V.mul(zz); zzz = W.mul(zzz);
  }
  void set_dbl_aff(const Aff<K>& a) {
    if (a.y.is_zero()) { *this = inf(); return; }
    K U = a.y.dbl(), V = U.sqr(), W = U.mul(V), S = a.x.mul(V), xx = a.x.sqr();
    K M = xx.dbl().add(xx);
    x = M.sqr().sub(S.dbl());
    y = M.mul(S.sub(x)).sub(W.mul(a.y));
    zz = V; zzz = W;
  }
  void add_aff(const Aff<K>& a, bool negate = false) {
    if (a.is_inf()) return;
    K ay = negate ? a.y.neg() : a.y;
    if (is_inf()) { x = a.x; y = ay; zz = K::one(); zzz = K::one(); return; }
    K P = a.x.mul(zz).sub(x), R = ay.mul(zzz).sub(y);
    if (P.is_zero()) {
      if (R.is_zero()) { Aff<K> t{a.x, ay}; set_dbl_aff(t); } else *this = inf();
      return;
    }
    K PP = P.sqr(), PPP = P.mul(PP), Q = x.mul(PP);
    K X3 = R.sqr().sub(PPP).sub(Q.dbl());
    y = R.mul(Q.sub(X3)).sub(y.mul(PPP));
    x = X3; zz = zz.mul(PP); zzz = zzz.mul(PPP);
  }
  void add_ext(const Ext& q) {
    if (q.is_inf()) return;
    if (is_inf()) { *this = q; return; }
    K U1 = x.mul(q.zz), U2 = q.x.mul(zz), S1 = y.mul(q.zzz), S2 = q.y.mul(zzz);
    K P = U2.sub(U1), R = S2.sub(S1);
    if (P.is_zero()) { if (R.is_zero()) dbl_self(); else *this = inf(); return; }
    K PP = P.sqr(), PPP = P.mul(PP), Q = U1.mul(PP);
    K X3 = R.sqr().sub(PPP).sub(Q.dbl());
    y = R.mul(Q.sub(X3)).sub(S1.mul(PPP));
    x = X3; zz = zz.mul(q.zz).mul(PP); zzz = zzz.mul(q.zzz).mul(PPP);
  }
  Jac<K> to_jac() const {
    if (is_inf()) return Jac<K>{K::one(), K::one(), K::zero()};
    return Jac<K>{x.mul(zz.sqr()), y.mul(zzz.sqr()), zzz};
  }
  Aff<K> to_aff() const {
    if (is_inf()) return Aff<K>{K::zero(), K::zero()};
    K zi = zzz.inv();
    K zi2 = zi.mul(zz).sqr();
    return Aff<K>{x.mul(zi2), y.mul(zi)};
  }
};

void run_threads(int nt, const std::function<void(int)>& f) {
  if (nt <= 1) { f(0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++) th.emplace_back(f, t);
  for (auto& t : th) t.join();
}

// ---- Pippenger ----------------------------------------------------------------
// ---- batch-affine bucket accumulation ------------------------------------------------
// gnark-crypto's MultiExp switches from extended-Jacobian buckets to AFFINE buckets with batched
// additions for large windows (Montgomery's simultaneous-inversion trick: one field inversion per batch,
// 3 multiplications per element; an affine addition then costs 2M + 1S instead of 8M + 2S).  Additions in
// one batch must touch distinct buckets; a point whose bucket is already in the batch waits in a queue.
template <class K>
struct BatchAffine {
  static const int MAX_BATCH = 512;
  std::vector<Aff<K>>& bk;          // affine buckets, (0,0) = empty
  std::vector<uint8_t> busy;        // bucket is the target of an operation in the current batch
  int batch;                        // operations per inversion (scaled to the bucket count, as gnark-crypto does)
  uint32_t idx[MAX_BATCH];
  Aff<K> rhs[MAX_BATCH];
  K num[MAX_BATCH], den[MAX_BATCH]; // slope = num / den
  K pre[MAX_BATCH];
  int cnt = 0;
  struct Pending { uint32_t b; Aff<K> p; };
  std::vector<Pending> queue, tmp;

  explicit BatchAffine(std::vector<Aff<K>>& buckets) : bk(buckets), busy(buckets.size(), 0) {
    size_t b = buckets.size() / 16;
    batch = (int)(b < 32 ? 32 : (b > (size_t)MAX_BATCH ? (size_t)MAX_BATCH : b));
  }

  // stage one addition; false when the bucket is already a target in this batch or the batch is full
  bool push(uint32_t b, const Aff<K>& p) {
    if (busy[b]) return false;
    Aff<K>& B = bk[b];
    if (B.is_inf()) { B = p; return true; }
    if (cnt == batch) return false;
    if (B.x.eq(p.x)) {
      if (!B.y.eq(p.y) || p.y.is_zero()) { B = Aff<K>{K::zero(), K::zero()}; return true; }   // P + (-P)
      K xx = p.x.sqr();
      num[cnt] = xx.dbl().add(xx);                                                // doubling: 3x^2 / 2y
      den[cnt] = p.y.dbl();
    } else {
      num[cnt] = p.y.sub(B.y);
      den[cnt] = p.x.sub(B.x);
    }
    idx[cnt] = b; rhs[cnt] = p; busy[b] = 1;
    cnt++;
    return true;
  }
  void flush() {
    if (!cnt) return;
    K acc = K::one();
    for (int i = 0; i < cnt; i++) { pre[i] = acc; acc = acc.mul(den[i]); }
    K inv = acc.inv();
    for (int i = cnt - 1; i >= 0; i--) {
      K dinv = inv.mul(pre[i]);
      inv = inv.mul(den[i]);
      Aff<K>& B = bk[idx[i]];
      K lam = num[i].mul(dinv);
      K x3 = lam.sqr().sub(B.x).sub(rhs[i].x);
      K y3 = lam.mul(B.x.sub(x3)).sub(B.y);
      B.x = x3; B.y = y3;
      busy[idx[i]] = 0;
    }
    cnt = 0;
  }
  // execute the staged batch, then stage whatever queued points have become possible
  void flush_and_requeue() {
    flush();
    tmp.swap(queue);
    queue.clear();
    for (auto& q : tmp) if (!push(q.b, q.p)) queue.push_back(q);
    tmp.clear();
  }
  void add(uint32_t b, const Aff<K>& p) {
    if (p.is_inf()) return;
    if (!push(b, p)) queue.push_back(Pending{b, p});
    if (cnt == batch || queue.size() >= (size_t)(2 * batch)) flush_and_requeue();
  }
  void finish() {
    while (cnt || !queue.empty()) flush_and_requeue();
  }
};

template <class K, class S>
Ext<K> msm_pippenger(const Aff<K>* pts, const S* sc_mont, size_t n, int c, int nthreads, int scalar_bits,
                     bool batch_affine = false) {
  if (n == 0) return Ext<K>::inf();
  const int NS = sizeof(S) / 8;
  const int nwin = scalar_bits / c + 1;
  // signed digits
  std::vector<int32_t> digits((size_t)nwin * n);
  run_threads(nthreads, [&](int t) {
    for (size_t i = n * t / nthreads, end = n * (t + 1) / nthreads; i < end; i++) {   // contiguous: no false sharing
      S s = sc_mont[i].from_mont();
      int carry = 0;
      for (int w = 0; w < nwin; w++) {
        int bit = w * c, limb = bit >> 6, sh = bit & 63;
        u64 raw = 0;
        if (limb < NS) { raw = s.v[limb] >> sh; if (sh + c > 64 && limb + 1 < NS) raw |= s.v[limb + 1] << (64 - sh); }
        raw &= ((u64)1 << c) - 1;
        int64_t d = (int64_t)raw + carry;
        carry = 0;
        if (d >= ((int64_t)1 << (c - 1)) && w != nwin - 1) { d -= (int64_t)1 << c; carry = 1; }
        digits[(size_t)w * n + i] = (int32_t)d;
      }
    }
  });
  // tasks: (window, point range)
  int split = 1;
  while (nwin * split < nthreads) split++;
  const int ntasks = nwin * split;
  std::vector<Ext<K>> partial(ntasks);
  std::vector<int> next_task(1, 0);
  std::atomic<int> counter(0);
  const size_t nb = (size_t)1 << (c - 1);
  // the top window holds only scalar_bits - (nwin-1)*c bits: with a handful of live buckets a batch cannot
  // fill, so (as gnark-crypto does for its last chunk) it keeps extended-Jacobian buckets
  const int top_bits = scalar_bits - (nwin - 1) * c;
  run_threads(nthreads, [&](int) {
    std::vector<Ext<K>> buckets;
    std::vector<Aff<K>> abuckets;
    for (;;) {
      int task = counter.fetch_add(1);
      if (task >= ntasks) break;
      int w = task / split, part = task % split;
      size_t lo = n * part / split, hi = n * (part + 1) / split;
      const int32_t* dg = &digits[(size_t)w * n];
      Ext<K> run = Ext<K>::inf(), acc = Ext<K>::inf();
      if (batch_affine && !(w == nwin - 1 && top_bits < 10)) {
        abuckets.assign(nb, Aff<K>{K::zero(), K::zero()});
        BatchAffine<K> ba(abuckets);
        for (size_t i = lo; i < hi; i++) {
          int32_t d = dg[i];
          if (d > 0) ba.add((uint32_t)(d - 1), pts[i]);
          else if (d < 0) ba.add((uint32_t)(-d - 1), Aff<K>{pts[i].x, pts[i].y.neg()});
        }
        ba.finish();
        for (size_t k = nb; k-- > 0;) { run.add_aff(abuckets[k]); acc.add_ext(run); }
      } else {
        buckets.assign(nb, Ext<K>::inf());
        for (size_t i = lo; i < hi; i++) {
          int32_t d = dg[i];
          if (d > 0) buckets[d - 1].add_aff(pts[i], false);
          else if (d < 0) buckets[-d - 1].add_aff(pts[i], true);
        }
        for (size_t k = nb; k-- > 0;) { run.add_ext(buckets[k]); acc.add_ext(run); }
      }
      partial[task] = acc;
    }
  });
  Ext<K> res = Ext<K>::inf();
  for (int w = nwin - 1; w >= 0; w--) {
    for (int k = 0; k < c; k++) res.dbl_self();
    for (int part = 0; part < split; part++) res.add_ext(partial[w * split + part]);
  }
  return res;
}

template <class K, class S>
Ext<K> scalar_mul(const Aff<K>& p, const S& k_mont) {
  S k = k_mont.from_mont();
  const int NS = sizeof(S) / 8;
  Ext<K> acc = Ext<K>::inf();
  for (int i = NS - 1; i >= 0; i--)
    for (int b = 63; b >= 0; b--) { acc.dbl_self(); if ((k.v[i] >> b) & 1) acc.add_aff(p); }
  return acc;
}

// ---- fixed-base batch ------------------------------------------------------------
template <class K, class S>
void fixed_base(const Aff<K>& base, const S* sc_mont, size_t n, Aff<K>* out, int nthreads, int scalar_bits) {
  const int c = 8;
  const int nwin = (scalar_bits + c - 1) / c;
  // table[w][d] = d * 2^(c*w) * base, d in 1..255 (affine)
  std::vector<Aff<K>> table((size_t)nwin * 255);
  Ext<K> wbase = Ext<K>::inf(); wbase.add_aff(base);
  for (int w = 0; w < nwin; w++) {
    Aff<K> wb = wbase.to_aff();
    Ext<K> acc = Ext<K>::inf();
    for (int d = 1; d <= 255; d++) { acc.add_aff(wb); table[(size_t)w * 255 + d - 1] = acc.to_aff(); }
    for (int k = 0; k < c; k++) wbase.dbl_self();
  }
  const int NS = sizeof(S) / 8;
  run_threads(nthreads, [&](int t) {
    size_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
    const size_t B = 256;
    std::vector<Ext<K>> acc(B);
    std::vector<K> pre(B);
    for (size_t s0 = lo; s0 < hi; s0 += B) {
      size_t cnt = std::min(B, hi - s0);
      for (size_t j = 0; j < cnt; j++) {
        S k = sc_mont[s0 + j].from_mont();
        Ext<K> a = Ext<K>::inf();
        for (int w = 0; w < nwin; w++) {
          int bit = w * c, limb = bit >> 6, sh = bit & 63;
          if (limb >= NS) break;
          unsigned d = (unsigned)((k.v[limb] >> sh) & 255);
          if (d) a.add_aff(table[(size_t)w * 255 + d - 1]);
        }
        acc[j] = a;
      }
      // batch inversion of zzz
      K run = K::one();
      for (size_t j = 0; j < cnt; j++) { pre[j] = run; if (!acc[j].is_inf()) run = run.mul(acc[j].zzz); }
      K inv = run.inv();
      for (size_t j = cnt; j-- > 0;) {
        if (acc[j].is_inf()) { out[s0 + j] = Aff<K>{K::zero(), K::zero()}; continue; }
        K zi = inv.mul(pre[j]);
        inv = inv.mul(acc[j].zzz);
        K zi2 = zi.mul(acc[j].zz).sqr();
        out[s0 + j] = Aff<K>{acc[j].x.mul(zi2), acc[j].y.mul(zi)};
      }
    }
  });
}

// ---- NTT (recursive radix-2, gnark conventions) -------------------------------------
template <class S>
void dif_rec(S* a, size_t n, const std::vector<std::vector<S>>& tw, int stage) {
  if (n == 1) return;
  size_t m = n >> 1;
  for (size_t i = 0; i < m; i++) {
    S x = a[i], y = a[i + m];
    a[i] = x.add(y);
    a[i + m] = x.sub(y).mul(tw[stage][i]);
  }
  dif_rec(a, m, tw, stage + 1);
  dif_rec(a + m, m, tw, stage + 1);
}
template <class S>
void dit_rec(S* a, size_t n, const std::vector<std::vector<S>>& tw, int stage) {
  if (n == 1) return;
  size_t m = n >> 1;
  dit_rec(a, m, tw, stage + 1);
  dit_rec(a + m, m, tw, stage + 1);
  for (size_t i = 0; i < m; i++) {
    S x = a[i], y = a[i + m].mul(tw[stage][i]);
    a[i] = x.add(y);
    a[i + m] = x.sub(y);
  }
}
inline size_t bitrev(size_t i, int logn) { size_t r = 0; for (int k = 0; k < logn; k++) { r = (r << 1) | (i & 1); i >>= 1; } return r; }

template <class S>
void ntt(S* a, int logn, bool inverse, int decimation, bool on_coset, const S& gen, const S& coset, int nthreads) {
  const size_t n = (size_t)1 << logn;
  S w = inverse ? gen.inv() : gen;
  // twiddles[stage][i] = w^(i * 2^stage), i < n / 2^(stage+1)
  std::vector<std::vector<S>> tw(logn);
  S ws = w;
  for (int s = 0; s < logn; s++) {
    size_t m = n >> (s + 1);
    tw[s].resize(m);
    tw[s][0] = S::one();
    for (size_t i = 1; i < m; i++) tw[s][i] = tw[s][i - 1].mul(ws);
    ws = ws.sqr();
  }
  auto scale = [&](const S& g, const S& c0, bool br) {
    // a[i] *= c0 * g^j, j = br ? bitrev(i) : i
    std::vector<S> pw(n);
    pw[0] = c0;
    for (size_t j = 1; j < n; j++) pw[j] = pw[j - 1].mul(g);
    run_threads(nthreads, [&](int t) {
      for (size_t i = t; i < n; i += nthreads) a[i] = a[i].mul(pw[br ? bitrev(i, logn) : i]);
    });
  };
  if (!inverse && on_coset) scale(coset, S::one(), decimation == 1);
  // top levels of the recursion are split over threads (maxSplits in lagrange.go:132-169)
  int split_log = 0;
  while ((1 << split_log) < nthreads && split_log < logn) split_log++;
  if (decimation == 0) {
    for (int s = 0; s < split_log; s++) {
      size_t len = n >> s, m = len >> 1;
      run_threads(nthreads, [&](int t) {
        for (size_t blk = 0; blk < ((size_t)1 << s); blk++) {
          S* b = a + blk * len;
          for (size_t i = t; i < m; i += nthreads) { S x = b[i], y = b[i + m]; b[i] = x.add(y); b[i + m] = x.sub(y).mul(tw[s][i]); }
        }
      });
    }
    size_t len = n >> split_log;
    std::atomic<size_t> ctr(0);
    run_threads(nthreads, [&](int) { for (;;) { size_t blk = ctr.fetch_add(1); if (blk >= ((size_t)1 << split_log)) break; dif_rec(a + blk * len, len, tw, split_log); } });
  } else {
    size_t len = n >> split_log;
    std::atomic<size_t> ctr(0);
    run_threads(nthreads, [&](int) { for (;;) { size_t blk = ctr.fetch_add(1); if (blk >= ((size_t)1 << split_log)) break; dit_rec(a + blk * len, len, tw, split_log); } });
    for (int s = split_log - 1; s >= 0; s--) {
      size_t ln = n >> s, m = ln >> 1;
      run_threads(nthreads, [&](int t) {
        for (size_t blk = 0; blk < ((size_t)1 << s); blk++) {
          S* b = a + blk * ln;
          for (size_t i = t; i < m; i += nthreads) { S x = b[i], y = b[i + m].mul(tw[s][i]); b[i] = x.add(y); b[i + m] = x.sub(y); }
        }
      });
    }
  }
  if (inverse) {
    S nn = S::one();
    for (int k = 0; k < logn; k++) nn = nn.dbl();
    S ninv = nn.inv();
    if (on_coset) scale(coset.inv(), ninv, decimation == 0);
    else run_threads(nthreads, [&](int t) { for (size_t i = t; i < n; i += nthreads) a[i] = a[i].mul(ninv); });
  }
}

template <class S>
void compute_h(S* a, S* b, S* c, int logn, const S& gen, const S& coset, int nthreads) {
  const size_t n = (size_t)1 << logn;
  S* v[3] = {a, b, c};
  for (int k = 0; k < 3; k++) { ntt(v[k], logn, true, 0, false, gen, coset, nthreads); ntt(v[k], logn, false, 1, true, gen, coset, nthreads); }
  S gn = coset;
  for (int k = 0; k < logn; k++) gn = gn.sqr();
  S den = gn.sub(S::one()).inv();
  run_threads(nthreads, [&](int t) { for (size_t i = t; i < n; i += nthreads) a[i] = a[i].mul(b[i]).sub(c[i]).mul(den); });
  ntt(a, logn, true, 0, true, gen, coset, nthreads);
}

// ---- dispatch ------------------------------------------------------------------
template <int NP, int NR>
struct Curve {
  typedef Fp<NP, 0> Fq;
  typedef Fp<NR, 1> Fr;
  typedef Fp2<Fq> Fq2;
};

template <class K, class S>
int do_msm(const void* pts, const void* sc, size_t n, int c, int nthreads, int scalar_bits, void* out_jac,
           bool batch_affine = false) {
  Ext<K> r = msm_pippenger<K, S>((const Aff<K>*)pts, (const S*)sc, n, c, nthreads, scalar_bits, batch_affine);
  *(Jac<K>*)out_jac = r.to_jac();
  return 0;
}
template <class K, class S>
int do_msm_naive(const void* pts, const void* sc, size_t n, void* out_jac) {
  Ext<K> acc = Ext<K>::inf();
  for (size_t i = 0; i < n; i++) acc.add_ext(scalar_mul<K, S>(((const Aff<K>*)pts)[i], ((const S*)sc)[i]));
  *(Jac<K>*)out_jac = acc.to_jac();
  return 0;
}
template <class K, class S>
int do_fixed(const void* base, const void* sc, size_t n, void* out, int nthreads, int scalar_bits) {
  fixed_base<K, S>(*(const Aff<K>*)base, (const S*)sc, n, (Aff<K>*)out, nthreads, scalar_bits);
  return 0;
}

int bitlen(const u64* p, int n) { for (int i = n - 1; i >= 0; i--) if (p[i]) return 64 * i + 64 - __builtin_clzll(p[i]); return 0; }

}  // namespace

#define DISPATCH(NP, NR, EXPR_G1, EXPR_G2)                                                   \
  if (np == NP && nr == NR) {                                                                 \
    typedef Curve<NP, NR> Cv;                                                                 \
    ctx_init<NP>(Cv::Fq::C, p_mod);                                                           \
    ctx_init<NR>(Cv::Fr::C, r_mod);                                                           \
    Cv::Fq2::BETA = beta;                                                                     \
    if (deg == 1) { typedef Cv::Fq K; typedef Cv::Fr S; (void)sizeof(K); (void)sizeof(S); return EXPR_G1; } \
    else { typedef Cv::Fq2 K; typedef Cv::Fr S; (void)sizeof(K); (void)sizeof(S); return EXPR_G2; }         \
  }

// ---- Fr vector helpers for full-size fixtures and checks (oracle/plonk_fast.py): all Montgomery in / out ----------
template <class S>
static int fr_vec(int op, const S* a, const S* b, S* out, size_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  std::vector<std::thread> th;
  const size_t per = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    const size_t lo = (size_t)t * per, hi = std::min(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back([=] {
      for (size_t i = lo; i < hi; i++) out[i] = op == 0 ? a[i].mul(b[i]) : (op == 1 ? a[i].add(b[i]) : a[i].sub(b[i]));
    });
  }
  for (auto& x : th) x.join();
  return 0;
}
// out[i] = start * ratio^i
template <class S>
static int fr_geometric(const S& start, const S& ratio, size_t n, S* out) {
  S acc = start;
  for (size_t i = 0; i < n; i++) { out[i] = acc; acc = acc.mul(ratio); }
  return 0;
}
// out[i] = L_i(x) = (x^n - 1) / n * w^i / (x - w^i), n = 2^logn, w = gen (x not in the domain); one inversion (Montgomery's trick)
template <class S>
static int fr_lagrange_at(int logn, const S& gen, const S& x, S* out) {
  const size_t n = (size_t)1 << logn;
  std::vector<S> wp(n), pre(n);
  S acc = S::one();
  for (size_t i = 0; i < n; i++) { wp[i] = acc; acc = acc.mul(gen); }
  S prod = S::one();
  for (size_t i = 0; i < n; i++) {
    const S d = x.sub(wp[i]);
    if (d.is_zero()) return -2;
    pre[i] = prod;
    prod = prod.mul(d);
  }
  S inv = prod.inv();
  S xn = x;
  for (int k = 0; k < logn; k++) xn = xn.sqr();
  S nn = S::one();
  for (int k = 0; k < logn; k++) nn = nn.dbl();
  const S scale = xn.sub(S::one()).mul(nn.inv());
  for (size_t i = n; i-- > 0;) {
    const S d = x.sub(wp[i]);
    const S di = inv.mul(pre[i]);      // 1 / (x - w^i)
    inv = inv.mul(d);
    out[i] = scale.mul(wp[i]).mul(di);
  }
  return 0;
}

extern "C" {

// np/nr: 64-bit limbs of Fp/Fr; deg: 1 (coords in Fp) or 2 (Fp2 with u^2 = beta)
int orc_msm(const u64* p_mod, int np, const u64* r_mod, int nr, int deg, int beta, const void* pts, const void* sc,
            size_t n, int c, int nthreads, void* out_jac) {
  int sb = bitlen(r_mod, nr);
  DISPATCH(4, 4, (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac)), (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac)))
  DISPATCH(6, 4, (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac)), (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac)))
  DISPATCH(12, 6, (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac)), (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac)))
  return -1;
}
// same result, buckets kept affine with batched additions (gnark-crypto's choice for large windows)
int orc_msm_batch_affine(const u64* p_mod, int np, const u64* r_mod, int nr, int deg, int beta, const void* pts,
                         const void* sc, size_t n, int c, int nthreads, void* out_jac) {
  int sb = bitlen(r_mod, nr);
  DISPATCH(4, 4, (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac, true)), (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac, true)))
  DISPATCH(6, 4, (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac, true)), (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac, true)))
  DISPATCH(12, 6, (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac, true)), (do_msm<K, S>(pts, sc, n, c, nthreads, sb, out_jac, true)))
  return -1;
}
int orc_msm_naive(const u64* p_mod, int np, const u64* r_mod, int nr, int deg, int beta, const void* pts,
                  const void* sc, size_t n, void* out_jac) {
  DISPATCH(4, 4, (do_msm_naive<K, S>(pts, sc, n, out_jac)), (do_msm_naive<K, S>(pts, sc, n, out_jac)))
  DISPATCH(6, 4, (do_msm_naive<K, S>(pts, sc, n, out_jac)), (do_msm_naive<K, S>(pts, sc, n, out_jac)))
  DISPATCH(12, 6, (do_msm_naive<K, S>(pts, sc, n, out_jac)), (do_msm_naive<K, S>(pts, sc, n, out_jac)))
  return -1;
}
int orc_fixed_base(const u64* p_mod, int np, const u64* r_mod, int nr, int deg, int beta, const void* base,
                   const void* sc, size_t n, void* out_aff, int nthreads) {
  int sb = bitlen(r_mod, nr);
  DISPATCH(4, 4, (do_fixed<K, S>(base, sc, n, out_aff, nthreads, sb)), (do_fixed<K, S>(base, sc, n, out_aff, nthreads, sb)))
  DISPATCH(6, 4, (do_fixed<K, S>(base, sc, n, out_aff, nthreads, sb)), (do_fixed<K, S>(base, sc, n, out_aff, nthreads, sb)))
  DISPATCH(12, 6, (do_fixed<K, S>(base, sc, n, out_aff, nthreads, sb)), (do_fixed<K, S>(base, sc, n, out_aff, nthreads, sb)))
  return -1;
}

#define DISPATCH_FR(NR, EXPR)                        \
  if (nr == NR) {                                     \
    typedef Fp<NR, 1> S;                              \
    ctx_init<NR>(S::C, r_mod);                        \
    return EXPR;                                      \
  }
int orc_ntt(const u64* r_mod, int nr, void* data, int logn, int inverse, int decimation, int on_coset,
            const void* gen_mont, const void* coset_mont, int nthreads) {
  DISPATCH_FR(4, (ntt<S>((S*)data, logn, inverse != 0, decimation, on_coset != 0, *(const S*)gen_mont, *(const S*)coset_mont, nthreads), 0))
  DISPATCH_FR(6, (ntt<S>((S*)data, logn, inverse != 0, decimation, on_coset != 0, *(const S*)gen_mont, *(const S*)coset_mont, nthreads), 0))
  return -1;
}
int orc_compute_h(const u64* r_mod, int nr, void* a, void* b, void* c, int logn, const void* gen_mont,
                  const void* coset_mont, int nthreads) {
  DISPATCH_FR(4, (compute_h<S>((S*)a, (S*)b, (S*)c, logn, *(const S*)gen_mont, *(const S*)coset_mont, nthreads), 0))
  DISPATCH_FR(6, (compute_h<S>((S*)a, (S*)b, (S*)c, logn, *(const S*)gen_mont, *(const S*)coset_mont, nthreads), 0))
  return -1;
}
int orc_fr_vec(const u64* r_mod, int nr, int op, const void* a, const void* b, void* out, size_t n, int nthreads) {
  DISPATCH_FR(4, (fr_vec<S>(op, (const S*)a, (const S*)b, (S*)out, n, nthreads)))
  DISPATCH_FR(6, (fr_vec<S>(op, (const S*)a, (const S*)b, (S*)out, n, nthreads)))
  return -1;
}
int orc_fr_geometric(const u64* r_mod, int nr, const void* start_mont, const void* ratio_mont, size_t n, void* out) {
  DISPATCH_FR(4, (fr_geometric<S>(*(const S*)start_mont, *(const S*)ratio_mont, n, (S*)out)))
  DISPATCH_FR(6, (fr_geometric<S>(*(const S*)start_mont, *(const S*)ratio_mont, n, (S*)out)))
  return -1;
}
int orc_fr_lagrange_at(const u64* r_mod, int nr, int logn, const void* gen_mont, const void* x_mont, void* out) {
  DISPATCH_FR(4, (fr_lagrange_at<S>(logn, *(const S*)gen_mont, *(const S*)x_mont, (S*)out)))
  DISPATCH_FR(6, (fr_lagrange_at<S>(logn, *(const S*)gen_mont, *(const S*)x_mont, (S*)out)))
  return -1;
}

// out = sum_i a[i]*b[i] mod r  (Montgomery in, Montgomery out) - known-dlog bookkeeping
int orc_fr_dot(const u64* r_mod, int nr, const void* a, const void* b, size_t n, void* out) {
  DISPATCH_FR(4, ([&] { S acc = S::zero(); for (size_t i = 0; i < n; i++) acc = acc.add(((const S*)a)[i].mul(((const S*)b)[i])); *(S*)out = acc; return 0; }()))
  DISPATCH_FR(6, ([&] { S acc = S::zero(); for (size_t i = 0; i < n; i++) acc = acc.add(((const S*)a)[i].mul(((const S*)b)[i])); *(S*)out = acc; return 0; }()))
  return -1;
}
}
