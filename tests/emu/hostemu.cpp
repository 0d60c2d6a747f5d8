// CPU emulation harness for the arithmetic / MSM / NTT templates (TEST ONLY).
// Compiles the SAME field.cuh / curve.cuh / msm.cuh / ntt.cuh per-thread logic for
// the host, with the PTX carry-chain primitives emulated (ptx.cuh), so the
// "not gpu" test-suite can check the device algorithms against the oracle on a
// box without a GPU.  This library is NOT loaded by the product (gnark_b200/lib.py
// loads libgnark_b200.so only and fails loudly without CUDA).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>


#include <fcntl.h>

#include "fixed_base.cuh"
#include "file_stage.h"
#include "host_fr.h"
#include "msm.cuh"
#include "ntt.cuh"
#include "plonk.cuh"
#include "points_decode.cuh"
#include "emu_plonk.h"

using namespace gb200;

namespace {

template <class F>
int field_op(int op, const void* a_, const void* b_, void* out_) {
  const F& a = *reinterpret_cast<const F*>(a_);
  F r;
  switch (op) {
    case 0: r = a + *reinterpret_cast<const F*>(b_); break;
    case 1: r = a - *reinterpret_cast<const F*>(b_); break;
    case 2: r = a * *reinterpret_cast<const F*>(b_); break;
    case 3: r = a.inverse(); break;
    case 4: r = a.neg(); break;
    case 5: r = a.sqr(); break;
    case 6: r = a.dbl(); break;
    default: return -1;
  }
  *reinterpret_cast<F*>(out_) = r;
  return 0;
}

// raw wide routines (field.cuh): op 0 wide_mul (a, b: N limbs -> 2N), 2 mont_reduce_wide (a: 2N -> N)
template <class F>
int wide_op(int op, const void* a_, const void* b_, void* out_) {
  constexpr int N = F::N;
  const uint32_t* a = reinterpret_cast<const uint32_t*>(a_);
  const uint32_t* b = reinterpret_cast<const uint32_t*>(b_);
  uint32_t* out = reinterpret_cast<uint32_t*>(out_);
  switch (op) {
    case 0: wide_mul_raw<N>(out, a, b); return 0;
    case 2: mont_reduce_wide<typename F::Params>(out, a); return 0;
  }
  return -1;
}

template <class Fr, class F>
int msm_emu(const void* points_, const void* scalars_, uint32_t n, int c, int precomp, uint32_t task_len,
            uint32_t chunk, void* out_jac) {
  const Affine<F>* points = reinterpret_cast<const Affine<F>*>(points_);
  const Fr* scalars = reinterpret_cast<const Fr*>(scalars_);
  MsmPlan pl = msm_make_plan(n, n, 0, Fr::Params::BITS, c, precomp, task_len, chunk);
  // table
  std::vector<Affine<F>> table(points, points + n);
  if (precomp) {
    table.resize((size_t)n * pl.nwin);
    for (int w = 1; w < pl.nwin; w++)
      for (uint32_t i = 0; i < n; i++) table[(size_t)w * n + i] = msm_shift_point(table[(size_t)(w - 1) * n + i], c);
  }
  const size_t m = (size_t)n * pl.nwin;
  std::vector<uint32_t> keys(m), vals(m);
  for (uint32_t i = 0; i < n; i++) msm_decompose_one<Fr>(pl, i, scalars, keys.data(), vals.data());
  std::vector<uint32_t> order(m);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return keys[x] < keys[y]; });
  std::vector<uint32_t> skeys(m), svals(m);
  for (size_t k = 0; k < m; k++) { skeys[k] = keys[order[k]]; svals[k] = vals[order[k]]; }
  // offsets
  std::vector<uint32_t> off(pl.total_buckets + 1);
  for (uint32_t b = 0; b <= pl.total_buckets; b++)
    off[b] = (uint32_t)(std::lower_bound(skeys.begin(), skeys.end(), b) - skeys.begin());
  // accumulate with tasks, then combine
  std::vector<XYZZ<F>> buckets(pl.total_buckets);
  for (uint32_t b = 0; b < pl.total_buckets; b++) {
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t s = off[b]; s < off[b + 1]; s += pl.task_len) {
      uint32_t e = std::min(off[b + 1], s + pl.task_len);
      acc.add(msm_accumulate_range<F>(table.data(), svals.data(), s, e));
    }
    buckets[b] = acc;
  }
  // reduce
  std::vector<XYZZ<F>> set_sums(pl.nsets);
  for (int s = 0; s < pl.nsets; s++) {
    XYZZ<F> tot = XYZZ<F>::inf();
    for (uint32_t lo = 0; lo < pl.set_size; lo += pl.chunk) {
      uint32_t hi = std::min(pl.set_size, lo + pl.chunk);
      tot.add(msm_reduce_chunk<F>(buckets.data() + (size_t)s * pl.set_size, lo, hi));
    }
    set_sums[s] = tot;
  }
  XYZZ<F> res = msm_horner<F>(set_sums.data(), pl.nsets, c);
  *reinterpret_cast<Jacobian<F>*>(out_jac) = res.to_jacobian();
  return 0;
}

int g_ntt_tile_log = NTT_MAX_TILE_LOG;
template <class Fr>
int ntt_emu(void* data_, unsigned logn, int inverse, int decimation, int on_coset, const void* gen_mont,
            const void* coset_mont) {
  Fr* data = reinterpret_cast<Fr*>(data_);
  NttDomainHost<Fr> dom;
  dom.tile_log = g_ntt_tile_log;
  dom.init(logn, gen_mont ? reinterpret_cast<const Fr*>(gen_mont) : nullptr,
           coset_mont ? reinterpret_cast<const Fr*>(coset_mont) : nullptr);
  dom.transform(data, inverse != 0, decimation, on_coset != 0);
  return 0;
}

}  // namespace

template <class Fr, class F>
int fixed_base_emu(const void* base_, const void* scalars_, uint32_t n, int c, void* out_) {
  const Affine<F> base = *reinterpret_cast<const Affine<F>*>(base_);
  const Fr* scalars = reinterpret_cast<const Fr*>(scalars_);
  Affine<F>* out = reinterpret_cast<Affine<F>*>(out_);
  const FixedBasePlan pl = fixed_base_plan(Fr::Params::BITS, c);
  if (pl.nwin > FB_MAX_WINDOWS) return -2;
  std::vector<Affine<F>> Pw(pl.nwin);
  for (int w = 0; w < pl.nwin; w++) Pw[w] = fixed_base_window_base<F>(base, c, w).to_affine();        // k_fb_window_bases
  const size_t entries = (size_t)pl.nwin * pl.half;
  std::vector<XYZZ<F>> tmp(std::max(entries, (size_t)n));
  for (size_t t = 0; t < entries; t++) tmp[t] = fixed_base_entry<F>(Pw[t / pl.half], (uint32_t)(t % pl.half));  // k_fb_table
  std::vector<Affine<F>> table(entries);
  for (size_t first = 0; first < entries; first += FB_CHUNK)                                            // k_fb_to_affine
    xyzz_batch_to_affine<F>(tmp.data() + first, table.data() + first, (uint32_t)std::min((size_t)FB_CHUNK, entries - first));
  for (uint32_t i = 0; i < n; i++) tmp[i] = fixed_base_eval<Fr, F>(pl, table.data(), scalars[i]);      // k_fb_eval
  for (size_t first = 0; first < n; first += FB_CHUNK)
    xyzz_batch_to_affine<F>(tmp.data() + first, out + first, (uint32_t)std::min((size_t)FB_CHUNK, (size_t)n - first));
  return 0;
}


// k_msm_precompute_dbl + k_msm_precompute_affine (msm_impl.cuh) for ONE point: the XYZZ doubling chain that is never
// normalised between slabs, then one inversion for the nwin - 1 slab points (Montgomery's trick); out: nwin affine points
template <class F>
int precompute_emu(const void* point_, int nwin, int c, void* out_) {
  const Affine<F> src = *reinterpret_cast<const Affine<F>*>(point_);
  Affine<F>* out = reinterpret_cast<Affine<F>*>(out_);
  std::vector<XYZZ<F>> tmp(nwin > 1 ? nwin - 1 : 0);
  XYZZ<F> q = XYZZ<F>::from_affine(src);
  for (int w = 1; w < nwin; w++) {
    for (int k = 0; k < c; k++) q.dbl();
    tmp[w - 1] = q;
  }
  out[0] = src;
  std::vector<F> pre(nwin + 1);
  F acc = F::one();
  for (int w = 1; w < nwin; w++) {
    pre[w] = acc;
    const F zzz = tmp[w - 1].zzz;
    if (!zzz.is_zero()) acc = acc * zzz;
  }
  F inv = acc.inverse();
  for (int w = nwin - 1; w >= 1; w--) {
    const XYZZ<F> t = tmp[w - 1];
    Affine<F> a = Affine<F>::inf();
    if (!t.zzz.is_zero()) {
      const F zi3 = inv * pre[w];
      inv = inv * t.zzz;
      const F zi2 = (zi3 * t.zz).sqr();
      a.x = t.x * zi2;
      a.y = t.y * zi3;
    }
    out[w] = a;
  }
  return 0;
}

// k_plonk_add_bsb22 walked over the n points of one coset (BN254 / BLS12-381 / BW6 Fr by curve id)
template <class Fr>
int bsb22_emu(const void* qcp, const void* pi2, void* out, uint32_t logn, uint32_t coset_index, uint32_t rho) {
  uint32_t log_rho = 0;
  while ((1u << log_rho) < rho) log_rho++;
  for (uint32_t j = 0; j < (1u << logn); j++)
    plonk_add_bsb22_point<Fr>((const Fr*)qcp, (const Fr*)pi2, (Fr*)out, j, coset_index, rho, logn, log_rho);
  return 0;
}

// points_decode.cuh: n serialised points -> affine Montgomery points; returns the first DECODE_* status that is not OK
template <class F>
static int decode_emu(const void* bytes, size_t n, int encoding, int curve, int group, void* out) {
  using FB = typename F::Base;
  const size_t stride = encoding == POINTS_RAW ? sizeof(Affine<F>) : sizeof(F);
  const DecodeConsts<FB> k = decode_make_consts<F, FB>(curve, group);
  int first = 0;
  for (size_t i = 0; i < n; i++) {
    Affine<F> a = Affine<F>::inf();
    const int rc = decode_point<F, FB>((const uint8_t*)bytes + i * stride, encoding == POINTS_COMPRESSED, k, a);
    ((Affine<F>*)out)[i] = a;
    if (rc && !first) first = rc;
  }
  return first;
}

extern "C" {

// polys: 12 pointers (l r o z s1 s2 s3 ql qr qm qo qk), each n fr.Elements ON the coset; abg: alpha, beta, gamma
int emu_plonk_constraints_coset(int curve, const void* const* polys, const void* abg, const void* const* blind,
                                const int* nblind, uint32_t logn, uint32_t coset_index, uint32_t rho, void* out) {
  switch (curve) {
    case 0: return plonk_coset_emu<bn254_fr>(polys, abg, blind, nblind, logn, coset_index, rho, out);
    case 1: return plonk_coset_emu<bls12_381_fr>(polys, abg, blind, nblind, logn, coset_index, rho, out);
    case 2: return plonk_coset_emu<bls12_377_fr>(polys, abg, blind, nblind, logn, coset_index, rho, out);
    case 3: return plonk_coset_emu<bw6_761_fr>(polys, abg, blind, nblind, logn, coset_index, rho, out);
  }
  return -1;
}






int emu_decode_points(int curve, int group, const void* bytes, size_t n, int encoding, void* out) {
  switch (curve * 2 + (group - 1)) {
    case 0: return decode_emu<bn254_fp>(bytes, n, encoding, curve, group, out);
    case 1: return decode_emu<bn254_fp2>(bytes, n, encoding, curve, group, out);
    case 2: return decode_emu<bls12_381_fp>(bytes, n, encoding, curve, group, out);
    case 3: return decode_emu<bls12_381_fp2>(bytes, n, encoding, curve, group, out);
    case 4: return decode_emu<bls12_377_fp>(bytes, n, encoding, curve, group, out);
    case 5: return decode_emu<bls12_377_fp2>(bytes, n, encoding, curve, group, out);
    case 6:
    case 7: return decode_emu<bw6_761_fp>(bytes, n, encoding, curve, group, out);
  }
  return -1;
}

int emu_precompute(int curve, int group, const void* point, int nwin, int c, void* out_affine) {
  switch (curve * 2 + (group - 1)) {
    case 0: return precompute_emu<bn254_fp>(point, nwin, c, out_affine);
    case 1: return precompute_emu<bn254_fp2>(point, nwin, c, out_affine);
    case 2: return precompute_emu<bls12_381_fp>(point, nwin, c, out_affine);
    case 3: return precompute_emu<bls12_381_fp2>(point, nwin, c, out_affine);
    case 4: return precompute_emu<bls12_377_fp>(point, nwin, c, out_affine);
    case 5: return precompute_emu<bls12_377_fp2>(point, nwin, c, out_affine);
    case 6:
    case 7: return precompute_emu<bw6_761_fp>(point, nwin, c, out_affine);
  }
  return -1;
}

int emu_fixed_base(int curve, int group, const void* base, const void* scalars, uint32_t n, int c, void* out_affine) {
  switch (curve * 2 + (group - 1)) {
    case 0: return fixed_base_emu<bn254_fr, bn254_fp>(base, scalars, n, c, out_affine);
    case 1: return fixed_base_emu<bn254_fr, bn254_fp2>(base, scalars, n, c, out_affine);
    case 2: return fixed_base_emu<bls12_381_fr, bls12_381_fp>(base, scalars, n, c, out_affine);
    case 3: return fixed_base_emu<bls12_381_fr, bls12_381_fp2>(base, scalars, n, c, out_affine);
    case 4: return fixed_base_emu<bls12_377_fr, bls12_377_fp>(base, scalars, n, c, out_affine);
    case 5: return fixed_base_emu<bls12_377_fr, bls12_377_fp2>(base, scalars, n, c, out_affine);
    case 6:
    case 7: return fixed_base_emu<bw6_761_fr, bw6_761_fp>(base, scalars, n, c, out_affine);
  }
  return -1;
}

// host_fr.h (run-time limb count, used by plonk_host.cu): op 0 add, 1 sub, 2 mul, 3 inv, 4 neg, 5 from_u64(a[0]),
// 6 domain_generator(log2n = a[0]), 7 mult_gen, 8 pow_u64(a, b[0])
int emu_hostfr_op(int curve, int op, const void* a_, const void* b_, void* out) {
  HostFrCtx c;
  switch (curve) {
    case 0: c = HostFrCtx::make<bn254_fr_params>(); break;
    case 1: c = HostFrCtx::make<bls12_381_fr_params>(); break;
    case 2: c = HostFrCtx::make<bls12_377_fr_params>(); break;
    case 3: c = HostFrCtx::make<bw6_761_fr_params>(); break;
    default: return -1;
  }
  const HostFr a = c.load(a_), b = c.load(b_);
  HostFr r;
  switch (op) {
    case 0: r = c.add(a, b); break;
    case 1: r = c.sub(a, b); break;
    case 2: r = c.mul(a, b); break;
    case 3: r = c.inv(a); break;
    case 4: r = c.neg(a); break;
    case 5: r = c.from_u64(*reinterpret_cast<const uint64_t*>(a_)); break;
    case 6: r = c.domain_generator((int)*reinterpret_cast<const uint64_t*>(a_)); break;
    case 7: r = c.mult_gen; break;
    case 8: r = c.pow_u64(a, *reinterpret_cast<const uint64_t*>(b_)); break;
    default: return -1;
  }
  c.store(out, r);
  return 0;
}

int emu_wide_op(int field_id, int op, const void* a, const void* b, void* out) {
  switch (field_id) {
    case 0: return wide_op<bn254_fp>(op, a, b, out);
    case 1: return wide_op<bn254_fr>(op, a, b, out);
    case 2: return wide_op<bls12_381_fp>(op, a, b, out);
    case 3: return wide_op<bls12_381_fr>(op, a, b, out);
    case 4: return wide_op<bls12_377_fp>(op, a, b, out);
    case 5: return wide_op<bls12_377_fr>(op, a, b, out);
    case 6: return wide_op<bw6_761_fp>(op, a, b, out);
    case 7: return wide_op<bw6_761_fr>(op, a, b, out);
  }
  return -1;
}

// field_id = curve*2 + (0: fp, 1: fr)
int emu_field_op(int field_id, int op, const void* a, const void* b, void* out) {
  switch (field_id) {
    case 0: return field_op<bn254_fp>(op, a, b, out);
    case 1: return field_op<bn254_fr>(op, a, b, out);
    case 2: return field_op<bls12_381_fp>(op, a, b, out);
    case 3: return field_op<bls12_381_fr>(op, a, b, out);
    case 4: return field_op<bls12_377_fp>(op, a, b, out);
    case 5: return field_op<bls12_377_fr>(op, a, b, out);
    case 6: return field_op<bw6_761_fp>(op, a, b, out);
    case 7: return field_op<bw6_761_fr>(op, a, b, out);
    case 100: return field_op<bn254_fp2>(op, a, b, out);
    case 102: return field_op<bls12_381_fp2>(op, a, b, out);
    case 104: return field_op<bls12_377_fp2>(op, a, b, out);
  }
  return -1;
}


int emu_msm(int curve, int group, const void* points, const void* scalars, uint32_t n, int c, int precomp,
            uint32_t task_len, uint32_t chunk, void* out_jac) {
  switch (curve * 2 + (group - 1)) {
    case 0: return msm_emu<bn254_fr, bn254_fp>(points, scalars, n, c, precomp, task_len, chunk, out_jac);
    case 1: return msm_emu<bn254_fr, bn254_fp2>(points, scalars, n, c, precomp, task_len, chunk, out_jac);
    case 2: return msm_emu<bls12_381_fr, bls12_381_fp>(points, scalars, n, c, precomp, task_len, chunk, out_jac);
    case 3: return msm_emu<bls12_381_fr, bls12_381_fp2>(points, scalars, n, c, precomp, task_len, chunk, out_jac);
    case 4: return msm_emu<bls12_377_fr, bls12_377_fp>(points, scalars, n, c, precomp, task_len, chunk, out_jac);
    case 5: return msm_emu<bls12_377_fr, bls12_377_fp2>(points, scalars, n, c, precomp, task_len, chunk, out_jac);
    case 6:
    case 7: return msm_emu<bw6_761_fr, bw6_761_fp>(points, scalars, n, c, precomp, task_len, chunk, out_jac);
  }
  return -1;
}

int emu_plonk_bsb22(int curve, const void* qcp, const void* pi2, void* out, uint32_t logn, uint32_t coset_index, uint32_t rho) {
  switch (curve) {
    case 0: return bsb22_emu<bn254_fr>(qcp, pi2, out, logn, coset_index, rho);
    case 1: return bsb22_emu<bls12_381_fr>(qcp, pi2, out, logn, coset_index, rho);
    case 2: return bsb22_emu<bls12_377_fr>(qcp, pi2, out, logn, coset_index, rho);
    case 3: return bsb22_emu<bw6_761_fr>(qcp, pi2, out, logn, coset_index, rho);
  }
  return -1;
}

// file_stage.h (host half of b200_table_upload_file): read [off, off + bytes) of `path` through two slots of slot_bytes
// into out; the slots are poisoned after each consumption so that a chunk delivered twice or a stale slot shows up.
// Returns 0, -1 (open), -2 (staging error: short file ...), -3 (a slot was overwritten before it was consumed)
int emu_stage_file(const char* path, uint64_t off, size_t bytes, size_t slot_bytes, void* out) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) return -1;
  std::vector<unsigned char> a(slot_bytes ? slot_bytes : 1), b(slot_bytes ? slot_bytes : 1);
  void* slots[2] = {a.data(), b.data()};
  bool pending[2] = {false, false};
  bool overwritten = false;
  std::string err;
  int r = stage_file_region(
      fd, off, bytes, slots, slot_bytes,
      [&](int slot, const void* data, size_t pos, size_t len) {
        memcpy((char*)out + pos, data, len);
        memset(slots[slot], 0xEE, slot_bytes);
        pending[slot] = true;
        return 0;
      },
      [&](int slot) {
        // the "copy engine" releases a slot only when asked: a slot still pending here must not have been touched
        if (pending[slot]) {
          const unsigned char* p = (const unsigned char*)slots[slot];
          for (size_t i = 0; i < slot_bytes; i++) if (p[i] != 0xEE) overwritten = true;
          pending[slot] = false;
        }
        return 0;
      },
      &err);
  close(fd);
  if (overwritten) return -3;
  return r == 0 ? 0 : -2;
}

// plan tile size used by emu_ntt (GB200_NTT_TILE_LOG on the device)
// 1: emu_ntt walks the register-round kernel (GB200_NTT_RADIX8 on the device) instead of the one-stage-per-barrier one
int emu_ntt_set_tile_log(int t) { if (t < 2 || t > NTT_MAX_TILE_LOG) return -1; g_ntt_tile_log = t; return 0; }

int emu_ntt(int curve, void* data, unsigned logn, int inverse, int decimation, int on_coset, const void* gen_mont,
            const void* coset_mont) {
  switch (curve) {
    case 0: return ntt_emu<bn254_fr>(data, logn, inverse, decimation, on_coset, gen_mont, coset_mont);
    case 1: return ntt_emu<bls12_381_fr>(data, logn, inverse, decimation, on_coset, gen_mont, coset_mont);
    case 2: return ntt_emu<bls12_377_fr>(data, logn, inverse, decimation, on_coset, gen_mont, coset_mont);
    case 3: return ntt_emu<bw6_761_fr>(data, logn, inverse, decimation, on_coset, gen_mont, coset_mont);
  }
  return -1;
}

}  // extern "C"
