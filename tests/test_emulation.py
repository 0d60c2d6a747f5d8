"""The device templates (field.cuh / curve.cuh / msm.cuh / ntt.cuh), compiled for the
host with the PTX carry chains emulated, against the oracle.  This checks the
kernels' per-thread logic on a box without a GPU; the -m gpu tests check the real
thing through the C ABI."""
import ctypes
import random

import numpy as np
import pytest

from oracle import ec, ff, ntt
from oracle.params import CURVES
from util import pick_base

ALL = list(CURVES.values())


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_field_ops(hostemu, c):
    rng = random.Random(5)
    for which, (q, L) in enumerate(((c.p, c.fp_limbs), (c.r, c.fr_limbs))):
        fid = c.curve_id * 2 + which
        for trial in range(40):
            a, b = rng.randrange(q), rng.randrange(q)
            if trial == 0: a, b = 0, 0
            if trial == 1: a, b = q - 1, q - 1
            if trial == 2: a, b = 1, q - 1
            if trial == 3: a, b = (1 << (64 * L)) % q, q - 2      # R mod q
            A, B = ff.pack_elements([a], q, L), ff.pack_elements([b], q, L)
            O = np.zeros_like(A)
            for op, exp in ((0, (a + b) % q), (1, (a - b) % q), (2, a * b % q), (4, (-a) % q), (5, a * a % q), (6, 2 * a % q)):
                assert hostemu.emu_field_op(fid, op, P(A), P(B), P(O)) == 0
                assert ff.unpack_elements(O, q, L)[0] == exp, (c.name, which, op)
            if a and trial < 6:
                hostemu.emu_field_op(fid, 3, P(A), P(B), P(O))
                assert ff.unpack_elements(O, q, L)[0] == pow(a, -1, q)
            if a:       # binary-GCD inversion (shift/add only): every trial incl. 1, q-1, R mod q, powers of two
                hostemu.emu_field_op(fid, 7, P(A), P(B), P(O))
                assert ff.unpack_elements(O, q, L)[0] == pow(a, -1, q), (c.name, which, "inverse_gcd", trial)


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_inverse_gcd_many(hostemu, c):
    """Fp::inverse_gcd (shift-and-add binary GCD, used by the batched-affine levels) against big-int inversion:
    random values, values with long runs of zero bits (whole-limb shifts), small values, p - small"""
    rng = random.Random(1234)
    for which, (q, L) in enumerate(((c.p, c.fp_limbs), (c.r, c.fr_limbs))):
        fid = c.curve_id * 2 + which
        vals = [rng.randrange(1, q) for _ in range(200)]
        vals += [1 << k for k in (1, 31, 32, 33, 63, 64, 65, 96, 127, 128, 200) if (1 << k) < q]
        vals += [(rng.randrange(1, 1 << 40) << 96) % q or 1 for _ in range(20)]
        vals += [k for k in range(1, 20)] + [q - k for k in range(1, 20)]
        R = 1 << (64 * L)
        for a in vals:
            # the stored limbs are a*R mod q; choose a so that the STORED value has the special shape too
            for stored in (a * R % q, a):
                val = stored * pow(R, -1, q) % q
                A = ff.pack_elements([val], q, L)
                O = np.zeros_like(A)
                assert hostemu.emu_field_op(fid, 7, P(A), P(A), P(O)) == 0
                assert ff.unpack_elements(O, q, L)[0] == pow(val, -1, q), (c.name, which, hex(a))


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_wide_arithmetic(hostemu, c):
    """wide_mul_raw / wide_sqr_raw / mont_reduce_wide (building blocks of the optional squaring and lazy Fp2 paths)
    against big-int arithmetic: ANY N-limb operands for the products (all-ones limbs included), T < p*R for the
    reduction (largest admissible values included)"""
    rng = random.Random(77)
    for which, (q, L) in enumerate(((c.p, c.fp_limbs), (c.r, c.fr_limbs))):
        fid = c.curve_id * 2 + which
        top = 1 << (64 * L)
        R = top
        def arr(x, limbs):
            return np.frombuffer(int(x).to_bytes(8 * limbs, "little"), dtype=np.uint64).copy()
        def val(a):
            return int.from_bytes(a.tobytes(), "little")
        cases = [(0, 0), (top - 1, top - 1), (1, top - 1), (q - 1, q - 1), (2 * q - 1, 2 * q - 1)]
        cases += [(rng.randrange(top), rng.randrange(top)) for _ in range(60)]
        for a, b in cases:
            O = np.zeros(2 * L, dtype=np.uint64)
            assert hostemu.emu_wide_op(fid, 0, P(arr(a, L)), P(arr(b, L)), P(O)) == 0
            assert val(O) == a * b, (c.name, which, "wide_mul")
            assert hostemu.emu_wide_op(fid, 1, P(arr(a, L)), P(arr(a, L)), P(O)) == 0
            assert val(O) == a * a, (c.name, which, "wide_sqr")
            assert hostemu.emu_wide_op(fid, 3, P(arr(a, L)), P(arr(b, L)), P(O)) == 0
            assert val(O) == a * b, (c.name, which, "wide_mul_karatsuba")
        Rinv = pow(R, -1, q)
        ts = [0, q * R - 1, q * q, (q - 1) * (q - 1), 6 * q * q if 6 * q < R else q * q]
        ts += [rng.randrange(q * R) for _ in range(60)]
        for t in ts:
            O = np.zeros(L, dtype=np.uint64)
            assert hostemu.emu_wide_op(fid, 2, P(arr(t, 2 * L)), P(arr(0, L)), P(O)) == 0
            assert val(O) == t * Rinv % q, (c.name, which, "mont_reduce_wide")


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
@pytest.mark.parametrize("opt", (False, True), ids=("default", "opt"))
def test_mul_sub_single_reduction(hostemu, hostemu_opt, c, opt):
    """Fp::mul_sub = a b - c d through ONE Montgomery reduction of a b + (p^2 - c d) (GB200_XYZZ_LAZY uses it for Y3 of
    the mixed addition); extremes of both products included; default wide products and the Karatsuba ones"""
    lib = hostemu_opt if opt else hostemu
    rng = random.Random(31)
    q, L = c.p, c.fp_limbs
    fid = c.curve_id * 2
    ext = [0, 1, q - 1, q - 2]
    cases = [(a, b, cc, d) for a in ext for b in (0, q - 1) for cc in (0, q - 1) for d in ext]
    cases += [tuple(rng.randrange(q) for _ in range(4)) for _ in range(60)]
    for a, b, cc, d in cases:
        A = ff.pack_elements([a], q, L)
        B = ff.pack_elements([b, cc, d], q, L)
        O = np.zeros_like(A)
        assert lib.emu_field_op(fid, 8, P(A), P(B), P(O)) == 0
        assert ff.unpack_elements(O, q, L)[0] == (a * b - cc * d) % q, (c.name, hex(a), hex(b), hex(cc), hex(d))


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_optional_paths_field_ops(hostemu_opt, c):
    """field / Fp2 operations of the library compiled with GB200_MONT_SQR + GB200_FP2_LAZY against big-int"""
    test_field_ops.__wrapped__(hostemu_opt, c) if hasattr(test_field_ops, "__wrapped__") else test_field_ops(hostemu_opt, c)
    if c.fp2_nonresidue is not None:
        test_fp2_ops(hostemu_opt, c)


@pytest.mark.parametrize("c,group", [(CURVES["bn254"], 1), (CURVES["bn254"], 2), (CURVES["bls12-377"], 1), (CURVES["bls12-377"], 2),
                                     (CURVES["bw6-761"], 1)], ids=lambda v: getattr(v, "name", str(v)))
def test_optional_paths_msm(hostemu_opt, c, group):
    """one MSM per group through the optional arithmetic paths (G2 of BLS12-377 exercises BETA = 5 in the lazy product)"""
    rng = random.Random(9 + group)
    F, base = pick_base(c, group, rng)
    n = 23
    pts = [ec.scalar_mul(F, rng.randrange(1, 1 << 40), base) for _ in range(n)]
    pts[3] = ec.INF
    pts[5] = pts[4]
    pts[7] = ec.affine_neg(F, pts[6])
    sc = [rng.randrange(c.r) for _ in range(n)]
    sc[0], sc[1], sc[2] = 0, c.r - 1, 1
    sc[4] = sc[5]
    sc[6] = sc[7] = 12345
    exp = ec.msm_naive(F, pts, sc)
    PA, SA = ec.pack_points(c, group, pts), ff.pack_elements(sc, c.r, c.fr_limbs)
    deg = 2 if (group == 2 and c.fp2_nonresidue is not None) else 1
    for (cw, pre, tl, ch) in ((7, 1, 2, 16), (5, 0, 3, 4)):
        out = np.zeros(3 * c.fp_limbs * deg, dtype=np.uint64)
        assert hostemu_opt.emu_msm(c.curve_id, group, P(PA), P(SA), n, cw, pre, tl, ch, P(out)) == 0
        assert ec.from_jac(F, ec.unpack_points(c, group, out, ncoords=3)[0]) == exp
    out = np.zeros(3 * c.fp_limbs * deg, dtype=np.uint64)
    assert hostemu_opt.emu_msm_ba(c.curve_id, group, P(PA), P(SA), n, 5, 1, 4, 8, 3, P(out)) == 0
    assert ec.from_jac(F, ec.unpack_points(c, group, out, ncoords=3)[0]) == exp


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_host_fr(hostemu, c):
    """host_fr.h: the run-time-limb-count Fr arithmetic the C++ PLONK orchestration uses between device stages"""
    rng = random.Random(99)
    q, L = c.r, c.fr_limbs
    pe = lambda v: ff.pack_elements([v], q, L)
    un = lambda A: ff.unpack_elements(A, q, L)[0]
    O = np.zeros(L, dtype=np.uint64)
    for trial in range(30):
        a, b = rng.randrange(q), rng.randrange(q)
        if trial == 0: a, b = 0, 0
        if trial == 1: a, b = q - 1, q - 1
        if trial == 2: a, b = 1, q - 1
        for op, exp in ((0, (a + b) % q), (1, (a - b) % q), (2, a * b % q), (4, (-a) % q)):
            assert hostemu.emu_hostfr_op(c.curve_id, op, P(pe(a)), P(pe(b)), P(O)) == 0
            assert un(O) == exp, (c.name, op)
        if a:
            hostemu.emu_hostfr_op(c.curve_id, 3, P(pe(a)), P(pe(b)), P(O))
            assert un(O) == pow(a, -1, q)
        e = np.array([rng.randrange(1 << 40)] + [0] * (L - 1), dtype=np.uint64)
        hostemu.emu_hostfr_op(c.curve_id, 8, P(pe(a)), P(e), P(O))
        assert un(O) == pow(a, int(e[0]), q)
        hostemu.emu_hostfr_op(c.curve_id, 5, P(e), P(e), P(O))
        assert un(O) == int(e[0]) % q
    from oracle import ntt as ontt
    for logn in (1, 4, 20):
        k = np.array([logn] + [0] * (L - 1), dtype=np.uint64)
        hostemu.emu_hostfr_op(c.curve_id, 6, P(k), P(k), P(O))
        assert un(O) == ontt.Domain(c, 1 << logn).generator
    hostemu.emu_hostfr_op(c.curve_id, 7, P(O.copy()), P(O.copy()), P(O))
    assert un(O) == c.mult_gen


@pytest.mark.parametrize("c", [c for c in ALL if c.fp2_nonresidue is not None], ids=lambda c: c.name)
def test_fp2_ops(hostemu, c):
    rng = random.Random(6)
    F2 = ff.Fp2(c.p, c.fp2_nonresidue)
    L = c.fp_limbs
    for _ in range(20):
        a = (rng.randrange(c.p), rng.randrange(c.p))
        b = (rng.randrange(c.p), rng.randrange(c.p))
        A = ff.pack_elements(list(a), c.p, L).reshape(-1)
        B = ff.pack_elements(list(b), c.p, L).reshape(-1)
        O = np.zeros_like(A)
        for op, exp in ((2, F2.mul(a, b)), (5, F2.sqr(a)), (3, F2.inv(a)), (7, F2.inv(a)), (0, F2.add(a, b)), (1, F2.sub(a, b))):
            assert hostemu.emu_field_op(100 + c.curve_id * 2, op, P(A), P(B), P(O)) == 0
            assert tuple(ff.unpack_elements(O, c.p, L)) == exp


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
@pytest.mark.parametrize("group", (1, 2))
def test_msm_logic(hostemu, c, group):
    rng = random.Random(9 + group)
    F, base = pick_base(c, group, rng)
    n = 37
    pts = [ec.scalar_mul(F, rng.randrange(1, 1 << 40), base) for _ in range(n)]
    pts[3] = ec.INF
    pts[5] = pts[4]
    pts[7] = ec.affine_neg(F, pts[6])
    sc = [rng.randrange(c.r) for _ in range(n)]
    sc[0], sc[1], sc[2] = 0, c.r - 1, 1
    sc[6] = sc[7] = 12345
    sc[8], sc[9] = 1 << 15, (1 << 16) - 1
    exp = ec.msm_naive(F, pts, sc)
    PA, SA = ec.pack_points(c, group, pts), ff.pack_elements(sc, c.r, c.fr_limbs)
    for (cw, pre, tl, ch) in ((4, 0, 3, 4), (7, 1, 2, 16), (16, 0, 64, 512), (13, 1, 5, 100)):
        if cw >= 13 and c.fp_limbs > 6:
            continue
        out = np.zeros(3 * F.degree * c.fp_limbs, dtype=np.uint64)
        assert hostemu.emu_msm(c.curve_id, group, P(PA), P(SA), n, cw, pre, tl, ch, P(out)) == 0
        got = ec.from_jac(F, ec.unpack_points(c, group, out, ncoords=3)[0])
        assert got == exp, (c.name, group, cw, pre)


@pytest.mark.parametrize("c,group", [(CURVES["bn254"], 1), (CURVES["bn254"], 2), (CURVES["bls12-381"], 1)],
                         ids=lambda v: getattr(v, "name", str(v)))
def test_msm_persistent_accumulate_logic(hostemu, c, group):
    """opt-in GB200_MSM_PERSISTENT: the accumulate stage on a fixed number of threads that take tasks from a counter
    (msm_accumulate_persistent) - every task exactly once, same result; thread counts below, equal to and above the
    number of tasks; skewed scalars (one very long bucket, many empty ones)"""
    rng = random.Random(40 + group)
    F, base = pick_base(c, group, rng)
    n = 61
    pts = [ec.scalar_mul(F, rng.randrange(1, 1 << 40), base) for _ in range(n)]
    pts[3] = ec.INF
    pts[5] = pts[4]
    pts[7] = ec.affine_neg(F, pts[6])
    sc = [rng.randrange(c.r) for _ in range(n)]
    sc[0], sc[1], sc[2] = 0, c.r - 1, 1
    sc[6] = sc[7] = 12345
    for i in range(20, 45):
        sc[i] = 3                      # 25 entries in one bucket of the lowest window
    exp = ec.msm_naive(F, pts, sc)
    PA, SA = ec.pack_points(c, group, pts), ff.pack_elements(sc, c.r, c.fr_limbs)
    # negative thread counts: the same loop with the accumulator in emulated shared memory (GB200_MSM_PERSISTENT=2)
    for (cw, pre, tl, ch, threads) in ((4, 0, 3, 4, 1), (4, 0, 3, 4, 7), (7, 1, 2, 16, 64), (5, 1, 4, 8, 5000), (4, 0, 3, 4, -5)):
        out = np.zeros(3 * F.degree * c.fp_limbs, dtype=np.uint64)
        assert hostemu.emu_msm_persistent(c.curve_id, group, P(PA), P(SA), n, cw, pre, tl, ch, threads, P(out)) == 0
        assert ec.from_jac(F, ec.unpack_points(c, group, out, ncoords=3)[0]) == exp, (c.name, group, cw, pre, threads)


@pytest.mark.parametrize("c,group", [(CURVES["bn254"], 2), (CURVES["bls12-381"], 1), (CURVES["bls12-377"], 2), (CURVES["bw6-761"], 1)],
                         ids=lambda v: getattr(v, "name", str(v)))
def test_msm_shared_memory_accumulator_logic(hostemu, c, group):
    """opt-in GB200_MSM_SMEM_ACC: the XYZZ accumulator of a task kept in strided shared-memory words (SmemXYZZ) -
    same edge cases as test_msm_logic (infinity base, equal points -> doubling, P and -P -> cancellation, zero / r-1
    scalars); the neighbouring threads' words must stay untouched"""
    rng = random.Random(50 + group)
    F, base = pick_base(c, group, rng)
    n = 37
    pts = [ec.scalar_mul(F, rng.randrange(1, 1 << 40), base) for _ in range(n)]
    pts[3] = ec.INF
    pts[5] = pts[4]
    pts[7] = ec.affine_neg(F, pts[6])
    sc = [rng.randrange(c.r) for _ in range(n)]
    sc[0], sc[1], sc[2] = 0, c.r - 1, 1
    sc[4] = sc[5]
    sc[6] = sc[7] = 12345
    exp = ec.msm_naive(F, pts, sc)
    PA, SA = ec.pack_points(c, group, pts), ff.pack_elements(sc, c.r, c.fr_limbs)
    for (cw, pre, tl, ch) in ((4, 0, 3, 4), (7, 1, 2, 16)):
        if pre and c.fp_limbs > 6:
            continue
        out = np.zeros(3 * F.degree * c.fp_limbs, dtype=np.uint64)
        assert hostemu.emu_msm_smem(c.curve_id, group, P(PA), P(SA), n, cw, pre, tl, ch, P(out)) == 0
        assert ec.from_jac(F, ec.unpack_points(c, group, out, ncoords=3)[0]) == exp, (c.name, group, cw, pre)


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_ntt_logic(hostemu, c):
    rng = random.Random(4)
    for logn in (1, 3, 6, 12, 13):
        n = 1 << logn
        dom = ntt.Domain(c, n)
        a = [rng.randrange(c.r) for _ in range(n)]
        A0 = ff.pack_elements(a, c.r, c.fr_limbs)
        for inv in (0, 1):
            for dec in (0, 1):
                for cos in (0, 1):
                    if logn >= 12 and (cos == 0 and dec == 1):
                        continue
                    A = A0.copy()
                    assert hostemu.emu_ntt(c.curve_id, P(A), logn, inv, dec, cos, None, None) == 0
                    exp = (dom.fft_inverse if inv else dom.fft)(a, dec, on_coset=bool(cos))
                    assert ff.unpack_elements(A, c.r, c.fr_limbs) == exp, (c.name, logn, inv, dec, cos)


@pytest.mark.parametrize("c", [CURVES["bn254"], CURVES["bw6-761"]], ids=lambda c: c.name)
def test_ntt_smaller_tiles(hostemu, c):
    """GB200_NTT_TILE_LOG: the same pass / tile / stage walk with smaller tiles (more passes, other cb splits)"""
    rng = random.Random(14)
    try:
        for tile_log, logn in ((10, 12), (10, 13), (6, 9), (4, 10), (7, 7)):
            assert hostemu.emu_ntt_set_tile_log(tile_log) == 0
            n = 1 << logn
            dom = ntt.Domain(c, n)
            a = [rng.randrange(c.r) for _ in range(n)]
            A0 = ff.pack_elements(a, c.r, c.fr_limbs)
            for inv, dec, cos in ((0, 0, 0), (1, 0, 1), (0, 1, 1), (1, 1, 0)):
                A = A0.copy()
                assert hostemu.emu_ntt(c.curve_id, P(A), logn, inv, dec, cos, None, None) == 0
                exp = (dom.fft_inverse if inv else dom.fft)(a, dec, on_coset=bool(cos))
                assert ff.unpack_elements(A, c.r, c.fr_limbs) == exp, (c.name, tile_log, logn, inv, dec, cos)
    finally:
        hostemu.emu_ntt_set_tile_log(11)


@pytest.mark.parametrize("c", [CURVES["bn254"], CURVES["bw6-761"]], ids=lambda c: c.name)
def test_ntt_register_rounds(hostemu, c):
    """opt-in GB200_NTT_RADIX8: up to three stages per shared-memory exchange, a group of 8 elements per thread in
    registers (ntt_round / ntt_round_dispatch), walked thread by thread over padded shared-memory slots - every stage
    count 1..11 per pass (round plans 3+3+3+2, 3+3+2+2, 2+2, 1 ...), passes with carried-along contiguous bits, all
    four mode combinations"""
    rng = random.Random(15)
    try:
        assert hostemu.emu_ntt_set_radix8(1) == 0
        for tile_log, logn in ((11, 1), (11, 2), (11, 3), (11, 4), (11, 5), (11, 7), (11, 10), (11, 11), (11, 13),
                               (10, 12), (6, 9), (4, 10), (7, 7), (6, 14 if c.fr_limbs <= 4 else 12)):
            assert hostemu.emu_ntt_set_tile_log(tile_log) == 0
            n = 1 << logn
            dom = ntt.Domain(c, n)
            a = [rng.randrange(c.r) for _ in range(n)]
            A0 = ff.pack_elements(a, c.r, c.fr_limbs)
            for inv, dec, cos in ((0, 0, 0), (1, 0, 1), (0, 1, 1), (1, 1, 0)):
                A = A0.copy()
                assert hostemu.emu_ntt(c.curve_id, P(A), logn, inv, dec, cos, None, None) == 0
                exp = (dom.fft_inverse if inv else dom.fft)(a, dec, on_coset=bool(cos))
                assert ff.unpack_elements(A, c.r, c.fr_limbs) == exp, (c.name, tile_log, logn, inv, dec, cos)
    finally:
        hostemu.emu_ntt_set_tile_log(11)
        hostemu.emu_ntt_set_radix8(0)


# ---- FP64-pipe path (field52.cuh / curve52.cuh): 52-bit limbs, DFMA round-toward-zero products ----
def _limbs52(v, L):
    return np.array([(v >> (52 * i)) & ((1 << 52) - 1) for i in range(L)], dtype=np.uint64)


def _l52(q):
    bits = q.bit_length()
    L = (bits + 51) // 52
    if 52 * L - bits < 5:
        L += 1
    return L


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_f52_mul_lazy_bounds_and_roundtrip(hostemu, c):
    rng = random.Random(1)
    for which, (q, L64) in enumerate(((c.p, c.fp_limbs), (c.r, c.fr_limbs))):
        fid = c.curve_id * 2 + which
        L = _l52(q)
        R = 1 << (52 * L)
        bound = 8 * q if 52 * L - q.bit_length() >= 6 else 4 * q     # lazy-reduction input bound
        for t in range(100):
            a, b = rng.randrange(bound), rng.randrange(bound)
            if t == 0: a = b = 0
            if t == 1: a = b = bound - 1
            if t == 2: a, b = q, q - 1
            out = np.zeros(L, dtype=np.uint64)
            assert hostemu.emu_f52_mul(fid, P(_limbs52(a, L)), P(_limbs52(b, L)), P(out)) == L
            got = sum(int(out[i]) << (52 * i) for i in range(L))
            assert all(int(x) < (1 << 52) for x in out)                 # normalised limbs
            assert got % q == a * b * pow(R, -1, q) % q and got < 2 * q   # value and output bound
        for t in range(40):                                             # gnark layout <-> 52-bit form
            v = [0, q - 1, 1][t] if t < 3 else rng.randrange(q)
            A = ff.pack_elements([v], q, L64)
            O = np.zeros_like(A)
            hostemu.emu_f52_roundtrip(fid, P(A), P(O))
            assert np.array_equal(A, O)


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_msm_fp64_path_logic(hostemu, c):
    """same edge cases as test_msm_logic, bucket accumulation through XYZZ52 (precomputed-table mode)"""
    rng = random.Random(19)
    F, base = pick_base(c, 1, rng)
    n = 37
    pts = [ec.scalar_mul(F, rng.randrange(1, 1 << 40), base) for _ in range(n)]
    pts[3] = ec.INF
    pts[5] = pts[4]
    pts[7] = ec.affine_neg(F, pts[6])
    sc = [rng.randrange(c.r) for _ in range(n)]
    sc[0], sc[1], sc[2] = 0, c.r - 1, 1
    sc[4] = sc[5]
    sc[6] = sc[7] = 12345
    sc[8], sc[9] = 1 << 15, (1 << 16) - 1
    exp = ec.msm_naive(F, pts, sc)
    PA, SA = ec.pack_points(c, 1, pts), ff.pack_elements(sc, c.r, c.fr_limbs)
    for (cw, tl, ch) in ((7, 2, 16), (13, 5, 100), (4, 3, 4)):
        if c.fp_limbs > 6 and cw != 7:
            continue                                        # 24-limb field (opt-in path): one configuration
        out = np.zeros(3 * c.fp_limbs, dtype=np.uint64)
        assert hostemu.emu_msm52(c.curve_id, P(PA), P(SA), n, cw, tl, ch, P(out)) == 0
        assert ec.from_jac(F, ec.unpack_points(c, 1, out, ncoords=3)[0]) == exp, (c.name, cw)


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
@pytest.mark.parametrize("group", (1, 2))
def test_msm_batched_affine_levels_logic(hostemu, c, group):
    """opt-in GB200_MSM_BATCH_AFFINE: batched-affine tree levels (msm_batch.cuh) in front of the XYZZ accumulate,
    walked over the launch geometry on the CPU.  Inputs hit every branch of the affine addition: equal points in
    one bucket (tangent), P and -P (infinity, then infinity + point at the next level), (0,0) bases, odd bucket
    sizes, empty buckets, more levels than any bucket needs, batches that straddle buckets."""
    if group == 2 and c.fp_limbs > 6:
        pytest.skip("BW6-761 G2 shares the Fp instantiation with G1")
    if group == 1 and c.name == "bls12-377":
        pytest.skip("same limb shape as BLS12-381 G1")
    rng = random.Random(41 + group)
    F, base = pick_base(c, group, rng)
    n = 90
    pts = [ec.scalar_mul(F, rng.randrange(1, 1 << 40), base) for _ in range(12)]
    pts = [pts[i % 12] for i in range(n)]                 # heavy repetition: equal points meet in buckets
    pts[3] = ec.INF
    pts[7] = ec.affine_neg(F, pts[6])
    sc = [rng.randrange(c.r) for _ in range(n)]
    for i in range(12, 60):                                 # same scalar for the copies of a base -> P + P, then 2P + 2P
        sc[i] = sc[i % 12]
    sc[0], sc[1], sc[2] = 0, c.r - 1, 1
    sc[6] = sc[7] = 12345                                   # P and -P with the same digits
    sc[18] = sc[19] = 12345                                 # ... twice: (P - P) + (P - P)
    exp = ec.msm_naive(F, pts, sc)
    PA, SA = ec.pack_points(c, group, pts), ff.pack_elements(sc, c.r, c.fr_limbs)
    for (cw, pre, tl, ch, levels) in ((4, 0, 3, 4, 1), (4, 1, 2, 4, 3), (6, 1, 4, 8, 12), (3, 0, 64, 2, 2)):
        if c.fp_limbs > 6 and (pre or levels == 2):
            continue                                        # BW6 (opt-in path): the two plain-table configurations
        out = np.zeros(3 * c.fp_limbs * (2 if (group == 2 and c.fp2_nonresidue is not None) else 1), dtype=np.uint64)
        assert hostemu.emu_msm_ba(c.curve_id, group, P(PA), P(SA), n, cw, pre, tl, ch, levels, P(out)) == 0
        assert ec.from_jac(F, ec.unpack_points(c, group, out, ncoords=3)[0]) == exp, (c.name, group, cw, pre, levels)


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
@pytest.mark.parametrize("group", (1, 2))
def test_fixed_base_batch_logic(hostemu, c, group):
    """fixed_base.cuh (windowed table of one base, signed digits, batched XYZZ -> affine conversion) walked on the
    CPU, against big-int scalar multiplication; scalars 0, 1, r-1, all-ones digits, a chunk boundary"""
    if group == 2 and c.fp_limbs > 6:
        pytest.skip("BW6-761 G2 shares the Fp instantiation with G1")
    rng = random.Random(31 + group)
    F, base = pick_base(c, group, rng)
    n = 19                                   # one full chunk of 16 + a ragged one
    ks = [rng.randrange(c.r) for _ in range(n)]
    ks[0], ks[1], ks[2], ks[3] = 0, 1, c.r - 1, (1 << (c.r.bit_length() - 1)) - 1
    KS = ff.pack_elements(ks, c.r, c.fr_limbs)
    BA = ec.pack_points(c, group, [base])
    for cw in (2, 5, 8):
        out = np.zeros((n, 2 * c.fp_limbs * (2 if (group == 2 and c.fp2_nonresidue is not None) else 1)), dtype=np.uint64)
        assert hostemu.emu_fixed_base(c.curve_id, group, P(BA), P(KS), n, cw, P(out)) == 0
        got = ec.unpack_points(c, group, out)
        for i in (0, 1, 2, 3, 7, 15, 16, 18):
            assert got[i] == ec.scalar_mul(F, ks[i], base), (c.name, group, cw, i)


@pytest.mark.parametrize("c", [CURVES["bn254"], CURVES["bls12-381"]], ids=lambda c: c.name)
def test_msm_hybrid_split_logic(hostemu, c):
    """opt-in hybrid accumulate (GB200_MSM_HYBRID): the launch geometry of the two split kernels walked block by
    block - every task produced exactly once, by the 32-bit-limb path or the FP64-pipe path according to its
    block, same result as the oracle.  Enough entries for several 128-task blocks."""
    rng = random.Random(29)
    F, base = pick_base(c, 1, rng)
    n = 300
    ks = [rng.randrange(1, 1 << 48) for _ in range(n)]
    KS = ff.pack_elements(ks, c.r, c.fr_limbs)
    from oracle import corelib
    PA = corelib.fixed_base(c, 1, ec.pack_points(c, 1, [base]), KS)
    sc = [rng.randrange(c.r) for _ in range(n)]
    sc[0], sc[1], sc[2] = 0, c.r - 1, 1
    SA = ff.pack_elements(sc, c.r, c.fr_limbs)
    exp = ec.scalar_mul(F, sum(s * k for s, k in zip(sc, ks)) % c.r, base)
    for (cw, tl, ch, k52) in (((5, 2, 4, 5), (6, 3, 8, 15)) if c.fp_limbs == 4 else ((6, 3, 8, 11),)):
        out = np.zeros(3 * c.fp_limbs, dtype=np.uint64)
        assert hostemu.emu_msm_hybrid(c.curve_id, P(PA), P(SA), n, cw, tl, ch, k52, P(out)) == 0
        assert ec.from_jac(F, ec.unpack_points(c, 1, out, ncoords=3)[0]) == exp, (c.name, cw, k52)


@pytest.mark.parametrize("c", [CURVES["bn254"], CURVES["bls12-381"], CURVES["bw6-761"]], ids=lambda c: c.name)
def test_plonk_constraint_kernel_logic(hostemu, c):
    """per-point logic of k_plonk_constraints (gate + permutation + L1 with blinding, bit-reversed scatter)
    walked on the CPU, against oracle/plonk.py (restating plonk/bn254/prove.go:841-1123)."""
    from oracle import plonk
    rng = random.Random(77)
    r, L = c.r, c.fr_limbs
    logn, rho = 4, 4
    n = 1 << logn
    polys = {k: [rng.randrange(r) for _ in range(n)] for k in plonk.POLYS}
    alpha, beta, gamma = (rng.randrange(r) for _ in range(3))
    blind = {"l": [rng.randrange(r) for _ in range(2)], "r": [rng.randrange(r) for _ in range(2)],
             "o": [rng.randrange(r) for _ in range(2)], "z": [rng.randrange(r) for _ in range(3)]}
    want = plonk.numerator(c, n, rho, polys, alpha, beta, gamma, blind)
    dom0, dom1 = ntt.Domain(c, n), ntt.Domain(c, rho * n)
    out = np.zeros((rho * n, L), dtype=np.uint64)
    pe = lambda v: ff.pack_elements(v, r, L)
    abg = pe([alpha, beta, gamma])
    bl = [pe(blind[k]) for k in ("l", "r", "o", "z")]
    nbl = (ctypes.c_int * 4)(2, 2, 2, 3)
    blp = (ctypes.c_void_p * 4)(*[b.ctypes.data for b in bl])
    for i in range(rho):
        coset = dom1.coset_gen * pow(dom1.generator, i, r) % r
        on_coset = [pe(plonk.coset_values(c, dom0, polys[k], coset)) for k in plonk.POLYS]
        pp = (ctypes.c_void_p * 12)(*[a.ctypes.data for a in on_coset])
        assert hostemu.emu_plonk_constraints_coset(c.curve_id, pp, P(abg), blp, nbl, logn, i, rho, P(out)) == 0
    assert ff.unpack_elements(out, r, L) == want
    # BSB22 commitment gates: k_plonk_add_bsb22 completes the gate term afterwards, one call per commitment and coset
    bsb = [([rng.randrange(r) for _ in range(n)], [rng.randrange(r) for _ in range(n)]) for _ in range(2)]
    want2 = plonk.numerator(c, n, rho, polys, alpha, beta, gamma, blind, bsb22=bsb)
    assert want2 != want
    for i in range(rho):
        coset = dom1.coset_gen * pow(dom1.generator, i, r) % r
        for qcp, pi2 in bsb:
            Q, PI = pe(plonk.coset_values(c, dom0, qcp, coset)), pe(plonk.coset_values(c, dom0, pi2, coset))
            assert hostemu.emu_plonk_bsb22(c.curve_id, P(Q), P(PI), P(out), logn, i, rho) == 0
    assert ff.unpack_elements(out, r, L) == want2
