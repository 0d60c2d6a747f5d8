"""The device templates (field.cuh / curve.cuh / msm.cuh / ntt.cuh), compiled for the
host with the PTX carry chains emulated, against the oracle.  This checks the
kernels' per-thread logic on a box without a GPU; the -m gpu tests check the real
thing through the C ABI."""
import ctypes
import os
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
        Rinv = pow(R, -1, q)
        ts = [0, q * R - 1, q * q, (q - 1) * (q - 1), 6 * q * q if 6 * q < R else q * q]
        ts += [rng.randrange(q * R) for _ in range(60)]
        for t in ts:
            O = np.zeros(L, dtype=np.uint64)
            assert hostemu.emu_wide_op(fid, 2, P(arr(t, 2 * L)), P(arr(0, L)), P(O)) == 0
            assert val(O) == t * Rinv % q, (c.name, which, "mont_reduce_wide")


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_mont_reduce_wide_saturated_high_words(hostemu, c):
    """Directed regression for the carry that round 2's first mont_reduce_wide lost once in ~2^33 reductions (a word of
    the HIGH half of T equal to 0xffffffff under a chain-end carry, field.cuh): T with saturated 32-bit words in its high
    half - every position, runs of them, and the all-ones high half that still satisfies T < p R - against big ints."""
    rng = random.Random(4242)
    for which, (q, L) in enumerate(((c.p, c.fp_limbs), (c.r, c.fr_limbs))):
        fid = c.curve_id * 2 + which
        R = 1 << (64 * L)
        Rinv = pow(R, -1, q)
        arr = lambda x, limbs: np.frombuffer(int(x).to_bytes(8 * limbs, "little"), dtype=np.uint64).copy()
        val = lambda a: int.from_bytes(a.tobytes(), "little")
        W = 2 * L                     # 32-bit words of one half
        ts = []
        for _ in range(40):
            lo = rng.randrange(R)
            hi = rng.randrange(q)
            for k in range(W):        # one saturated word at every position of the high half, then runs
                t = lo + (((hi | (0xFFFFFFFF << (32 * k))) % R) << (64 * L))
                if t < q * R:
                    ts.append(t)
            for k in range(1, W):
                t = lo + ((((1 << (32 * k)) - 1) | (hi >> (32 * k) << (32 * k))) << (64 * L))
                if t < q * R:
                    ts.append(t)
        ts.append(q * R - 1)
        ts.append((R - 1) + (((q - 1) | 0xFFFFFFFF) << (64 * L)) if ((q - 1) | 0xFFFFFFFF) < q else q * R - 1)
        assert len(ts) > 500
        for t in ts:
            O = np.zeros(L, dtype=np.uint64)
            assert hostemu.emu_wide_op(fid, 2, P(arr(t, 2 * L)), P(arr(0, L)), P(O)) == 0
            assert val(O) == t * Rinv % q, (c.name, which, hex(t))


def test_g2_doubling_chain_known_failure(hostemu):
    """The BN254 G2 point whose precomputed slabs 11..15 came out wrong on the GPU (found by the verified 2^20 Groth16
    proof of bench.py: index 116963 of G2.B, seed 20): the device table build - an XYZZ doubling chain that is never
    normalised, then one inversion per point - emulated on the CPU, every slab against 2^(16 w) P of the big-int oracle;
    and the 1-point MSM of the scalar it was paired with."""
    c = CURVES["bn254"]
    F = ff.base_field(c, 2)
    here = os.path.dirname(os.path.abspath(__file__))
    pt = np.load(os.path.join(here, "golden", "bn254_g2_lost_carry_point.npy"))
    sc = np.load(os.path.join(here, "golden", "bn254_g2_lost_carry_scalar.npy"))
    Pt = ec.unpack_points(c, 2, pt)[0]
    out = np.zeros((16, 4 * c.fp_limbs), dtype=np.uint64)
    assert hostemu.emu_precompute(c.curve_id, 2, P(pt), 16, 16, P(out)) == 0
    for w in range(16):
        assert ec.unpack_points(c, 2, out[w])[0] == ec.scalar_mul(F, 1 << (16 * w), Pt), w
    s = ff.unpack_elements(sc, c.r, c.fr_limbs)[0]
    res = np.zeros(6 * c.fp_limbs, dtype=np.uint64)
    assert hostemu.emu_msm(c.curve_id, 2, P(pt), P(sc), 1, 16, 1, 32, 8, P(res)) == 0
    assert ec.from_jac(F, ec.unpack_points(c, 2, res, ncoords=3)[0]) == ec.scalar_mul(F, s, Pt)


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
        for op, exp in ((2, F2.mul(a, b)), (5, F2.sqr(a)), (3, F2.inv(a)), (0, F2.add(a, b)), (1, F2.sub(a, b))):
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


def _limbs52(v, L):
    return np.array([(v >> (52 * i)) & ((1 << 52) - 1) for i in range(L)], dtype=np.uint64)


def _l52(q):
    bits = q.bit_length()
    L = (bits + 51) // 52
    if 52 * L - bits < 5:
        L += 1
    return L


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


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_point_decoding_logic(hostemu, c):
    """points_decode.cuh (per-thread logic of k_points_decode, host build) against oracle/encoding.py: gnark-crypto's raw
    and compressed G1 encodings and the raw G2 encoding - random points, the point at infinity, both roots of y, and
    the rejected inputs (wrong metadata bits, x not on the curve, coordinate >= p)"""
    from oracle import encoding, derive
    rng = random.Random(77)
    F1 = ff.Fp(c.p)
    b_small = {"bn254": 3, "bls12-381": 4, "bls12-377": 1, "bw6-761": -1}[c.name]
    hostemu.emu_decode_points.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    if c.g1 is not None:
        pts = [ec.scalar_mul(F1, rng.randrange(1, c.r), c.g1) for _ in range(12)] + [None]
    else:           # no generator recalled for this curve (BW6-761): random points of y^2 = x^3 + b
        pts = []
        while len(pts) < 12:
            x = rng.randrange(c.p)
            y = encoding.sqrt_fp(c.p, x ** 3 + b_small)
            if y is not None:
                pts.append((x, y if rng.random() < 0.5 else c.p - y))
        pts.append(None)
    pts.append(ec.affine_neg(F1, pts[0]))                  # the other root for the same x
    for compressed in (False, True):
        blob = b"".join(encoding.encode_g1(c, P_, compressed) for P_ in pts)
        raw = np.frombuffer(blob, dtype=np.uint8).copy()
        out = np.zeros((len(pts), 2 * c.fp_limbs), dtype=np.uint64)
        rc = hostemu.emu_decode_points(c.curve_id, 1, P(raw), len(pts), 2 if compressed else 1, P(out))
        assert rc == 0, (c.name, compressed)                 # BLS12-377 (p = 1 mod 4): Tonelli-Shanks
        assert ec.unpack_points(c, 1, out) == pts, (c.name, compressed)
        assert [encoding.decode_g1(c, blob[i * len(blob) // len(pts):(i + 1) * len(blob) // len(pts)]) for i in range(len(pts))] == pts
    # rejected inputs
    one = lambda bts, enc: hostemu.emu_decode_points(c.curve_id, 1, P(np.frombuffer(bts, dtype=np.uint8).copy()), 1, enc,
                                                     P(np.zeros(2 * c.fp_limbs, dtype=np.uint64)))
    good = encoding.encode_g1(c, pts[0], False)
    bad_curve = good[:-1] + bytes([good[-1] ^ 1])          # y changed: not on the curve
    assert one(bad_curve, 1) == 3
    too_big = (c.p).to_bytes(len(good) // 2, "big") + good[len(good) // 2:]
    if c.p.bit_length() % 8 != 0 and (c.p >> (8 * (len(good) // 2) - (2 if c.name == "bn254" else 3))) == 0:
        assert one(too_big, 1) == 2                        # x = p is not reduced
    flagged = bytes([good[0] | 0x80]) + good[1:]           # "compressed" bit on an uncompressed point
    assert one(flagged, 1) == 1
    if True:
        comp = encoding.encode_g1(c, pts[0], True)
        # find an x that is not on the curve
        x = 5
        while pow((x ** 3 + (b_small % c.p)) % c.p, (c.p - 1) // 2, c.p) == 1:
            x += 1
        off_curve = encoding.encode_g1(c, (x, 0), True, check=False)
        assert one(off_curve, 2) == 3
        assert one(bytes([comp[0] & 0x3f if c.name == "bn254" else comp[0] & 0x1f]) + comp[1:], 2) == 1   # flags cleared
    # G2, raw (A1 || A0 ordering)
    if c.fp2_nonresidue is not None:
        F2 = ff.base_field(c, 2)
        _, g2 = pick_base(c, 2, rng)
        g2pts = [ec.scalar_mul(F2, rng.randrange(1, c.r), g2) for _ in range(4)] + [None]
        blob = b"".join(encoding.encode_g2_raw(c, Q) for Q in g2pts)
        out = np.zeros((len(g2pts), 4 * c.fp_limbs), dtype=np.uint64)
        assert hostemu.emu_decode_points(c.curve_id, 2, P(np.frombuffer(blob, dtype=np.uint8).copy()), len(g2pts), 1, P(out)) == 0
        assert ec.unpack_points(c, 2, out) == g2pts


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_g2_compressed_decoding(hostemu, c):
    """compressed G2 through points_decode.cuh (square roots in Fp2 by the norm, Tonelli-Shanks under them for
    BLS12-377; BW6-761: G2 over Fp on y^2 = x^3 + 4) against oracle/encoding.py - random points of the twist, both roots,
    infinity, an x that is not on the twist; and the EXTERNAL vectors: the compressed G2 generators inside gnark's
    serialised verifying keys (BN254, BLS12-381) and the 65 compressed G2 points of the Ethereum KZG ceremony file."""
    import json
    import os
    from oracle import encoding
    rng = random.Random(99)
    hostemu.emu_decode_points.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    p = c.p
    deg = c.g2_degree
    F2 = ff.base_field(c, 2)
    bt = encoding.twist_b(c)
    if c.g2 is not None:                       # the recalled generator is on the twist this module states
        x, y = c.g2
        assert F2.eq(F2.sqr(y), F2.add(F2.mul(F2.sqr(x), x), bt if deg == 2 else F2.from_int(bt)))
    pts = []
    while len(pts) < 10:                       # random points of the twist (the decoder does not check the subgroup)
        if deg == 2:
            x = (rng.randrange(p), rng.randrange(p) if len(pts) != 3 else 0)
            y = encoding.sqrt_fp2(c, F2.add(F2.mul(F2.sqr(x), x), bt))
        else:
            x = rng.randrange(p)
            y = encoding.sqrt_fp(p, x ** 3 + bt)
        if y is not None:
            pts.append((x, y))
    pts.append((pts[0][0], F2.neg(pts[0][1])))
    pts.append(None)
    if c.g2 is not None:
        pts.append(c.g2)
    blob = b"".join(encoding.encode_g2(c, Q, True) for Q in pts)
    per = len(blob) // len(pts)
    assert per == deg * 8 * c.fp_limbs
    assert [encoding.decode_g2(c, blob[i * per:(i + 1) * per]) for i in range(len(pts))] == pts
    out = np.zeros((len(pts), 2 * deg * c.fp_limbs), dtype=np.uint64)
    assert hostemu.emu_decode_points(c.curve_id, 2, P(np.frombuffer(blob, dtype=np.uint8).copy()), len(pts), 2, P(out)) == 0
    assert ec.unpack_points(c, 2, out) == pts
    # raw encoding of the same points (BW6-761 G2: the on-curve check must use b = 4, not G1's -1)
    rawb = b"".join(encoding.encode_g2(c, Q, False) for Q in pts)
    assert hostemu.emu_decode_points(c.curve_id, 2, P(np.frombuffer(rawb, dtype=np.uint8).copy()), len(pts), 1, P(out)) == 0
    assert ec.unpack_points(c, 2, out) == pts
    # an x that is not on the twist is refused
    while True:
        x = (rng.randrange(p), rng.randrange(p)) if deg == 2 else rng.randrange(p)
        rhs = F2.add(F2.mul(F2.sqr(x), x), bt) if deg == 2 else (x ** 3 + bt) % p
        if (encoding.sqrt_fp2(c, rhs) if deg == 2 else encoding.sqrt_fp(p, rhs)) is None:
            break
    fake = ((x, (1, 1)) if deg == 2 else (x, 1))
    bad = encoding.encode_g2(c, fake, True)
    assert hostemu.emu_decode_points(c.curve_id, 2, P(np.frombuffer(bad, dtype=np.uint8).copy()), 1, 2,
                                     P(np.zeros(2 * deg * c.fp_limbs, dtype=np.uint64))) == 3
    # external vectors
    here = os.path.dirname(os.path.abspath(__file__))
    vk = json.load(open(os.path.join(here, "golden", "gnark_vk_constants_v1.json")))
    name = {"bn254": "bn254", "bls12-381": "bls12-381"}.get(c.name)
    for key in vk["keys"]:
        if name and key["curve"] == name:
            raw = bytes.fromhex(key["kzg_g2_0_compressed"])
            assert encoding.decode_g2(c, raw) == c.g2 and encoding.encode_g2(c, c.g2, True) == raw
            o1 = np.zeros((1, 4 * c.fp_limbs), dtype=np.uint64)
            assert hostemu.emu_decode_points(c.curve_id, 2, P(np.frombuffer(raw, dtype=np.uint8).copy()), 1, 2, P(o1)) == 0
            assert ec.unpack_points(c, 2, o1) == [c.g2]
    if c.name == "bls12-381":
        from oracle import kzg_srs
        blob = open(kzg_srs.PATH, "rb").read()
        off = 2 * kzg_srs.N * 48
        g2b = blob[off:off + 96 * kzg_srs.N_G2]
        want = kzg_srs.load()[2]
        o65 = np.zeros((kzg_srs.N_G2, 4 * c.fp_limbs), dtype=np.uint64)
        assert hostemu.emu_decode_points(c.curve_id, 2, P(np.frombuffer(g2b, dtype=np.uint8).copy()), kzg_srs.N_G2, 2, P(o65)) == 0
        assert ec.unpack_points(c, 2, o65) == want
        assert b"".join(encoding.encode_g2(c, Q, True) for Q in want) == g2b


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_point_decoding_round_trip_many(hostemu, c):
    """wider net for the square roots of points_decode.cuh (exponentiation, Tonelli-Shanks with its data-dependent loop,
    the two candidate branches of the Fp2 root): 150 random points per group through encode (oracle) -> decode
    (device logic, host build), compressed, both groups"""
    from oracle import encoding
    rng = random.Random(2024)
    hostemu.emu_decode_points.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    p = c.p
    b1 = {"bn254": 3, "bls12-381": 4, "bls12-377": 1, "bw6-761": -1}[c.name]
    deg = c.g2_degree
    F2 = ff.base_field(c, 2)
    bt = encoding.twist_b(c)
    g1, g2 = [], []
    while len(g1) < 150:
        x = rng.randrange(p)
        y = encoding.sqrt_fp(p, x ** 3 + b1)
        if y is not None:
            g1.append((x, y if rng.random() < 0.5 else (p - y) % p))
    while len(g2) < 150:
        if deg == 2:
            x = (rng.randrange(p), rng.randrange(p))
            y = encoding.sqrt_fp2(c, F2.add(F2.mul(F2.sqr(x), x), bt))
        else:
            x = rng.randrange(p)
            y = encoding.sqrt_fp(p, x ** 3 + bt)
        if y is not None:
            g2.append((x, y if rng.random() < 0.5 else F2.neg(y)))
    for group, pts, enc in ((1, g1, lambda Q: encoding.encode_g1(c, Q, True)), (2, g2, lambda Q: encoding.encode_g2(c, Q, True))):
        blob = b"".join(enc(Q) for Q in pts)
        width = 2 * (deg if group == 2 else 1) * c.fp_limbs
        out = np.zeros((len(pts), width), dtype=np.uint64)
        assert hostemu.emu_decode_points(c.curve_id, group, P(np.frombuffer(blob, dtype=np.uint8).copy()), len(pts), 2, P(out)) == 0
        assert ec.unpack_points(c, group, out) == pts, (c.name, group)
