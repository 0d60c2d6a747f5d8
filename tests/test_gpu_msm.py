"""MSM parity through the C ABI on the GPU (bit-exact affine results).

Oracles: exhaustive small-N double-and-add (Python big ints),
mid-size C++ Pippenger, and at BASELINE size the known-discrete-log identity
sum s_i (k_i B) = (sum s_i k_i) B."""
import random

import numpy as np
import pytest

from oracle import corelib, ec, ff
from oracle.params import CURVES
from util import jac_to_affine, known_dlog_instance, pick_base

pytestmark = pytest.mark.gpu
ALL = list(CURVES.values())


def edge_case_instance(c, group, n, rng):
    F, base = pick_base(c, group, rng)
    pts = [ec.scalar_mul(F, rng.randrange(1, 1 << 48), base) for _ in range(n)]
    sc = [rng.randrange(c.r) for _ in range(n)]
    if n > 12:
        pts[3] = ec.INF                       # (0,0) base
        pts[5] = pts[4]                       # repeated point
        pts[7] = ec.affine_neg(F, pts[6])     # P and -P
        sc[0], sc[1], sc[2] = 0, c.r - 1, 1   # zero scalar, r-1, one
        sc[4] = sc[5]                         # same bucket, same point -> doubling path
        sc[6] = sc[7] = 0x1234567             # same buckets, opposite points -> infinity
        sc[8], sc[9], sc[10] = 1 << 15, (1 << 16) - 1, 1 << 16
    return F, pts, sc


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
@pytest.mark.parametrize("group", (1, 2))
@pytest.mark.parametrize("precomp", (False, True))
def test_small_exhaustive(gpu, c, group, precomp):
    rng = random.Random(100 + c.curve_id * 4 + group)
    for n in (1, 2, 33, 100):               # N=1, N not a multiple of the block
        F, pts, sc = edge_case_instance(c, group, n, rng)
        want = ec.msm_naive(F, pts, sc)
        t = gpu.Table(c.curve_id, group, ec.pack_points(c, group, pts), precomp=precomp)
        got = jac_to_affine(c, group, t.msm(ff.pack_elements(sc, c.r, c.fr_limbs)))
        assert got == want, (c.name, group, n, precomp)
        # sub-range + all-zero scalars + empty
        if n >= 33:
            got = jac_to_affine(c, group, t.msm(ff.pack_elements(sc[10:30], c.r, c.fr_limbs), off=10, n=20))
            assert got == ec.msm_naive(F, pts[10:30], sc[10:30])
            assert jac_to_affine(c, group, t.msm(np.zeros((n, c.fr_limbs), dtype=np.uint64))) is ec.INF
            assert jac_to_affine(c, group, t.msm(np.zeros((0, c.fr_limbs), dtype=np.uint64), n=0)) is ec.INF
        t.free()


def test_all_equal_bases_dummy_setup(gpu):
    """groth16.DummySetup (backend/groth16/bn254/setup.go:516-540): every base is the same point."""
    c = CURVES["bn254"]
    rng = random.Random(77)
    F = ff.Fp(c.p)
    n = 3000
    P = ec.scalar_mul(F, 987654321, c.g1)
    PTS = np.tile(ec.pack_points(c, 1, [P]), (n, 1))
    sc = [rng.randrange(c.r) for _ in range(n)]
    want = ec.scalar_mul(F, sum(sc) % c.r, P)
    for precomp in (False, True):
        t = gpu.Table(c.curve_id, 1, PTS, precomp=precomp)
        assert jac_to_affine(c, 1, t.msm(ff.pack_elements(sc, c.r, 4))) == want
        t.free()


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
@pytest.mark.parametrize("group", (1, 2))
def test_mid_vs_cpp_oracle(gpu, c, group):
    n = 1 << (14 if c.fp_limbs <= 6 else 12)
    F, base, PTS, SC, expected = known_dlog_instance(c, group, n, seed=0x6e61726b + c.curve_id * 2 + group)
    want = jac_to_affine(c, group, corelib.msm(c, group, PTS, SC))
    assert want == expected                                   # oracle vs known dlog
    for precomp in (False, True):
        t = gpu.Table(c.curve_id, group, PTS, precomp=precomp)
        assert jac_to_affine(c, group, t.msm(SC)) == want
        t.free()


def test_skewed_scalars(gpu):
    """50% zeros, 25% in {1,2}, rest uniform (SURVEY.md §8d config 2 variant): giant buckets."""
    c = CURVES["bn254"]
    n = 1 << 15
    F, base, PTS, SC, expected = known_dlog_instance(c, 1, n, seed=31337, skew=True)
    for precomp in (False, True):
        t = gpu.Table(c.curve_id, 1, PTS, precomp=precomp)
        assert jac_to_affine(c, 1, t.msm(SC)) == expected
        t.free()


def test_baseline_size_known_dlog(gpu):
    """BASELINE config 2: BN254 G1 MSM, N = 2^20, checked exactly by one scalar multiplication."""
    c = CURVES["bn254"]
    n = 1 << 20
    F, base, PTS, SC, expected = known_dlog_instance(c, 1, n, seed=0x6e61726b00000002)
    t = gpu.Table(c.curve_id, 1, PTS, precomp=True)
    assert t.info()["n"] == n
    assert jac_to_affine(c, 1, t.msm(SC)) == expected
    # linearity: MSM(2s) = 2 MSM(s)
    s2 = ff.pack_elements([2 * v % c.r for v in ff.unpack_elements(SC[:4096], c.r, 4)], c.r, 4)
    a = jac_to_affine(c, 1, t.msm(SC[:4096].copy(), n=4096))
    b = jac_to_affine(c, 1, t.msm(s2, n=4096))
    assert b == ec.affine_add(F, a, a)
    t.free()


def test_device_resident_scalars(gpu):
    import torch
    c = CURVES["bn254"]
    n = 5000
    F, base, PTS, SC, expected = known_dlog_instance(c, 1, n, seed=5)
    t = gpu.Table(c.curve_id, 1, PTS, precomp=True)
    d_sc = torch.from_numpy(SC.view(np.int64)).cuda()
    assert jac_to_affine(c, 1, t.msm(d_sc, n=n, on_device=True)) == expected
    t.free()


def test_errors(gpu):
    c = CURVES["bn254"]
    PTS = ec.pack_points(c, 1, [c.g1] * 8)
    t = gpu.Table(c.curve_id, 1, PTS, precomp=False)
    with pytest.raises(gpu.B200Error):
        t.msm(np.zeros((8, 4), dtype=np.uint64), off=4, n=8)     # range exceeds table
    with pytest.raises(gpu.B200Error):
        gpu.check(gpu.load().b200_msm_g2(t.handle, 0, 1, gpu.ptr(np.zeros(4, dtype=np.uint64)), 0,
                                         gpu.ptr(np.zeros(24, dtype=np.uint64))))
    t.free()
