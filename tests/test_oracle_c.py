"""C++ oracle (oracle/c/oracle.cpp) validated against the big-int Python oracle."""
import random

import pytest

from oracle import corelib, ec, ff, ntt
from oracle.params import CURVES
from util import pick_base

ALL = list(CURVES.values())


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
@pytest.mark.parametrize("group", (1, 2))
def test_msm_and_fixed_base(c, group):
    rng = random.Random(11 + group)
    F, base = pick_base(c, group, rng)
    n = 41
    ks = [rng.randrange(1, c.r) for _ in range(n)]
    KS = ff.pack_elements(ks, c.r, c.fr_limbs)
    PTS = corelib.fixed_base(c, group, ec.pack_points(c, group, [base]), KS)
    pts = ec.unpack_points(c, group, PTS)
    for i in (0, 1, 17, 40):
        assert pts[i] == ec.scalar_mul(F, ks[i], base)
    sc = [rng.randrange(c.r) for _ in range(n)]
    sc[0], sc[1], sc[2] = 0, c.r - 1, 1
    SC = ff.pack_elements(sc, c.r, c.fr_limbs)
    exp = ec.scalar_mul(F, sum(s * k for s, k in zip(sc, ks)), base)
    for cw in (3, 8, 13):
        got = ec.from_jac(F, ec.unpack_points(c, group, corelib.msm(c, group, PTS, SC, c=cw), ncoords=3)[0])
        assert got == exp
    got = ec.from_jac(F, ec.unpack_points(c, group, corelib.msm_naive(c, group, PTS, SC), ncoords=3)[0])
    assert got == exp


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
@pytest.mark.parametrize("group", (1, 2))
def test_msm_batch_affine_variant(c, group):
    """the CPU arm's batch-affine buckets (gnark-crypto's large-window path) against the big-int oracle, on inputs
    that hit every special case of an affine addition: repeated points (doubling), P and -P in one bucket
    (cancellation), points at infinity, zero scalars, more operations than one batch, bucket collisions."""
    rng = random.Random(21 + group)
    F, base = pick_base(c, group, rng)
    n = 700
    ks = [rng.randrange(1, c.r) for _ in range(n)]
    for i in range(0, 60, 3):            # same base three times in a row ...
        ks[i + 1] = ks[i]
        ks[i + 2] = c.r - ks[i]          # ... and its negative
    KS = ff.pack_elements(ks, c.r, c.fr_limbs)
    PTS = corelib.fixed_base(c, group, ec.pack_points(c, group, [base]), KS)
    PTS[100] = 0                          # (0,0) = infinity
    ks[100] = 0
    sc = [rng.randrange(c.r) for _ in range(n)]
    for i in range(0, 60, 3):            # identical digits -> identical buckets in every window
        sc[i + 1] = sc[i + 2] = sc[i]
    sc[200], sc[201], sc[202] = 0, c.r - 1, 1
    for i in range(300, 400):             # tiny scalars: one hot bucket in the lowest window only
        sc[i] = 1 + (i & 1)
    SC = ff.pack_elements(sc, c.r, c.fr_limbs)
    exp = ec.scalar_mul(F, sum(s * k for s, k in zip(sc, ks)) % c.r, base)
    for cw in (4, 10, 13):
        for threads in (1, 3):
            got = corelib.msm(c, group, PTS, SC, c=cw, nthreads=threads, batch_affine=True)
            assert ec.from_jac(F, ec.unpack_points(c, group, got, ncoords=3)[0]) == exp, (cw, threads)


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_ntt(c):
    rng = random.Random(5)
    for logn in (0, 1, 5, 9):
        n = 1 << logn
        dom = ntt.Domain(c, n)
        a = [rng.randrange(c.r) for _ in range(n)]
        for inv in (0, 1):
            for dec in (0, 1):
                for cos in (0, 1):
                    A = ff.pack_elements(a, c.r, c.fr_limbs)
                    corelib.ntt(c, A, logn, inv, dec, cos)
                    exp = (dom.fft_inverse if inv else dom.fft)(a, dec, on_coset=bool(cos))
                    assert ff.unpack_elements(A, c.r, c.fr_limbs) == exp


def test_compute_h_and_dot():
    from oracle import groth16 as g16
    c = CURVES["bn254"]
    rng = random.Random(8)
    n = 64
    a = [rng.randrange(c.r) for _ in range(n)]
    b = [rng.randrange(c.r) for _ in range(n)]
    cc = [rng.randrange(c.r) for _ in range(n)]
    dom = ntt.Domain(c, n)
    want = g16.compute_h(dom, a, b, cc)
    A, B, C = (ff.pack_elements(v, c.r, c.fr_limbs) for v in (a, b, cc))
    corelib.compute_h(c, A, B, C, 6)
    assert ff.unpack_elements(A, c.r, c.fr_limbs) == want
    assert corelib.fr_dot(c, ff.pack_elements(a, c.r, 4), ff.pack_elements(b, c.r, 4)) == sum(x * y for x, y in zip(a, b)) % c.r
