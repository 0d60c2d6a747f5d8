"""Oracle self-checks (CPU).  Beyond the external fixtures of tests/test_golden_kzg.py the reference pins
nothing at the bit level for this path, so the oracle is also anchored on properties:
published constants property-checked, NTT against the O(n^2) definition, the
Groth16 pipeline against the pairing equation evaluated in the exponent (the
reference's own test is prove -> Verify, test/assert_checkcircuit.go:140-144)."""
import random

import pytest

from oracle import ec, ff, groth16 as g16, ntt
from oracle.params import CURVES

ALL = list(CURVES.values())


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_constants(c):
    r, p, s = c.r, c.p, c.two_adicity
    assert (r - 1) % (1 << s) == 0 and ((r - 1) >> s) % 2 == 1
    w = c.root_of_unity
    assert pow(w, 1 << s, r) == 1 and pow(w, 1 << (s - 1), r) == r - 1
    assert pow(c.mult_gen, (r - 1) // 2, r) == r - 1          # non-residue => g^n != 1 for n | 2^s
    assert p.bit_length() <= 64 * c.fp_limbs and r.bit_length() <= 64 * c.fr_limbs
    if c.g1 is not None:
        F = ff.Fp(p)
        assert ec.is_on_curve(F, c.g1, c.b % p)
        assert ec.scalar_mul(F, r, c.g1) is ec.INF
    if c.g2 is not None:
        F2 = ff.base_field(c, 2)          # Fp2, or Fp for BW6-761
        assert ec.scalar_mul(F2, r, c.g2) is ec.INF


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_ntt_conventions(c):
    rng = random.Random(3)
    for logn in (0, 1, 4, 6):
        n = 1 << logn
        dom = ntt.Domain(c, n)
        a = [rng.randrange(c.r) for _ in range(n)]
        ev = ntt.dft_naive(c, a, dom.generator)
        evc = ntt.dft_naive(c, a, dom.generator, dom.coset_gen)
        assert ntt.bit_reverse(dom.fft(a, ntt.DIF)) == ev            # DIF: natural -> bit-reversed
        assert dom.fft(ntt.bit_reverse(a), ntt.DIT) == ev            # DIT: bit-reversed -> natural
        assert ntt.bit_reverse(dom.fft_inverse(ev, ntt.DIF)) == a
        assert dom.fft_inverse(ntt.bit_reverse(ev), ntt.DIT) == a
        assert dom.fft(ntt.bit_reverse(a), ntt.DIT, on_coset=True) == evc
        assert ntt.bit_reverse(dom.fft(a, ntt.DIF, on_coset=True)) == evc
        assert ntt.bit_reverse(dom.fft_inverse(evc, ntt.DIF, on_coset=True)) == a
        assert dom.fft_inverse(ntt.bit_reverse(evc), ntt.DIT, on_coset=True) == a


def test_lagrange_srs_convention():
    """test/unsafekzg/kzgsrs.go:186-194: iFFT-DIF of (1,tau,tau^2,...) + BitReverse = L_i(tau)."""
    c = CURVES["bn254"]
    n = 8
    dom = ntt.Domain(c, n)
    tau = 123456789
    pows = [pow(tau, i, c.r) for i in range(n)]
    lag = ntt.bit_reverse(dom.fft_inverse(pows, ntt.DIF))
    # L_i(tau) = prod_{j != i} (tau - w^j) / (w^i - w^j)
    w = [pow(dom.generator, i, c.r) for i in range(n)]
    for i in range(n):
        num = den = 1
        for j in range(n):
            if j != i:
                num = num * (tau - w[j]) % c.r
                den = den * (w[i] - w[j]) % c.r
        assert lag[i] == num * pow(den, -1, c.r) % c.r


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_groth16_cubic_and_chain(c):
    """config 1 (examples/cubic, x=3 y=35: examples/cubic/cubic_test.go:47-50) restated."""
    for make, wit in ((g16.cubic_r1cs, lambda: g16.cubic_witness(c.r)),
                      (lambda: g16.square_chain_r1cs(13), lambda: g16.square_chain_witness(c.r, 13))):
        cs, W = make(), wit()
        A, B, C = g16.solve_abc(cs, W, c.r)
        assert all((x * y - z) % c.r == 0 for x, y, z in zip(A, B, C))
        pk = g16.setup_dlog(c, cs, g16.random_toxic(c, 7))
        pr = g16.prove_dlog(c, cs, pk, W, 0x1234567, 0x7654321)
        assert g16.verify_dlog(c, cs, pk, pr, W)
        assert pr.h[pk.domain.n - 1] == 0                 # deg h <= n-2 (prove.go:225)
        W2 = list(W); W2[2] = (W2[2] + 1) % c.r
        assert not g16.verify_dlog(c, cs, pk, g16.prove_dlog(c, cs, pk, W2, 1, 2), W2)


def test_compute_h_identity():
    """h(X) (X^n - 1) = A(X)B(X) - C(X) at a random point."""
    c = CURVES["bn254"]
    cs = g16.square_chain_r1cs(29)
    W = g16.square_chain_witness(c.r, 29)
    A, B, C = g16.solve_abc(cs, W, c.r)
    dom = ntt.Domain(c, cs.nb_constraints)
    n = dom.n
    h = ntt.bit_reverse(g16.compute_h(dom, A, B, C))
    pad = [0] * (n - len(A))
    ca = ntt.bit_reverse(dom.fft_inverse(A + pad, ntt.DIF))
    cb = ntt.bit_reverse(dom.fft_inverse(B + pad, ntt.DIF))
    cc = ntt.bit_reverse(dom.fft_inverse(C + pad, ntt.DIF))
    x = 0xabcdef1234567
    r = c.r
    lhs = ntt.poly_eval(r, h, x) * (pow(x, n, r) - 1) % r
    rhs = (ntt.poly_eval(r, ca, x) * ntt.poly_eval(r, cb, x) - ntt.poly_eval(r, cc, x)) % r
    assert lhs == rhs


def test_filter_semantics():
    """backend/groth16/bn254/utils_test.go:17-38 (filterHeap) - the only exact-value test on
    the reference's prove path; our K scalars are W[nb_public:] minus removed wires."""
    def filter_heap(slice_, first, to_remove):
        rm = set(to_remove)
        return [v for i, v in enumerate(slice_) if (i + first) not in rm]
    elems = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]
    assert filter_heap(elems, 2, [2, 3, 5]) == [3, 5, 6, 7, 8, 9, 10]   # indices 2,3,5 -> drop 1,2,4
    assert filter_heap(elems, 0, []) == elems


def test_plonk_quotient_identity():
    """A satisfying trace (sigma = identity => Z = 1) makes the numerator vanish on H, so the
    quotient h from numerator/divide_by_zh satisfies h(x)(x^n - 1) = N(x) at a random point
    (independent of the coset machinery)."""
    from oracle import plonk
    c = CURVES["bn254"]
    r = c.r
    n, rho = 8, 4
    rng = random.Random(12)
    dom0 = ntt.Domain(c, n)
    dom1 = ntt.Domain(c, rho * n)
    g = dom1.coset_gen
    w = [pow(dom0.generator, j, r) for j in range(n)]
    l = [rng.randrange(r) for _ in range(n)]
    rr = [rng.randrange(r) for _ in range(n)]
    qm = [rng.randrange(r) for _ in range(n)]
    ql = [rng.randrange(r) for _ in range(n)]
    qr = [rng.randrange(r) for _ in range(n)]
    qk = [rng.randrange(r) for _ in range(n)]
    qo = [r - 1] * n
    o = [(ql[j] * l[j] + qr[j] * rr[j] + qm[j] * l[j] * rr[j] + qk[j]) % r for j in range(n)]   # gate holds
    polys = {"l": l, "r": rr, "o": o, "z": [1] * n, "ql": ql, "qr": qr, "qm": qm, "qo": qo, "qk": qk,
             "s1": w, "s2": [g * x % r for x in w], "s3": [g * g * x % r for x in w]}
    alpha, beta, gamma = (rng.randrange(r) for _ in range(3))
    cres = plonk.numerator(c, n, rho, polys, alpha, beta, gamma, {})
    h = plonk.divide_by_zh(c, n, rho, cres)
    assert all(v == 0 for v in h[3 * n:])                       # deg h < 3n
    # independent evaluation of the numerator at a random point
    x = rng.randrange(r)
    can = {k: ntt.bit_reverse(dom0.fft_inverse(v, ntt.DIF)) for k, v in polys.items()}
    u = {k: ntt.poly_eval(r, can[k], x) for k in plonk.POLYS}
    u["zs"] = ntt.poly_eval(r, can["z"], x * dom0.generator % r)
    N = plonk.all_constraints(r, n, dom0.cardinality_inv, u, x, x * dom0.generator % r, alpha, beta, gamma, g, {},
                              (pow(x, n, r) - 1) % r)
    assert ntt.poly_eval(r, h, x) * (pow(x, n, r) - 1) % r == N


@pytest.mark.parametrize("cname", ("bn254", "bls12-381", "bw6-761"))
def test_plonk_prover_oracle_verifies(cname):
    """oracle/plonk_prover.py (restating backend/plonk/bn254/prove.go with injected challenges) produces
    proofs that satisfy the verifier's equations in the exponent; a tampered witness is rejected."""
    from oracle import plonk_prover as pp
    c = CURVES[cname]
    rng = random.Random(5)
    rnd = lambda: rng.randrange(c.r)
    for n in (8, 16):
        circ, l, rr, o = pp.random_satisfied_instance(c, n, seed=n)
        ch = pp.Challenges(gamma=rnd(), beta=rnd(), alpha=rnd(), zeta=rnd(), v=rnd(), bl=[rnd(), rnd()],
                           br=[rnd(), rnd()], bo=[rnd(), rnd()], bz=[rnd(), rnd(), rnd()])
        tau = rnd()
        pr = pp.prove(c, circ, l, rr, o, ch, tau)
        assert pp.verify(c, circ, pr, ch, tau)
        assert len(pr.lin_poly) == n + 3 and pr.z_lagrange[0] == 1
        bad = pp.prove(c, circ, l, rr, o, ch, tau)
        bad.claimed[1] = (bad.claimed[1] + 1) % c.r            # a wrong opened value
        assert not pp.verify(c, circ, bad, ch, tau)
        l2 = list(l)
        l2[1] = (l2[1] + 1) % c.r                                # an unsatisfied trace does not divide
        with pytest.raises(AssertionError):
            pp.prove(c, circ, l2, rr, o, ch, tau)


@pytest.mark.parametrize("cname", ["bn254", "bls12-381"])
def test_groth16_verifies_under_the_real_pairing(cname):
    """prove -> Verify as the reference tests it (test/assert_checkcircuit.go:140-144): the proof points of the
    trapdoor prover satisfy the pairing equation, a tampered proof or another public input does not."""
    c = CURVES[cname]
    cs, W = g16.cubic_r1cs(), g16.cubic_witness(c.r)
    pk = g16.setup_dlog(c, cs, g16.random_toxic(c, 3))
    pr = g16.prove_dlog(c, cs, pk, W, 1234567, 7654321)
    F1, F2 = ff.Fp(c.p), ff.base_field(c, 2)
    pts = (ec.scalar_mul(F1, pr.ar, c.g1), ec.scalar_mul(F2, pr.bs, c.g2), ec.scalar_mul(F1, pr.krs, c.g1))
    assert g16.verify_pairing(c, pk, *pts, W)
    assert not g16.verify_pairing(c, pk, pts[0], pts[1], ec.scalar_mul(F1, pr.krs + 1, c.g1), W)
    W2 = list(W); W2[1] = (W2[1] + 1) % c.r        # another public input
    assert not g16.verify_pairing(c, pk, *pts, W2)


@pytest.mark.parametrize("cname", ["bn254", "bls12-381"])
def test_pairing_oracle_self_checks(cname):
    """oracle/pairing.py (reduced Tate): non-degenerate, of order r, bilinear in both arguments; on BLS12-381 every
    product check agrees with the independent ate implementation (oracle/pairing_bls12_381.py)."""
    from oracle import pairing
    c = CURVES[cname]
    T = pairing.get(c)
    F1, F2 = ff.Fp(c.p), ff.base_field(c, 2)
    e = T.pairing(c.g1, c.g2)
    assert e != T.K.one and T.K.pow(e, c.r) == T.K.one
    a, b = 0xDEADBEEF, 0xC0FFEE123
    assert T.pairing(ec.scalar_mul(F1, a, c.g1), ec.scalar_mul(F2, b, c.g2)) == T.K.pow(e, a * b % c.r)
    good = [(ec.scalar_mul(F1, a, c.g1), ec.scalar_mul(F2, b, c.g2)), (ec.affine_neg(F1, c.g1), ec.scalar_mul(F2, a * b % c.r, c.g2))]
    bad = [good[0], (good[1][0], ec.scalar_mul(F2, (a * b + 1) % c.r, c.g2))]
    assert T.product_is_one(good) and not T.product_is_one(bad)
    if cname == "bls12-381":
        from oracle import pairing_bls12_381 as ate
        assert ate.pairing_product_is_one(good) and not ate.pairing_product_is_one(bad)
