"""PLONK prover restated on big ints, with a trapdoor SRS.  TEST INFRASTRUCTURE ONLY
(see oracle/params.py header for what pins parity and what does not).

Follows backend/plonk/bn254/prove.go (StatisticalZK off; BSB22 commitment gates :867-884 with the committed
polynomials PI2_i GIVEN - in the reference they come from the solver hint :280-318 - their digests
Bsb22Commitments :300, the linearised-polynomial term sum_i Qcp_i(zeta) PI2_i(X) :1362,1457 and the Qcp openings
:805-817):
  Prove :98-153; blinding polynomials :259-266,1239-1253 (orders 1,1,1,2 :72-75);
  commitToLRO :404-489 + commitBlindingFactor :1223-1236 (digest of p + b*(X^n - 1));
  buildRatioCopyConstraint :635-667; computeQuotient :558-633 -> computeNumerator :841-1123,
  divideByZH :1287-1324, h split h1,h2,h3 of n+2 coefficients :689-722, commitToQuotient :1263-1282;
  openZ :670-687 (blinded Z at w*zeta); computeLinearizedPolynomial :724-794 +
  innerComputeLinearizedPoly :1366-1487; batchOpening :796-837.

Randomness (blinding coefficients) and the Fiat-Shamir challenges (gamma, beta, alpha, zeta and the
KZG folding challenge) are INJECTED: the transcript encoding lives in gnark-crypto (absent), and
injected values make every intermediate result comparable bit for bit.
With a known tau every KZG digest is p(tau) * G, so digests are carried as discrete logs and the
verifier's pairing checks become equalities in Fr (test/unsafekzg pattern).
The batched opening folds with powers of the challenge: f = sum_i v^i p_i (kzg.BatchOpenSinglePoint
semantics as relied upon at :827-834).
"""

from dataclasses import dataclass, field
from typing import List

from . import plonk
from .ntt import DIF, Domain, bit_reverse, poly_eval


@dataclass
class Circuit:
    """Lagrange-form (regular layout) trace on domain0 and the wire permutation
    (Trace / NewTrace, backend/plonk/bn254/setup.go:67-231,289-392)."""
    n: int
    ql: List[int]
    qr: List[int]
    qm: List[int]
    qo: List[int]
    qk: List[int]           # complete Qk (public inputs folded in, prove.go:349-373)
    perm: List[int]         # 3n entries
    qcp: List[List[int]] = field(default_factory=list)    # trace.Qcp: one selector per BSB22 commitment


@dataclass
class Challenges:
    gamma: int
    beta: int
    alpha: int
    zeta: int
    v: int                  # KZG folding challenge
    bl: List[int] = field(default_factory=lambda: [0, 0])
    br: List[int] = field(default_factory=lambda: [0, 0])
    bo: List[int] = field(default_factory=lambda: [0, 0])
    bz: List[int] = field(default_factory=lambda: [0, 0, 0])
    # StatisticalZK (backend.WithStatisticalZeroKnowledge, prove.go:239-242): the two quotientShardsRandomizers, or
    # None.  h1 += b1 X^(n+2);  h2 += -b1 + b2 X^(n+2);  h3 += -b2   (prove.go:689-722) - the quotient itself,
    # h1 + X^(n+2) h2 + X^(2(n+2)) h3, does not change, its three commitments and the linearised polynomial do.
    hr: List[int] = None


@dataclass
class Proof:
    # digests as discrete logs (p(tau))
    L: int
    R: int
    O: int
    Z: int
    H: List[int]
    lin: int
    batch_opening: int       # [ (f - f(zeta)) / (X - zeta) ]
    z_opening: int           # [ (Z_b - Z_b(w zeta)) / (X - w zeta) ]
    claimed: List[int]       # lin(zeta), l(zeta), r(zeta), o(zeta), s1(zeta), s2(zeta), then Qcp_j(zeta)
    zu: int                  # Z_b(w zeta)
    # intermediates kept for parity tests
    h: List[int] = field(default_factory=list)
    lin_poly: List[int] = field(default_factory=list)
    z_lagrange: List[int] = field(default_factory=list)
    bsb22: List[int] = field(default_factory=list)       # Bsb22Commitments: [PI2_j]


def canonical(curve, dom: Domain, lagrange):
    return bit_reverse(dom.fft_inverse(lagrange, DIF))


def blinded(r, n, coeffs, b):
    """getBlindedCoefficients :1211-1220: p + b*(X^n - 1)."""
    out = list(coeffs) + list(b)
    for i, bi in enumerate(b):
        out[i] = (out[i] - bi) % r
    return out


def sigma_polys(curve, dom0: Domain, perm):
    supp = plonk.support_permutation(curve, dom0)
    n = dom0.n
    return [[supp[perm[j * n + i]] for i in range(n)] for j in range(3)]


def prove(curve, circ: Circuit, l, rr, o, ch: Challenges, tau: int, pi2=()) -> Proof:
    """pi2: the committed polynomials (Lagrange/regular), one per circ.qcp entry"""
    r, n = curve.r, circ.n
    assert len(pi2) == len(circ.qcp)
    dom0 = Domain(curve, n)
    g, w = dom0.coset_gen, dom0.generator
    ev = lambda p, x: poly_eval(r, p, x)
    s1, s2, s3 = sigma_polys(curve, dom0, circ.perm)
    # --- L, R, O
    cl, cr, co = (canonical(curve, dom0, v) for v in (l, rr, o))
    lb, rb, ob = blinded(r, n, cl, ch.bl), blinded(r, n, cr, ch.br), blinded(r, n, co, ch.bo)
    # --- Z
    z = plonk.build_ratio_copy_constraint(curve, dom0, l, rr, o, circ.perm, ch.beta, ch.gamma)
    zb = blinded(r, n, canonical(curve, dom0, z), ch.bz)
    # --- quotient
    polys = {"l": l, "r": rr, "o": o, "z": z, "s1": s1, "s2": s2, "s3": s3,
             "ql": circ.ql, "qr": circ.qr, "qm": circ.qm, "qo": circ.qo, "qk": circ.qk}
    blind = {"l": ch.bl, "r": ch.br, "o": ch.bo, "z": ch.bz}
    cres = plonk.numerator(curve, n, 4, polys, ch.alpha, ch.beta, ch.gamma, blind, bsb22=list(zip(circ.qcp, pi2)))
    h = plonk.divide_by_zh(curve, n, 4, cres)
    assert all(x == 0 for x in h[3 * (n + 2):]), "numerator not divisible by X^n - 1: unsatisfied trace"
    h1, h2, h3 = h[:n + 2], h[n + 2:2 * (n + 2)], h[2 * (n + 2):3 * (n + 2)]
    if ch.hr is not None:                   # h1(), h2(), h3() with StatisticalZK, prove.go:689-722
        b1, b2 = ch.hr
        h1 = h1 + [b1 % r]
        h2 = [(h2[0] - b1) % r] + h2[1:] + [b2 % r]
        h3 = [(h3[0] - b2) % r] + h3[1:]
    # --- openings at zeta
    zeta = ch.zeta
    zu = ev(zb, zeta * w % r)
    lz, rz, oz = ev(lb, zeta), ev(rb, zeta), ev(ob, zeta)
    cs1, cs2, cs3 = (canonical(curve, dom0, v) for v in (s1, s2, s3))
    cql, cqr, cqm, cqo, cqk = (canonical(curve, dom0, v) for v in (circ.ql, circ.qr, circ.qm, circ.qo, circ.qk))
    s1z, s2z = ev(cs1, zeta), ev(cs2, zeta)
    cqcp = [canonical(curve, dom0, v) for v in circ.qcp]
    cpi2 = [canonical(curve, dom0, v) for v in pi2]
    qcpz = [ev(p, zeta) for p in cqcp]
    # innerComputeLinearizedPoly :1366-1487
    alpha, beta, gamma = ch.alpha, ch.beta, ch.gamma
    rl = rz * lz % r
    c1 = (lz + beta * s1z + gamma) % r * ((rz + beta * s2z + gamma) % r) % r * zu % r * beta % r * alpha % r
    uz, uuz = zeta * g % r, zeta * g % r * g % r
    c2 = (lz + beta * zeta + gamma) % r * ((rz + beta * uz + gamma) % r) % r * ((oz + beta * uuz + gamma) % r) % r
    c2 = (-c2 * alpha) % r
    zn = pow(zeta, n, r)
    zn2 = zn * zeta % r * zeta % r
    zh = (zn - 1) % r
    a2l1 = zh * pow((zeta - 1) % r, -1, r) % r * alpha % r * alpha % r * dom0.cardinality_inv % r
    lin = []
    for i in range(len(zb)):
        t = zb[i] * c2 % r
        if i < n:
            t = (t + cs3[i] * c1 + cqm[i] * rl + cql[i] * lz + cqr[i] * rz + cqo[i] * oz + cqk[i]) % r
            for j in range(len(cqcp)):              # + sum_j Qcp_j(zeta) PI2_j(X)   (:1457-1460)
                t = (t + cpi2[j][i] * qcpz[j]) % r
        t = (t + zb[i] * a2l1) % r
        if i < len(h3):
            t = (t - zh * ((h3[i] * zn2 + h2[i]) % r * zn2 % r + h1[i])) % r
        elif ch.hr is not None:             # prove.go:1476-1481: h1, h2 are one coefficient longer than h3
            t = (t - zh * ((h2[i] * zn2 + h1[i]) % r)) % r
        lin.append(t)
    # batchOpening :796-837
    to_open = [lin, lb, rb, ob, cs1, cs2] + cqcp
    claimed = [ev(p, zeta) for p in to_open]
    size = max(len(p) for p in to_open)
    f = [0] * size
    vp = 1
    for p in to_open:
        for i, c in enumerate(p):
            f[i] = (f[i] + vp * c) % r
        vp = vp * ch.v % r
    qf, _ = plonk.div_by_linear(r, f, zeta)
    qz, zu2 = plonk.div_by_linear(r, zb, zeta * w % r)
    assert zu2 == zu
    com = lambda p: ev(p, tau)
    return Proof(L=com(lb), R=com(rb), O=com(ob), Z=com(zb), H=[com(h1), com(h2), com(h3)], lin=com(lin),
                 batch_opening=com(qf), z_opening=com(qz), claimed=claimed, zu=zu, h=h, lin_poly=lin, z_lagrange=z,
                 bsb22=[com(p) for p in cpi2])


def verify(curve, circ: Circuit, proof: Proof, ch: Challenges, tau: int) -> bool:
    """The verifier's checks (backend/plonk/bn254/verify.go:38-320) in the exponent: the linearised
    identity at zeta and the two KZG openings, with every digest replaced by its discrete log."""
    r, n = curve.r, circ.n
    dom0 = Domain(curve, n)
    g, w = dom0.coset_gen, dom0.generator
    zeta, alpha, beta, gamma = ch.zeta, ch.alpha, ch.beta, ch.gamma
    lin_z, lz, rz, oz, s1z, s2z = proof.claimed[:6]
    qcpz = proof.claimed[6:]
    assert len(qcpz) == len(circ.qcp) == len(proof.bsb22)
    zn = pow(zeta, n, r)
    zh = (zn - 1) % r
    l1 = zh * pow((zeta - 1) % r, -1, r) % r * dom0.cardinality_inv % r
    # lin(zeta) = alpha^2 L1(zeta) - alpha (l + beta s1 + gamma)(r + beta s2 + gamma)(o + gamma) zu
    want = (alpha * alpha % r * l1
            - alpha * ((lz + beta * s1z + gamma) % r) % r * ((rz + beta * s2z + gamma) % r) % r * ((oz + gamma) % r) % r * proof.zu) % r
    if lin_z != want:
        return False
    # the verifier rebuilds [lin] from the verifying key digests and the claimed values
    s1, s2, s3 = sigma_polys(curve, dom0, circ.perm)
    com_l = lambda lag: poly_eval(r, canonical(curve, dom0, lag), tau)
    c1 = (lz + beta * s1z + gamma) % r * ((rz + beta * s2z + gamma) % r) % r * proof.zu % r * beta % r * alpha % r
    uz, uuz = zeta * g % r, zeta * g % r * g % r
    c2 = (-(lz + beta * zeta + gamma) % r * ((rz + beta * uz + gamma) % r) % r * ((oz + beta * uuz + gamma) % r) % r * alpha) % r
    zn2 = zn * zeta % r * zeta % r
    lin_digest = (proof.Z * ((c2 + alpha * alpha % r * l1) % r) + com_l(s3) * c1 + com_l(circ.qm) * (rz * lz % r)
                  + com_l(circ.ql) * lz + com_l(circ.qr) * rz + com_l(circ.qo) * oz + com_l(circ.qk)
                  + sum(q_ * d_ for q_, d_ in zip(qcpz, proof.bsb22))       # sum_j Qcp_j(zeta) [PI2_j]  (verify.go)
                  - zh * ((proof.H[2] * zn2 + proof.H[1]) % r * zn2 % r + proof.H[0])) % r
    if lin_digest != proof.lin:
        return False
    # batched KZG opening at zeta: e([f] - f(zeta)[1], [1]) = e([H], [tau - zeta])
    digests = [proof.lin, proof.L, proof.R, proof.O, com_l(s1), com_l(s2)] + [com_l(q_) for q_ in circ.qcp]
    f_tau = f_z = 0
    vp = 1
    for d, c in zip(digests, proof.claimed):
        f_tau = (f_tau + vp * d) % r
        f_z = (f_z + vp * c) % r
        vp = vp * ch.v % r
    if proof.batch_opening * ((tau - zeta) % r) % r != (f_tau - f_z) % r:
        return False
    # opening of Z at w*zeta
    if proof.z_opening * ((tau - zeta * w) % r) % r != (proof.Z - proof.zu) % r:
        return False
    return True


def verify_pairing(curve, circ: Circuit, points, values, ch: Challenges, tau: int = None, bsb22_points=(), srs_g1=None,
                   tau_g2=None, msm=None) -> bool:
    """The verifier of backend/plonk/bn254/verify.go:38-320 on the PROOF POINTS, with real pairings (oracle/pairing.py;
    BN254 and BLS12-381): what the reference's own test does with a proof (prove -> Verify).
      points: the ten G1 points a prover returns - [L] [R] [O] [Z] [H1] [H2] [H3] [linearised] [batch opening] [Z opening]
              (affine, canonical ints);  values: lin(zeta), l, r, o, s1, s2 at zeta, Z(w zeta), then Qcp_j(zeta).
    The verifying key (digests of the selectors / permutation polynomials, [1]_2, [tau]_2) is derived from the circuit
    as Setup would: from tau when the trapdoor is known, or - no trapdoor anywhere - from an SRS given as points
    (srs_g1 = [tau^k]_1 affine, tau_g2 = [tau]_2; e.g. the Ethereum KZG ceremony SRS the reference ships).  The proof
    side uses nothing but the points and values handed in."""
    from . import ec, ff, pairing
    r, n = curve.r, circ.n
    F1, F2 = ff.Fp(curve.p), ff.base_field(curve, 2)
    G1, G2 = curve.g1, curve.g2
    dom0 = Domain(curve, n)
    assert len(values) - 7 == len(circ.qcp) == len(bsb22_points)
    mul = lambda k, P: ec.scalar_mul(F1, k % r, P)
    # verifying key
    s1, s2, s3 = sigma_polys(curve, dom0, circ.perm)
    if srs_g1 is not None:
        assert tau_g2 is not None and len(srs_g1) >= n
        msm = msm or (lambda pts, sc: ec.msm_naive(F1, pts, sc))     # msm: optional faster MSM for large n
        vk = lambda lag: msm(list(srs_g1[:n]), canonical(curve, dom0, lag))
    else:
        vk = lambda lag: mul(poly_eval(r, canonical(curve, dom0, lag), tau), G1)
        tau_g2 = ec.scalar_mul(F2, tau, G2)
    key = {"S1": vk(s1), "S2": vk(s2), "S3": vk(s3), "Ql": vk(circ.ql), "Qr": vk(circ.qr), "Qm": vk(circ.qm),
           "Qo": vk(circ.qo), "Qk": vk(circ.qk), "Qcp": [vk(q_) for q_ in circ.qcp]}
    T = pairing.get(curve)
    opening = lambda lhs, H: T.product_is_one([(lhs, G2), (ec.affine_neg(F1, H), tau_g2)])
    return verify_core(curve, n, key, points, values, ch, opening, bsb22_points)


def verify_core(curve, n, key, points, values, ch: Challenges, opening, bsb22_points=()) -> bool:
    """The verifier proper (backend/plonk/bn254/verify.go:38-320) from a verifying KEY given as points
    (key: S1 S2 S3 Ql Qr Qm Qo Qk as affine G1 points, Qcp a list) - what Setup hands a verifier - so that it can run at
    sizes where deriving the key from the circuit in Python is not affordable (oracle/plonk_fast.py derives it with the
    C++ oracle).  opening(lhs, H) -> bool decides one KZG opening equation  e(lhs, [1]_2) == e(H, [tau]_2):
    with real pairings (verify_pairing) or, when the trapdoor is known, as lhs == tau * H."""
    from . import ec, ff
    r = curve.r
    F1 = ff.Fp(curve.p)
    G1 = curve.g1
    dom0 = Domain(curve, n)
    g, w = dom0.coset_gen, dom0.generator
    zeta, alpha, beta, gamma = ch.zeta, ch.alpha, ch.beta, ch.gamma
    cmL, cmR, cmO, cmZ, cmH1, cmH2, cmH3, cmLin, cmBatch, cmZopen = points
    lin_z, lz, rz, oz, s1z, s2z, zu = values[:7]
    qcpz = list(values[7:])
    mul = lambda k, P: ec.scalar_mul(F1, k % r, P)
    add = lambda P, Q: ec.affine_add(F1, P, Q)
    neg = lambda P: ec.affine_neg(F1, P)
    vS1, vS2, vS3 = key["S1"], key["S2"], key["S3"]
    vQl, vQr, vQm, vQo, vQk = key["Ql"], key["Qr"], key["Qm"], key["Qo"], key["Qk"]
    vQcp = list(key.get("Qcp", ()))
    # 1. the opened value of the linearised polynomial (verify.go: the constant part moved to the right-hand side)
    zn = pow(zeta, n, r)
    zh = (zn - 1) % r
    l1 = zh * pow((zeta - 1) % r, -1, r) % r * dom0.cardinality_inv % r
    want = (alpha * alpha % r * l1
            - alpha * ((lz + beta * s1z + gamma) % r) % r * ((rz + beta * s2z + gamma) % r) % r * ((oz + gamma) % r) % r * zu) % r
    if lin_z != want:
        return False
    # 2. the linearised digest, rebuilt from the key and the proof
    c1 = (lz + beta * s1z + gamma) % r * ((rz + beta * s2z + gamma) % r) % r * zu % r * beta % r * alpha % r
    uz, uuz = zeta * g % r, zeta * g % r * g % r
    c2 = (-(lz + beta * zeta + gamma) % r * ((rz + beta * uz + gamma) % r) % r * ((oz + beta * uuz + gamma) % r) % r * alpha) % r
    zn2 = zn * zeta % r * zeta % r
    acc = mul((c2 + alpha * alpha % r * l1) % r, cmZ)
    for k, P in ((c1, vS3), (rz * lz, vQm), (lz, vQl), (rz, vQr), (oz, vQo), (1, vQk)):
        acc = add(acc, mul(k, P))
    for q_, P in zip(qcpz, bsb22_points):
        acc = add(acc, mul(q_, P))
    hfold = add(add(mul(zn2 * zn2, cmH3), mul(zn2, cmH2)), cmH1)
    acc = add(acc, neg(mul(zh, hfold)))
    if acc != cmLin:
        return False
    # 3. the two KZG openings: e([f] - f(z)[1] + z [H], [1]_2) = e([H], [tau]_2)
    digests = [cmLin, cmL, cmR, cmO, vS1, vS2] + vQcp
    claimed = [lin_z, lz, rz, oz, s1z, s2z] + qcpz
    Fd, fz, vp = None, 0, 1
    for D, cval in zip(digests, claimed):
        Fd = add(Fd, mul(vp, D))
        fz = (fz + vp * cval) % r
        vp = vp * ch.v % r
    lhs = add(add(Fd, neg(mul(fz, G1))), mul(zeta, cmBatch))
    if not opening(lhs, cmBatch):
        return False
    lhs = add(add(cmZ, neg(mul(zu, G1))), mul(zeta * w, cmZopen))
    return opening(lhs, cmZopen)


def random_satisfied_instance(curve, n, seed, n_commit=0):
    """A random satisfied trace: random gates L*R-style with O solved, and a permutation built from
    cycles over slots that are FORCED to carry equal values (so the copy constraints hold).
    n_commit > 0: also n_commit BSB22 gates (random selectors Qcp_j, random committed polynomials PI2_j entering
    the gate equation); returns (circuit, l, r, o, pi2)."""
    import random
    rng = random.Random(seed)
    r = curve.r
    # pick wire values with repeated values so that non-trivial cycles exist
    pool = [rng.randrange(r) for _ in range(max(4, n // 2))]
    l = [rng.choice(pool) for _ in range(n)]
    rr = [rng.choice(pool) for _ in range(n)]
    ql = [rng.randrange(r) for _ in range(n)]
    qr = [rng.randrange(r) for _ in range(n)]
    qm = [rng.randrange(r) for _ in range(n)]
    qk = [rng.randrange(r) for _ in range(n)]
    qo = [r - 1] * n
    qcp = [[rng.randrange(r) if rng.random() < 0.5 else 0 for _ in range(n)] for _ in range(n_commit)]
    pi2 = [[rng.randrange(r) for _ in range(n)] for _ in range(n_commit)]
    o = [(ql[i] * l[i] + qr[i] * rr[i] + qm[i] * l[i] * rr[i] + qk[i]
          + sum(qcp[j][i] * pi2[j][i] for j in range(n_commit))) % r for i in range(n)]
    vals = l + rr + o
    # permutation: link all slots holding the same value into one cycle
    groups = {}
    for idx, v in enumerate(vals):
        groups.setdefault(v, []).append(idx)
    perm = list(range(3 * n))
    for idxs in groups.values():
        if len(idxs) > 1:
            rng.shuffle(idxs)
            for a, b in zip(idxs, idxs[1:] + idxs[:1]):
                perm[a] = b
    circ = Circuit(n=n, ql=ql, qr=qr, qm=qm, qo=qo, qk=qk, perm=perm, qcp=qcp)
    if n_commit:
        return circ, l, rr, o, pi2
    return circ, l, rr, o
