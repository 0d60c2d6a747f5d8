"""Groth16 trapdoor Setup / Prove restated on big ints.  TEST INFRASTRUCTURE ONLY
(see oracle/params.py header for what pins parity and what does not).

Follows:
  Setup      backend/groth16/bn254/setup.go:75-331 (scalar layout, InfinityA/B
             filtering :194-219, bit-reversed Z :247-249), setupABC :346-428.
  Prove      backend/groth16/bn254/prove.go:52-315 (wire filtering :147-168,
             deltas :185, Ar :207-214, Bs1 :194-200, Krs :227-269, Bs2 :283-292).
  computeH   backend/groth16/bn254/prove.go:346-389.
No BSB22 commitments (commitmentInfo empty): rows a8 of SURVEY.md §8 are handled
by the plain MSM entry point.

Because the toxic waste is known here, every proving-key point has a known
discrete log, so every MSM output and the Groth16 relation
  Ar*Bs = alpha*beta + (sum_pub vkK_i w_i)*gamma + Krs*delta
can be checked in Fr without a pairing.
"""

import random
from dataclasses import dataclass, field
from typing import List, Tuple

from . import ec, ff
from .ntt import DIF, DIT, Domain, bit_reverse


@dataclass
class R1CS:
    """constraints: list of (L, R, O); each a list of (coeff, wire_id).
    Wire order as in gnark: public (wire 0 = constant one), secret, internal."""
    nb_public: int
    nb_secret: int
    nb_internal: int
    constraints: List[Tuple[list, list, list]]

    @property
    def nb_wires(self):
        return self.nb_public + self.nb_secret + self.nb_internal

    @property
    def nb_constraints(self):
        return len(self.constraints)


def cubic_r1cs():
    """examples/cubic/cubic.go:12-25: x^3 + x + 5 == y.
    wires: 0=one, 1=y (public) | 2=x (secret) | 3=x^2, 4=x^3 (internal)."""
    cons = [
        ([(1, 2)], [(1, 2)], [(1, 3)]),                       # x*x = x2
        ([(1, 3)], [(1, 2)], [(1, 4)]),                       # x2*x = x3
        ([(1, 4), (1, 2), (5, 0)], [(1, 0)], [(1, 1)]),       # (x3+x+5)*1 = y
    ]
    return R1CS(nb_public=2, nb_secret=1, nb_internal=2, constraints=cons)


def cubic_witness(r, x=3):
    y = (x ** 3 + x + 5) % r
    return [1, y, x % r, x * x % r, x ** 3 % r]


def square_chain_r1cs(nb_constraints: int):
    """x_{i+1} = x_i^2 chain, the shape of backend/groth16/groth16_test.go:126-156.
    wires: 0=one, 1=y (public) | 2=x0 (secret) | internal x1..x_{m-1}; last square = y."""
    m = nb_constraints
    cons = []
    # wire ids: x0 = 2, x_k = 2 + k for k < m ; x_m = y = wire 1
    for k in range(m):
        src = 2 + k
        dst = 1 if k == m - 1 else 3 + k
        cons.append(([(1, src)], [(1, src)], [(1, dst)]))
    return R1CS(nb_public=2, nb_secret=1, nb_internal=m - 1, constraints=cons)


def square_chain_witness(r, nb_constraints: int, x0=3):
    xs = [x0 % r]
    for _ in range(nb_constraints):
        xs.append(xs[-1] * xs[-1] % r)
    return [1, xs[-1]] + xs[:-1]


def solve_abc(r1cs: R1CS, W, r):
    """R1CSSolution{A,B,C}: per-constraint evaluations <L,W>, <R,W>, <O,W>
    (constraint/bn254/system.go:162-165)."""
    def ev(lin):
        return sum(c * W[w] for c, w in lin) % r
    A = [ev(L) for L, _, _ in r1cs.constraints]
    B = [ev(R) for _, R, _ in r1cs.constraints]
    C = [ev(O) for _, _, O in r1cs.constraints]
    return A, B, C


@dataclass
class Toxic:
    t: int
    alpha: int
    beta: int
    gamma: int
    delta: int


@dataclass
class ProvingKeyDlog:
    """Discrete logs (w.r.t. g1 / g2) of every proving-key element, in the
    ProvingKey layout of backend/groth16/bn254/setup.go:25-48."""
    domain: Domain
    alpha: int
    beta: int
    delta: int
    A: List[int]            # filtered (no zeros)
    B: List[int]            # filtered
    Z: List[int]            # n-1 entries, bit-reversed order
    K: List[int]            # private wires
    infinity_a: List[bool]
    infinity_b: List[bool]
    vk_K: List[int]
    gamma: int


def setup_dlog(curve, r1cs: R1CS, toxic: Toxic) -> ProvingKeyDlog:
    r = curve.r
    dom = Domain(curve, r1cs.nb_constraints)
    n = dom.n
    t = toxic.t
    nw = r1cs.nb_wires
    A = [0] * nw
    B = [0] * nw
    C = [0] * nw
    # setupABC (setup.go:346-428): L_j(t), j over constraints
    w = dom.generator
    tn1 = (pow(t, n, r) - 1) % r
    L = tn1 * pow((t - 1) % r, -1, r) % r * dom.cardinality_inv % r
    wi = 1
    for j, (cl, cr, co) in enumerate(r1cs.constraints):
        for c, wid in cl:
            A[wid] = (A[wid] + c * L) % r
        for c, wid in cr:
            B[wid] = (B[wid] + c * L) % r
        for c, wid in co:
            C[wid] = (C[wid] + c * L) % r
        # L_{j+1} = w * L_j * (t - w^j) / (t - w^(j+1))
        wn = wi * w % r
        L = L * w % r * ((t - wi) % r) % r * pow((t - wn) % r, -1, r) % r
        wi = wn
    gamma_inv = pow(toxic.gamma, -1, r)
    delta_inv = pow(toxic.delta, -1, r)
    vkK, pkK = [], []
    for i in range(nw):
        k = (toxic.beta * A[i] + toxic.alpha * B[i] + C[i]) % r
        if i < r1cs.nb_public:
            vkK.append(k * gamma_inv % r)
        else:
            pkK.append(k * delta_inv % r)
    zdt = tn1 * delta_inv % r
    Z = []
    for _ in range(n):
        Z.append(zdt)
        zdt = zdt * t % r
    Z = bit_reverse(Z)[:n - 1]
    inf_a = [a == 0 for a in A]
    inf_b = [b == 0 for b in B]
    return ProvingKeyDlog(domain=dom, alpha=toxic.alpha, beta=toxic.beta, delta=toxic.delta,
                          A=[a for a in A if a], B=[b for b in B if b], Z=Z, K=pkK,
                          infinity_a=inf_a, infinity_b=inf_b, vk_K=vkK, gamma=toxic.gamma)


def compute_h(dom: Domain, a, b, c):
    """prove.go:346-389; returns h in bit-reversed order (length n)."""
    r = dom.r
    n = dom.n
    pad = [0] * (n - len(a))
    a = list(a) + pad
    b = list(b) + pad
    c = list(c) + pad
    a = dom.fft_inverse(a, DIF)
    b = dom.fft_inverse(b, DIF)
    c = dom.fft_inverse(c, DIF)
    a = dom.fft(a, DIT, on_coset=True)
    b = dom.fft(b, DIT, on_coset=True)
    c = dom.fft(c, DIT, on_coset=True)
    den = pow((pow(dom.coset_gen, n, r) - 1) % r, -1, r)
    a = [((x * y - z) % r) * den % r for x, y, z in zip(a, b, c)]
    return dom.fft_inverse(a, DIF, on_coset=True)


@dataclass
class ProofDlog:
    ar: int
    bs: int
    krs: int
    # the five raw MSM results (dlogs), in prove.go order
    msm_a: int
    msm_b: int
    msm_z: int
    msm_k: int
    h: list = field(default_factory=list)
    wire_values_a: list = field(default_factory=list)
    wire_values_b: list = field(default_factory=list)
    wire_values_k: list = field(default_factory=list)


def prove_dlog(curve, r1cs: R1CS, pk: ProvingKeyDlog, W, rr: int, ss: int) -> ProofDlog:
    r = curve.r
    A, B, C = solve_abc(r1cs, W, r)
    h = compute_h(pk.domain, A, B, C)
    n = pk.domain.n
    wa = [W[i] for i in range(len(W)) if not pk.infinity_a[i]]
    wb = [W[i] for i in range(len(W)) if not pk.infinity_b[i]]
    wk = list(W[r1cs.nb_public:])
    msm_a = sum(x * y for x, y in zip(pk.A, wa)) % r
    msm_b = sum(x * y for x, y in zip(pk.B, wb)) % r
    msm_z = sum(x * y for x, y in zip(pk.Z, h[:n - 1])) % r
    msm_k = sum(x * y for x, y in zip(pk.K, wk)) % r
    ar = (msm_a + pk.alpha + rr * pk.delta) % r
    bs = (msm_b + pk.beta + ss * pk.delta) % r
    krs = (msm_k + msm_z + (-rr * ss % r) * pk.delta + ss * ar + rr * bs) % r
    return ProofDlog(ar=ar, bs=bs, krs=krs, msm_a=msm_a, msm_b=msm_b, msm_z=msm_z, msm_k=msm_k,
                     h=h, wire_values_a=wa, wire_values_b=wb, wire_values_k=wk)


def verify_dlog(curve, r1cs: R1CS, pk: ProvingKeyDlog, proof: ProofDlog, W) -> bool:
    """Groth16 pairing equation in the exponent (backend/groth16/bn254/verify.go:38-140)."""
    r = curve.r
    pub = sum(k * W[i] for i, k in enumerate(pk.vk_K)) % r
    lhs = proof.ar * proof.bs % r
    rhs = (pk.alpha * pk.beta + pub * pk.gamma + proof.krs * pk.delta) % r
    return lhs == rhs


def verify_pairing(curve, pk: ProvingKeyDlog, ar_pt, bs_pt, krs_pt, W) -> bool:
    """The reference's acceptance test for a proof (backend/groth16/bls12-381/verify.go:38-140, reached from
    test/assert_checkcircuit.go:140-144): e(Ar, Bs) = e(alpha, beta) * e(sum_i w_i K_i, gamma) * e(Krs, delta), on the
    PROOF POINTS themselves with a real pairing (oracle/pairing.py: BN254 and BLS12-381, the curves whose G1 / G2
    generators the oracle holds).  The verifying key is derived from the trapdoor: alpha G1, beta G2, gamma G2,
    delta G2, K_i G1."""
    from . import ec, ff, pairing
    assert curve.name in pairing.TOWER
    r = curve.r
    F1, F2 = ff.Fp(curve.p), ff.base_field(curve, 2)
    pub = sum(k * W[i] for i, k in enumerate(pk.vk_K)) % r
    neg = lambda P: ec.affine_neg(F1, P)
    return pairing.get(curve).product_is_one([
        (ar_pt, bs_pt),
        (neg(ec.scalar_mul(F1, pk.alpha, curve.g1)), ec.scalar_mul(F2, pk.beta, curve.g2)),
        (neg(ec.scalar_mul(F1, pub, curve.g1)), ec.scalar_mul(F2, pk.gamma, curve.g2)),
        (neg(krs_pt), ec.scalar_mul(F2, pk.delta, curve.g2)),
    ])


def random_toxic(curve, seed: int) -> Toxic:
    rng = random.Random(seed)
    r = curve.r
    return Toxic(*[rng.randrange(2, r) for _ in range(5)])
