"""Full-size PLONK fixtures and the verifier at sizes where Python lists of big ints are not affordable
(BASELINE configs[3]: BLS12-381, 2^22 gates).  TEST INFRASTRUCTURE ONLY (see oracle/params.py header): imported by
tests/ and by bench.py to BUILD the workload and to CHECK a proof, never by the product.

  satisfied_instance  a random SATISFIED trace in gnark's memory layout (Montgomery limb arrays): general gates
                      ql l + qr r + qm l r - o + qk = 0 with random selectors, and a wire permutation with n/4 L-R
                      copy constraints and n/8 O-L copy constraints (2-cycles) whose slots are forced equal
                      (trace layout: backend/plonk/bn254/setup.go:67-231, permutation :289-392)
  trapdoor_srs_*      [tau^i] G1, i < n + 3 (test/unsafekzg/kzgsrs.go:142-172)
  verifying_key       the eight key digests as q(tau) G with q(tau) = sum_i q_i L_i(tau) (Lagrange form, C++ oracle)
  verify              backend/plonk/bn254/verify.go:38-320 through oracle/plonk_prover.verify_core on the ten proof
                      points and the opened values; the two KZG openings are checked twice: with the trapdoor
                      (lhs == tau H, a G1 identity) and, on the curves with an oracle pairing, with real pairings.
All vector arithmetic runs in the C++ oracle (corelib.fr_vec / fr_dot / fr_lagrange_at / fr_geometric); the group
arithmetic of the verifier (about 30 scalar multiplications) is the big-int oracle's.
"""

from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import corelib, ec, ff
from .plonk_prover import Challenges, verify_core


@dataclass
class Instance:
    curve: object
    log2n: int
    n: int
    ql: np.ndarray
    qr: np.ndarray
    qm: np.ndarray
    qo: np.ndarray
    qk: np.ndarray
    perm: np.ndarray          # int64, 3n
    l: np.ndarray
    r: np.ndarray
    o: np.ndarray
    ch: Challenges
    tau: int
    _vk: dict = field(default=None, repr=False)

    def challenges_packed(self):
        """(gamma, beta, alpha, zeta, v, bl, br, bo, bz) as Montgomery limb arrays - lib.PlonkKey.prove's order"""
        c, ch = self.curve, self.ch
        pe = lambda v: ff.pack_elements(v, c.r, c.fr_limbs)
        return (pe([ch.gamma]), pe([ch.beta]), pe([ch.alpha]), pe([ch.zeta]), pe([ch.v]), pe(ch.bl), pe(ch.br), pe(ch.bo),
                pe(ch.bz))


def _rand_fr(rs, curve, count):
    """count uniform residues below 2^(bits(r) - 1) as raw limbs (used as Montgomery residues: uniform field elements)"""
    L = curve.fr_limbs
    a = rs.integers(0, 1 << 64, size=(count, L), dtype=np.uint64)
    a[:, L - 1] &= np.uint64((1 << (curve.r.bit_length() - 64 * (L - 1) - 1)) - 1)     # < r
    return a


def satisfied_instance(curve, log2n, seed) -> Instance:
    import random
    c, r, L = curve, curve.r, curve.fr_limbs
    n = 1 << log2n
    rs = np.random.Generator(np.random.PCG64(seed))
    ql, qr, qm, qk = (_rand_fr(rs, c, n) for _ in range(4))
    qo = np.tile(ff.pack_elements([r - 1], r, L), (n, 1))
    l, rr = _rand_fr(rs, c, n), _rand_fr(rs, c, n)
    half = n // 2
    perm = np.arange(3 * n, dtype=np.int64)
    # L-R copy constraints: l[ia] (first half of the rows) == r[ib] (any row)
    m1 = max(1, n // 4)
    ia = rs.permutation(half)[:m1].astype(np.int64)
    ib = rs.permutation(n)[:m1].astype(np.int64)
    rr[ib] = l[ia]
    perm[ia] = n + ib
    perm[n + ib] = ia

    def gate(rows):
        lr = corelib.fr_vec(c, 0, l[rows], rr[rows])
        acc = corelib.fr_vec(c, 0, qm[rows], lr)
        acc = corelib.fr_vec(c, 1, acc, corelib.fr_vec(c, 0, ql[rows], l[rows]))
        acc = corelib.fr_vec(c, 1, acc, corelib.fr_vec(c, 0, qr[rows], rr[rows]))
        return corelib.fr_vec(c, 1, acc, qk[rows])
    o = np.zeros((n, L), dtype=np.uint64)
    first = slice(0, half)
    o[first] = gate(first)
    # O-L copy constraints: o[ka] (first half, already solved) == l[jb] (second half, not yet used)
    m2 = max(1, n // 8)
    ka = rs.permutation(half)[:m2].astype(np.int64)
    jb = half + rs.permutation(half)[:m2].astype(np.int64)
    l[jb] = o[ka]
    perm[2 * n + ka] = jb
    perm[jb] = 2 * n + ka
    second = slice(half, n)
    o[second] = gate(second)
    rng = random.Random(seed + 1)
    rnd = lambda: rng.randrange(1, r)
    ch = Challenges(gamma=rnd(), beta=rnd(), alpha=rnd(), zeta=rnd(), v=rnd(), bl=[rnd(), rnd()], br=[rnd(), rnd()],
                    bo=[rnd(), rnd()], bz=[rnd(), rnd(), rnd()])
    return Instance(curve=c, log2n=log2n, n=n, ql=ql, qr=qr, qm=qm, qo=qo, qk=qk, perm=perm, l=l, r=rr, o=o, ch=ch,
                    tau=rnd())


def check_satisfied(inst: Instance) -> bool:
    """gate equation on every row and value equality along the permutation (the fixture's own sanity check)"""
    c = inst.curve
    lr = corelib.fr_vec(c, 0, inst.l, inst.r)
    acc = corelib.fr_vec(c, 0, inst.qm, lr)
    acc = corelib.fr_vec(c, 1, acc, corelib.fr_vec(c, 0, inst.ql, inst.l))
    acc = corelib.fr_vec(c, 1, acc, corelib.fr_vec(c, 0, inst.qr, inst.r))
    acc = corelib.fr_vec(c, 1, acc, corelib.fr_vec(c, 0, inst.qo, inst.o))
    acc = corelib.fr_vec(c, 1, acc, inst.qk)
    vals = np.concatenate([inst.l, inst.r, inst.o])
    return not acc.any() and np.array_equal(vals, vals[inst.perm])


def tau_powers(inst: Instance) -> np.ndarray:
    return corelib.fr_geometric(inst.curve, 1, inst.tau, inst.n + 3)


def trapdoor_srs_cpu(inst: Instance) -> np.ndarray:
    c = inst.curve
    return corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), tau_powers(inst))


def trapdoor_srs_gpu(lib, curve, inst: Instance, dev=0) -> np.ndarray:
    """the same SRS built by the library's fixed-base batch (seconds instead of a minute at 2^22; the fixed-base kernels
    have their own parity tests, and a wrong SRS point cannot go unnoticed: the proof would not verify)"""
    return lib.fixed_base_batch(curve.curve_id, 1, ec.pack_points(curve, 1, [curve.g1]), tau_powers(inst), n=inst.n + 3,
                                dev=dev)


def verifying_key(inst: Instance) -> dict:
    """digests of ql qr qm qo qk s1 s2 s3 as q(tau) G (what Setup computes with the SRS, setup.go:120-160)"""
    if inst._vk is not None:
        return inst._vk
    c, n, r = inst.curve, inst.n, inst.curve.r
    F1 = ff.Fp(c.p)
    lag = corelib.fr_lagrange_at(c, inst.log2n, inst.tau)
    w = pow(c.root_of_unity, 1 << (c.two_adicity - inst.log2n), r)
    g = c.mult_gen
    supp = np.concatenate([corelib.fr_geometric(c, 1, w, n), corelib.fr_geometric(c, g, w, n),
                           corelib.fr_geometric(c, g * g % r, w, n)])
    digest = lambda q: ec.scalar_mul(F1, corelib.fr_dot(c, np.ascontiguousarray(q), lag), c.g1)
    s = [supp[inst.perm[j * n:(j + 1) * n]] for j in range(3)]
    inst._vk = {"S1": digest(s[0]), "S2": digest(s[1]), "S3": digest(s[2]), "Ql": digest(inst.ql), "Qr": digest(inst.qr),
                "Qm": digest(inst.qm), "Qo": digest(inst.qo), "Qk": digest(inst.qk), "Qcp": []}
    return inst._vk


def verify(curve, inst: Instance, points_jac: np.ndarray, values: np.ndarray, with_pairing=True) -> bool:
    """points_jac: (10, 3 * fp_limbs) Jacobian limbs as b200_plonk_prove returns them; values: (7, fr_limbs)"""
    c, r = curve, curve.r
    F1 = ff.Fp(c.p)
    pts = [ec.from_jac(F1, ec.unpack_points(c, 1, np.ascontiguousarray(points_jac[k]), ncoords=3)[0]) for k in range(10)]
    vals = ff.unpack_elements(np.ascontiguousarray(values), r, c.fr_limbs)
    key = verifying_key(inst)
    by_trapdoor = lambda lhs, H: lhs == ec.scalar_mul(F1, inst.tau, H)
    if not verify_core(c, inst.n, key, pts, vals, inst.ch, by_trapdoor):
        return False
    if with_pairing:
        from . import pairing
        try:
            T = pairing.get(c)
        except Exception:
            return True          # no oracle pairing for this curve: the trapdoor identity above is the check
        F2 = ff.base_field(c, 2)
        tau_g2 = ec.scalar_mul(F2, inst.tau, c.g2)
        by_pairing = lambda lhs, H: T.product_is_one([(lhs, c.g2), (ec.affine_neg(F1, H), tau_g2)])
        return verify_core(c, inst.n, key, pts, vals, inst.ch, by_pairing)
    return True
