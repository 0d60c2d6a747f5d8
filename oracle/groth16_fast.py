"""Full-size Groth16 fixtures and checks at sizes where Python lists of big ints are not affordable (BASELINE
configs[2]: BN254, 2^20 constraints).  TEST INFRASTRUCTURE ONLY (see oracle/params.py header): imported by tests/
and by bench.py to BUILD the workload and to CHECK a proof, never by the product.

The circuit is a SATISFIED product network of m constraints over S interleaved lanes (vectorisable, unlike the
square chain of backend/groth16/groth16_test.go:126-156 whose m squarings are sequential):

    constraint k (k < m):   v_k * v_max(k-1,0) = v_(k+S)            values v_0 .. v_(m+S-1)
    wires (gnark order):    0 = one, 1 = y = v_(m+S-1) (public) | 2 + j = v_j, j < m+S-1 (secret v_0..v_(S-1), internal)

so A, B and K filter different wire sets (InfinityA / InfinityB as in backend/groth16/bn254/setup.go:194-219) and one
wire sits in two B rows.  The trapdoor is known, hence every proving-key element has a known discrete logarithm:

    setupABC   backend/groth16/bn254/setup.go:346-428   A_i = sum_k L_k(t) [wire i in L-row k], ...
    K, Z       setup.go:128-149,247-249                   K_i = (beta A_i + alpha B_i + C_i) / delta (private wires),
                                                          Z_i = t^i (t^n - 1) / delta, bit-reversed, n - 1 entries
    Prove      backend/groth16/bn254/prove.go:185-292      Ar, Bs, Krs from the five MSMs and r, s

and the expected value of each MSM is one dot product in Fr (corelib.fr_dot), the proof points are dlog * G, and the
verifier's equation e(Ar, Bs) = e(alpha, beta) e(sum w_i K_i, gamma) e(Krs, delta) (verify.go:38-140) is checked both in
the exponent and with a real pairing.  All vector arithmetic runs in the C++ oracle; `as_r1cs` rebuilds the same
circuit as an oracle/groth16.R1CS so that, at small sizes, this module is itself checked against the big-int oracle
(tests/test_groth16_fast.py).
"""

import random
from dataclasses import dataclass

import numpy as np

from . import corelib, ec, ff
from . import groth16 as g16


@dataclass
class Instance:
    curve: object
    logn: int
    n: int                # domain size
    m: int                # constraints (<= n)
    lanes: int            # S
    v: np.ndarray         # (m + S, fr_limbs) Montgomery: every value of the network
    toxic: g16.Toxic
    a_dl: np.ndarray      # discrete logs of G1.A / G1.B / G2.B / G1.K / G1.Z in ProvingKey order (Montgomery arrays)
    b_dl: np.ndarray
    k_dl: np.ndarray
    z_dl: np.ndarray
    vk_k: list            # [0, dlog of the public wire's K] (ints)
    inf_a: np.ndarray
    inf_b: np.ndarray

    nb_public = 2

    @property
    def nb_wires(self):
        return self.m + self.lanes + 1

    # ---- the solver's output (constraint/bn254/system.go:162-165), Montgomery limb arrays --------------------------
    def wires(self) -> np.ndarray:
        c = self.curve
        one = ff.pack_elements([1], c.r, c.fr_limbs)
        return np.ascontiguousarray(np.concatenate([one, self.v[-1:], self.v[:-1]]))

    def solution_abc(self):
        m, S = self.m, self.lanes
        a = self.v[:m]
        b = np.concatenate([self.v[:1], self.v[:m - 1]])
        return (np.ascontiguousarray(a), np.ascontiguousarray(b), np.ascontiguousarray(self.v[S:S + m]))

    # ---- the same circuit for the big-int oracle (small sizes only) -------------------------------------------------
    def wire_of(self, j):
        return 1 if j == self.m + self.lanes - 1 else 2 + j

    def as_r1cs(self) -> g16.R1CS:
        cons = [([(1, self.wire_of(k))], [(1, self.wire_of(max(k - 1, 0)))], [(1, self.wire_of(k + self.lanes))])
                for k in range(self.m)]
        return g16.R1CS(nb_public=2, nb_secret=self.lanes, nb_internal=self.m - 1, constraints=cons)

    def witness_ints(self):
        c = self.curve
        return ff.unpack_elements(self.wires(), c.r, c.fr_limbs)


def _const(curve, value, count):
    return np.tile(ff.pack_elements([value % curve.r], curve.r, curve.fr_limbs), (count, 1))


def _bitrev_perm(logn):
    idx = np.arange(1 << logn, dtype=np.uint32)
    rev = np.zeros_like(idx)
    for k in range(logn):
        rev |= ((idx >> k) & 1) << (logn - 1 - k)
    return rev


def satisfied_instance(curve, logn, seed, m=None, lanes=1024) -> Instance:
    c, r, L = curve, curve.r, curve.fr_limbs
    n = 1 << logn
    m = n - 1 if m is None else m
    S = min(lanes, m)
    assert 2 <= S <= m <= n
    rs = np.random.Generator(np.random.PCG64(seed))
    v = np.zeros((m + S, L), dtype=np.uint64)
    seedv = rs.integers(0, 1 << 64, size=(S, L), dtype=np.uint64)
    seedv[:, L - 1] &= np.uint64((1 << (r.bit_length() - 64 * (L - 1) - 1)) - 1)      # < r: uniform Montgomery residues
    v[:S] = seedv
    # v_(k+S) = v_k * v_max(k-1,0), a block of S constraints at a time (block t needs v up to index t S + S - 1)
    for lo in range(0, m, S):
        hi = min(lo + S, m)
        left = v[lo:hi]
        right = np.concatenate([v[:1], v[:hi - 1]])[lo:hi] if lo == 0 else v[lo - 1:hi - 1]
        v[lo + S:hi + S] = corelib.fr_vec(c, 0, left, right)
    tox = g16.random_toxic(c, seed + 1)
    t = tox.t
    lag = corelib.fr_lagrange_at(c, logn, t)[:m]                  # L_k(t), k < m   (setup.go:346-428)
    npriv = m + S - 1                                             # values v_0 .. v_(m+S-2) are the private wires
    a_dl = np.ascontiguousarray(lag)                              # wire 2 + j carries L_j in A, j < m
    b_full = np.zeros((m - 1, L), dtype=np.uint64)                # wire 2 + j carries L_(j+1) in B (and L_0 for j = 0)
    b_full[:] = lag[1:m]
    b_full[:1] = corelib.fr_vec(c, 1, lag[1:2], lag[0:1])
    c_priv = np.zeros((npriv, L), dtype=np.uint64)                # value j is the output of constraint j - S
    c_priv[S:] = lag[:npriv - S]
    a_pad = np.zeros((npriv, L), dtype=np.uint64); a_pad[:m] = a_dl
    b_pad = np.zeros((npriv, L), dtype=np.uint64); b_pad[:m - 1] = b_full
    dinv = pow(tox.delta, -1, r)
    k_dl = corelib.fr_vec(c, 1, corelib.fr_vec(c, 0, a_pad, _const(c, tox.beta * dinv, npriv)),
                          corelib.fr_vec(c, 0, b_pad, _const(c, tox.alpha * dinv, npriv)))
    k_dl = corelib.fr_vec(c, 1, k_dl, corelib.fr_vec(c, 0, c_priv, _const(c, dinv, npriv)))
    lag_last = ff.unpack_elements(lag[m - 1:m], r, L)[0]          # the public output y = v_(m+S-1) closes constraint m - 1
    vk_k = [0, lag_last * pow(tox.gamma, -1, r) % r]
    zdt = (pow(t, n, r) - 1) * dinv % r
    z_dl = np.ascontiguousarray(corelib.fr_geometric(c, zdt, t, n)[_bitrev_perm(logn)][:n - 1])
    nw = m + S + 1
    inf_a = np.ones(nw, dtype=np.uint8); inf_a[2:2 + m] = 0
    inf_b = np.ones(nw, dtype=np.uint8); inf_b[2:2 + m - 1] = 0
    return Instance(curve=c, logn=logn, n=n, m=m, lanes=S, v=v, toxic=tox, a_dl=a_dl, b_dl=b_full, k_dl=k_dl, z_dl=z_dl,
                    vk_k=vk_k, inf_a=inf_a, inf_b=inf_b)


def check_satisfied(inst: Instance) -> bool:
    a, b, cc = inst.solution_abc()
    return np.array_equal(corelib.fr_vec(inst.curve, 0, a, b), cc)


def compute_h(inst: Instance) -> np.ndarray:
    """prove.go:346-389 on the C++ oracle; h in bit-reversed order, n entries"""
    c, n, L = inst.curve, inst.n, inst.curve.fr_limbs
    pads = []
    for x in inst.solution_abc():
        p = np.zeros((n, L), dtype=np.uint64)
        p[:inst.m] = x
        pads.append(p)
    return corelib.compute_h(c, pads[0], pads[1], pads[2], inst.logn)


@dataclass
class Expected:
    msm_a: int
    msm_b: int
    msm_z: int
    msm_k: int
    ar: int
    bs: int
    krs: int


def expected(inst: Instance, rr: int, ss: int) -> Expected:
    """discrete logs of the five MSM results and of the three proof points (prove.go:185-292)"""
    c, r, m, S = inst.curve, inst.curve.r, inst.m, inst.lanes
    tox = inst.toxic
    h = compute_h(inst)
    msm_a = corelib.fr_dot(c, inst.a_dl, np.ascontiguousarray(inst.v[:m]))
    msm_b = corelib.fr_dot(c, inst.b_dl, np.ascontiguousarray(inst.v[:m - 1]))
    msm_k = corelib.fr_dot(c, inst.k_dl, np.ascontiguousarray(inst.v[:m + S - 1]))
    msm_z = corelib.fr_dot(c, inst.z_dl, np.ascontiguousarray(h[:inst.n - 1]))
    ar = (msm_a + tox.alpha + rr * tox.delta) % r
    bs = (msm_b + tox.beta + ss * tox.delta) % r
    krs = (msm_k + msm_z + (-rr * ss % r) * tox.delta + ss * ar + rr * bs) % r
    return Expected(msm_a, msm_b, msm_z, msm_k, ar, bs, krs)


def public_input_dlog(inst: Instance) -> int:
    """sum_i w_i vkK_i over the public wires (verify.go:96-116)"""
    c = inst.curve
    y = ff.unpack_elements(inst.v[-1:], c.r, c.fr_limbs)[0]
    return inst.vk_k[1] * y % c.r


def verify_in_exponent(inst: Instance, e: Expected) -> bool:
    r, tox = inst.curve.r, inst.toxic
    return e.ar * e.bs % r == (tox.alpha * tox.beta + public_input_dlog(inst) * tox.gamma + e.krs * tox.delta) % r


def verify_points(inst: Instance, ar_pt, bs_pt, krs_pt, e: Expected, with_pairing=True) -> bool:
    """the proof POINTS (affine, oracle/ec conventions) against dlog * generator, then the verifier's pairing equation on
    the points themselves (curves with an oracle pairing)"""
    from . import pairing
    c = inst.curve
    F1, F2 = ff.Fp(c.p), ff.base_field(c, 2)
    ok = (ar_pt == ec.scalar_mul(F1, e.ar, c.g1) and bs_pt == ec.scalar_mul(F2, e.bs, c.g2)
          and krs_pt == ec.scalar_mul(F1, e.krs, c.g1) and verify_in_exponent(inst, e))
    if ok and with_pairing and c.name in pairing.TOWER:
        class _Vk:                      # what groth16.verify_pairing reads of a ProvingKeyDlog
            alpha, beta, gamma, delta, vk_K = inst.toxic.alpha, inst.toxic.beta, inst.toxic.gamma, inst.toxic.delta, inst.vk_k
        y = ff.unpack_elements(inst.v[-1:], c.r, c.fr_limbs)[0]
        ok = g16.verify_pairing(c, _Vk, ar_pt, bs_pt, krs_pt, [1, y])
    return ok


def key_points(inst: Instance, fixed_base):
    """the proving key as points: fixed_base(group, dlogs_montgomery_array) -> affine points (the C++ oracle's
    corelib.fixed_base on the CPU, or the library's b200_fixed_base_batch in a GPU test / the bench)"""
    c, tox = inst.curve, inst.toxic
    three = ff.pack_elements([tox.alpha, tox.beta, tox.delta], c.r, c.fr_limbs)
    g1s = fixed_base(1, three)
    g2s = fixed_base(2, three[1:])
    return dict(alpha=g1s[0], beta=g1s[1], delta=g1s[2], A=fixed_base(1, inst.a_dl), B=fixed_base(1, inst.b_dl),
                Z=fixed_base(1, inst.z_dl), K=fixed_base(1, inst.k_dl), beta2=g2s[0], delta2=g2s[1],
                B2=fixed_base(2, inst.b_dl))
