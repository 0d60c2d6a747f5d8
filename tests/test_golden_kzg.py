"""EXTERNAL known-answer vectors: the Ethereum KZG ceremony SRS that the reference ships and tests with
(std/evmprecompiles/kzg_trusted_setup.json, used at std/evmprecompiles/10-kzg_point_evaluation_test.go:50-71,853-903;
extracted to tests/golden/eth_kzg_srs_v1.bin by tests/golden/make_golden_kzg.py).  tau is unknown, both G1 bases
are given, so every check below compares a computation of OURS with points computed by SOMEBODY ELSE:

  * constants: g1_monomial[0] / g2_monomial[0] are gnark-crypto's BLS12-381 generators; the size-4096 domain generator
    of oracle/ntt.py (gnark-crypto's root of unity) is the w the ceremony's Lagrange basis is built on.
  * G1 MSM (4096 points, 255-bit scalars):  sum_i w^(ik) lagrange[i] = monomial[k],
                                             (1/n) sum_k w^(-ik) monomial[k] = lagrange[i].
  * Fr NTT (2^12, all orderings):           MSM(monomial, c) = MSM(lagrange, NTT(c)); if one output of the NTT were
                                             wrong, or in the wrong place, the two commitments would differ.
  * KZG commit as PLONK uses it (a9):       commitment in Lagrange form = commitment in canonical form.
  * G2 MSM (65 points over Fp2):            e(MSM(g1_monomial[:65], c), G2) = e(G1, MSM(g2_monomial, c)) with the big-int
                                             pairing of oracle/pairing_bls12_381.py - the fixture ties its G2 half to
                                             its G1 half only through the pairing, so this is what pins Fp2 / G2 results.

CPU here: the big-int oracle, the C++ oracle and the device templates compiled for the host (emulation).  The CUDA path
runs the same checks in tests/test_gpu_zz_late.py::test_cuda_reproduces_eth_kzg_srs.
"""
import ctypes
import random

import numpy as np
import pytest

from oracle import corelib, ec, ff, kzg_srs, ntt
from oracle.params import BLS12_381 as C

N, LOGN = kzg_srs.N, kzg_srs.LOGN
F = ff.Fp(C.p)
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def srs():
    mono, lag, g2 = kzg_srs.load()
    return dict(mono=mono, lag=lag, g2=g2, MONO=ec.pack_points(C, 1, mono), LAG=ec.pack_points(C, 1, lag))


def kat_cases(rng):
    """(which base, scalars, expected point index in the other base)"""
    w = ntt.Domain(C, N).generator
    ninv = pow(N, C.r - 2, C.r)
    ks = [1, 2, N - 1, rng.randrange(3, N - 1)]
    fwd = [("LAG", [pow(w, i * k % N, C.r) for i in range(N)], ("mono", k)) for k in ks]
    idx = [0, 1, rng.randrange(2, N)]
    bwd = [("MONO", [ninv * pow(w, (-i * k) % N, C.r) % C.r for k in range(N)], ("lag", i)) for i in idx]
    return fwd + bwd


def cpp_msm(pts, sc):
    out = corelib.msm(C, 1, pts, ff.pack_elements(sc, C.r, C.fr_limbs))
    return ec.from_jac(F, ec.unpack_points(C, 1, out, ncoords=3)[0])


def test_constants_pinned(srs):
    assert srs["mono"][0] == C.g1 and srs["g2"][0] == C.g2
    assert all(ec.is_on_curve(F, p, C.b) for p in srs["mono"][:8] + srs["lag"][:8])
    F2 = ff.base_field(C, 2)
    assert all(ec.curve_b_of(F2, p) == (4, 4) for p in srs["g2"])
    # the ceremony's w (EIP-4844: 7^((r-1)/4096)) is gnark-crypto's generator of the size-4096 domain, 7 = FrMultiplicativeGen
    assert ntt.Domain(C, N).generator == pow(7, (C.r - 1) // N, C.r) and C.mult_gen == 7
    # pure big-int group law, no MSM: sum of the Lagrange basis = [1]
    acc = None
    for p in srs["lag"]:
        acc = ec.affine_add(F, acc, p)
    assert acc == C.g1


def test_msm_known_answers_cpp_oracle(srs):
    for base, sc, (other, j) in kat_cases(random.Random(1)):
        assert cpp_msm(srs[base], sc) == srs[other][j]
    # a smaller one through the big-int oracle alone: the 64 Lagrange points i = 0 mod 64 carry sum_{j<64} L_{64j},
    # i.e. the polynomial (X^4096 - 1) / (64 (X^64 - 1)) = (1/64) sum_{m<64} X^(64 m)
    inv64 = pow(64, C.r - 2, C.r)
    lhs = None
    for j in range(64):
        lhs = ec.affine_add(F, lhs, srs["lag"][64 * j])
    assert lhs == ec.msm_naive(F, [srs["mono"][64 * m] for m in range(64)], [inv64] * 64)


def test_ntt_pinned_through_commitments(srs):
    rng = random.Random(2)
    dom = ntt.Domain(C, N)
    c = [rng.randrange(C.r) for _ in range(N)]
    want = cpp_msm(srs["MONO"], c)
    # big-int oracle: DIF leaves bit-reversed order, DIT consumes it
    e_br = dom.fft(c, ntt.DIF)
    e = ntt.bit_reverse(list(e_br))
    assert cpp_msm(srs["LAG"], e) == want
    assert dom.fft(ntt.bit_reverse(list(c)), ntt.DIT) == e
    assert dom.fft_inverse(e, ntt.DIF) == ntt.bit_reverse(list(c)) and dom.fft_inverse(e_br, ntt.DIT) == c
    # C++ oracle
    A = ff.pack_elements(c, C.r, C.fr_limbs)
    corelib.ntt(C, A, LOGN, False, ntt.DIF, False)
    assert ff.unpack_elements(A, C.r, C.fr_limbs) == e_br


def test_device_templates_on_external_vectors(hostemu, srs):
    """the CUDA kernels' per-thread code, compiled for the host (tests/test_emulation.py), on the external vectors;
    several window / task / chunk geometries"""
    base, sc, (other, j) = kat_cases(random.Random(3))[3]
    for lib, (cw, pre, tl, ch) in ((hostemu, (8, 0, 16, 64)), (hostemu, (10, 0, 32, 128)), (hostemu, (9, 0, 64, 128))):
        out = np.zeros(3 * C.fp_limbs, dtype=np.uint64)
        assert lib.emu_msm(C.curve_id, 1, P(srs[base]), P(ff.pack_elements(sc, C.r, C.fr_limbs)), N, cw, pre, tl, ch, P(out)) == 0
        assert ec.from_jac(F, ec.unpack_points(C, 1, out, ncoords=3)[0]) == srs[other][j]
    # precomputed-window tables (one emulated doubling chain per point and window: kept to the 64-point identity
    # (1/64) sum_m monomial[64 m] = sum_j lagrange[64 j], right-hand side by big-int additions)
    want = None
    for jj in range(64):
        want = ec.affine_add(F, want, srs["lag"][64 * jj])
    sub = ec.pack_points(C, 1, [srs["mono"][64 * m] for m in range(64)])
    inv64 = ff.pack_elements([pow(64, C.r - 2, C.r)] * 64, C.r, C.fr_limbs)
    for (cw, pre, tl, ch) in ((12, 1, 64, 256), (16, 1, 64, 512)):
        out = np.zeros(3 * C.fp_limbs, dtype=np.uint64)
        assert hostemu.emu_msm(C.curve_id, 1, P(sub), P(inv64), 64, cw, pre, tl, ch, P(out)) == 0
        assert ec.from_jac(F, ec.unpack_points(C, 1, out, ncoords=3)[0]) == want
    rng = random.Random(4)
    c = [rng.randrange(C.r) for _ in range(N)]
    A = ff.pack_elements(c, C.r, C.fr_limbs)
    assert hostemu.emu_ntt(C.curve_id, P(A), LOGN, 0, ntt.DIF, 0, None, None) == 0
    e = ntt.bit_reverse(ff.unpack_elements(A, C.r, C.fr_limbs))
    assert cpp_msm(srs["LAG"], e) == cpp_msm(srs["MONO"], c)


def g2_case(srs):
    rng = random.Random(5)
    c = [rng.randrange(C.r) for _ in range(kzg_srs.N_G2)]
    c[3], c[7] = 0, C.r - 1
    SC = ff.pack_elements(c, C.r, C.fr_limbs)
    A = cpp_msm(ec.pack_points(C, 1, srs["mono"][:kzg_srs.N_G2]), c)      # G1 side: pinned by the known answers above
    return c, SC, A


def test_pairing_oracle_and_srs_consistency(srs):
    from oracle import pairing_bls12_381 as pr
    e = pr.pairing(C.g1, C.g2)
    assert e != pr.Fp12.one() and e ** C.r == pr.Fp12.one()
    F2 = ff.base_field(C, 2)
    assert pr.pairing(ec.scalar_mul(F, 5, C.g1), C.g2) == pr.pairing(C.g1, ec.scalar_mul(F2, 5, C.g2)) == e ** 5
    # the fixture: e(tau^k G1, G2) = e(G1, tau^k G2)
    neg_g1 = ec.affine_neg(F, C.g1)
    for k in (1, 2, kzg_srs.N_G2 - 1):
        assert pr.pairing_product_is_one([(srs["mono"][k], C.g2), (neg_g1, srs["g2"][k])])
    assert not pr.pairing_product_is_one([(srs["mono"][2], C.g2), (neg_g1, srs["g2"][3])])


def test_g2_msm_pinned_by_pairing(hostemu, srs):
    from oracle import pairing_bls12_381 as pr
    F2 = ff.base_field(C, 2)
    c, SC, A = g2_case(srs)
    G2P = ec.pack_points(C, 2, srs["g2"])
    B = ec.msm_naive(F2, srs["g2"], c)                                     # big-int oracle
    assert pr.pairing_product_is_one([(A, C.g2), (ec.affine_neg(F, C.g1), B)])
    out = corelib.msm(C, 2, G2P, SC)                                       # C++ oracle
    assert ec.from_jac(F2, ec.unpack_points(C, 2, out, ncoords=3)[0]) == B
    for (cw, pre, tl, ch) in ((8, 0, 16, 64), (12, 1, 64, 256)):           # device templates on the host
        out = np.zeros(3 * 2 * C.fp_limbs, dtype=np.uint64)
        assert hostemu.emu_msm(C.curve_id, 2, P(G2P), P(SC), kzg_srs.N_G2, cw, pre, tl, ch, P(out)) == 0
        assert ec.from_jac(F2, ec.unpack_points(C, 2, out, ncoords=3)[0]) == B


def test_gnark_vk_constants():
    """Constants gnark-crypto itself computed, read out of gnark's own serialised PLONK verifying keys
    (backend/solidity/testdata/blank_plonk_*.vk -> tests/golden/gnark_vk_constants_v1.json, made by
    tests/golden/make_golden_vk_constants.py): fft.Domain generator of size 8 / 16, CardinalityInv, coset shift,
    G1 and G2 generators (compressed: x plus the 'y is the larger root' flag) for BN254 and BLS12-381."""
    import json
    import os
    from oracle.params import BN254
    kat = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gnark_vk_constants_v1.json")))
    assert len(kat["keys"]) == 4
    for k in kat["keys"]:
        c = {"bn254": BN254, "bls12381": C}[k["curve"]]
        dom = ntt.Domain(c, k["size"])
        assert dom.generator == int(k["generator"], 16)
        assert int(k["size_inv"], 16) * k["size"] % c.r == 1
        assert c.mult_gen == int(k["coset_shift"], 16) == dom.coset_gen
        nflag = 2 if c is BN254 else 3          # gnark-crypto: BN254 keeps 2 flag bits, BLS12-381 the 3 ZCash bits
        half = (c.p - 1) // 2
        g1 = bytes.fromhex(k["kzg_g1_compressed"])
        x = int.from_bytes(g1, "big") & ((1 << (8 * len(g1) - nflag)) - 1)
        larger = (g1[0] >> 6) == 3 if c is BN254 else bool(g1[0] & 0x20)
        assert g1[0] & 0x80 and x == c.g1[0] and larger == (c.g1[1] > half)
        g2 = bytes.fromhex(k["kzg_g2_0_compressed"])
        fpb = len(g2) // 2
        x1 = int.from_bytes(g2[:fpb], "big") & ((1 << (8 * fpb - nflag)) - 1)
        x0 = int.from_bytes(g2[fpb:], "big")
        (gx0, gx1), (gy0, gy1) = c.g2
        larger = (g2[0] >> 6) == 3 if c is BN254 else bool(g2[0] & 0x20)
        assert g2[0] & 0x80 and (x0, x1) == (gx0, gx1)
        assert larger == ((gy1 > half) if gy1 != 0 else (gy0 > half))


def test_encoding_matches_reference_bytes():
    """oracle/encoding.py (gnark-crypto's point encodings, which the device decoder of b200_table_upload_encoded is
    checked against) on bytes the reference holds: re-encoding the decoded Ethereum ceremony points reproduces the
    fixture's 8192 compressed BLS12-381 G1 encodings byte for byte, and the compressed G1 generators inside gnark's own
    serialised PLONK verifying keys decode to the generators (BN254: the 2-bit metadata scheme, BLS12-381: 3 bits)."""
    import json
    import os
    from oracle import encoding, kzg_srs
    from oracle.params import CURVES
    c = CURVES["bls12-381"]
    blob = open(kzg_srs.PATH, "rb").read()
    mono, lag, _ = kzg_srs.load()
    for i, P_ in enumerate(mono + lag):
        enc = blob[48 * i:48 * (i + 1)]
        assert encoding.encode_g1(c, P_, True) == enc, i
        if i % 97 == 0:
            assert encoding.decode_g1(c, enc) == P_
            assert encoding.decode_g1(c, encoding.encode_g1(c, P_, False)) == P_
    vk = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gnark_vk_constants_v1.json")))
    for k in vk["keys"]:
        cc = CURVES["bn254" if k["curve"] == "bn254" else "bls12-381"]
        assert encoding.decode_g1(cc, bytes.fromhex(k["kzg_g1_compressed"])) == cc.g1
        assert encoding.encode_g1(cc, cc.g1, True).hex() == k["kzg_g1_compressed"]
