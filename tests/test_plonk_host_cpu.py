"""The C++ PLONK orchestration (gnark_b200/csrc/plonk_host.cu: b200_plonk_pk_load / b200_plonk_prove) on the CPU.
plonk_host.cu is compiled as plain C++ against host stand-ins for the device entry points it calls
(tests/mock/mock_capi.cpp -> libgb200_plonkmock.so; test infrastructure, BN254 and BLS12-381), and its ten digests and
seven opened values are compared with the big-int oracle prover (oracle/plonk_prover.py) under the same injected
challenges and blinding.  Together with the hardware parity tests of each entry point (tests/test_gpu_plonk.py)
this pins everything but the GPU plumbing of tests/test_gpu_zz_late.py::test_plonk_prove_c_abi_vs_oracle."""
import ctypes
import os
import random

import numpy as np
import pytest

from gnark_b200 import lib as b200
from oracle import corelib, ec, ff, plonk_prover as pp
from oracle.params import CURVES
from util import jac_to_affine

LIB = os.environ.get("GB200_PLONKMOCK") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libgb200_plonkmock.so")


@pytest.fixture(scope="module")
def mock():
    if not os.path.exists(LIB):
        pytest.skip("libgb200_plonkmock.so not built (make -C gnark_b200/csrc)")
    m = ctypes.CDLL(LIB)
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    m.b200_plonk_pk_load.argtypes = [i32, i32, ctypes.POINTER(b200.PlonkPkDesc), ctypes.POINTER(vp)]
    m.b200_plonk_pk_free.argtypes = [vp]
    m.b200_plonk_prove.argtypes = [vp, vp, vp, vp, ctypes.POINTER(b200.PlonkChallenges), vp, vp]
    m.b200_plonk_begin.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(vp), vp, ctypes.POINTER(vp), vp]
    m.b200_plonk_commit_z.argtypes = [vp, vp, vp, vp, vp]
    m.b200_plonk_quotient.argtypes = [vp, vp, vp]
    m.b200_plonk_linearise.argtypes = [vp, vp, vp, vp]
    m.b200_plonk_batch_open.argtypes = [vp, vp, vp]
    m.b200_plonk_end.argtypes = [vp]
    m.b200_last_error.restype = ctypes.c_char_p
    return m


@pytest.mark.parametrize("cname,logn", [("bn254", 3), ("bn254", 5), ("bls12-381", 4)])
def test_plonk_host_orchestration_vs_oracle(mock, cname, logn):
    c = CURVES[cname]
    rng = random.Random(3000 + logn)
    r, L = c.r, c.fr_limbs
    n = 1 << logn
    circ, l, rr, o = pp.random_satisfied_instance(c, n, seed=logn + 7)
    rnd = lambda: rng.randrange(r)
    ch = pp.Challenges(gamma=rnd(), beta=rnd(), alpha=rnd(), zeta=rnd(), v=rnd(), bl=[rnd(), rnd()], br=[rnd(), rnd()],
                       bo=[rnd(), rnd()], bz=[rnd(), rnd(), rnd()])
    tau = rnd()
    want = pp.prove(c, circ, l, rr, o, ch, tau)
    assert pp.verify(c, circ, want, ch, tau)
    pe = lambda v: np.ascontiguousarray(ff.pack_elements(v, r, L))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    srs = np.ascontiguousarray(corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), pe([pow(tau, i, r) for i in range(n + 3)])))
    keep = {k: pe(getattr(circ, k)) for k in ("ql", "qr", "qm", "qo", "qk")}
    perm = np.ascontiguousarray(np.array(circ.perm, dtype=np.int64))
    d = b200.PlonkPkDesc()
    d.log2n = logn
    for k, a in keep.items():
        setattr(d, k, P(a).value)
    d.perm, d.srs_canonical = P(perm).value, P(srs).value
    h = ctypes.c_void_p(0)
    assert mock.b200_plonk_pk_load(0, c.curve_id, ctypes.byref(d), ctypes.byref(h)) == 0, mock.b200_last_error()
    sc = {k: pe(v) for k, v in (("gamma", [ch.gamma]), ("beta", [ch.beta]), ("alpha", [ch.alpha]), ("zeta", [ch.zeta]),
                                ("v", [ch.v]), ("bl", ch.bl), ("br", ch.br), ("bo", ch.bo), ("bz", ch.bz))}
    cs = b200.PlonkChallenges()
    for k, a in sc.items():
        setattr(cs, k, P(a).value)
    pts = np.zeros((10, 3 * c.fp_limbs), dtype=np.uint64)
    vals = np.zeros((7, L), dtype=np.uint64)
    L_, R_, O_ = pe(l), pe(rr), pe(o)
    assert mock.b200_plonk_prove(h, P(L_), P(R_), P(O_), ctypes.byref(cs), P(pts), P(vals)) == 0, mock.b200_last_error()
    F = ff.Fp(c.p)
    dl = [want.L, want.R, want.O, want.Z, want.H[0], want.H[1], want.H[2], want.lin, want.batch_opening, want.z_opening]
    for k, name in enumerate(("L", "R", "O", "Z", "H1", "H2", "H3", "lin", "batch", "zopen")):
        assert jac_to_affine(c, 1, pts[k]) == ec.scalar_mul(F, dl[k], c.g1), name
    got = ff.unpack_elements(vals, r, L)
    assert got[:6] == want.claimed and got[6] == want.zu
    # prove -> Verify as the reference tests it: the verifier's equations on the proof points, real pairings
    proof_pts = [jac_to_affine(c, 1, pts[k]) for k in range(10)]
    assert pp.verify_pairing(c, circ, proof_pts, got, ch, tau)
    tampered = list(proof_pts); tampered[8] = ec.affine_add(F, tampered[8], c.g1)
    assert not pp.verify_pairing(c, circ, tampered, got, ch, tau)
    bad_vals = list(got); bad_vals[2] = (bad_vals[2] + 1) % r
    assert not pp.verify_pairing(c, circ, proof_pts, bad_vals, ch, tau)
    # the same proof round by round (the way a Fiat-Shamir transcript drives it), and the stage order is enforced
    jl = 3 * c.fp_limbs
    s_ = ctypes.c_void_p(0)
    lro, zpt, hpts = np.zeros((3, jl), dtype=np.uint64), np.zeros(jl, dtype=np.uint64), np.zeros((3, jl), dtype=np.uint64)
    two, vals2, bpt = np.zeros((2, jl), dtype=np.uint64), np.zeros((7, L), dtype=np.uint64), np.zeros(jl, dtype=np.uint64)
    assert mock.b200_plonk_begin(h, P(L_), P(R_), P(O_), P(sc["bl"]), P(sc["br"]), P(sc["bo"]), None, None,
                                 ctypes.byref(s_), P(lro)) == 0
    assert mock.b200_plonk_quotient(s_, P(sc["alpha"]), P(hpts)) != 0 and b"after plonk_commit_z" in mock.b200_last_error()
    assert mock.b200_plonk_commit_z(s_, P(sc["beta"]), P(sc["gamma"]), P(sc["bz"]), P(zpt)) == 0
    assert mock.b200_plonk_batch_open(s_, P(sc["v"]), P(bpt)) != 0
    assert mock.b200_plonk_quotient(s_, P(sc["alpha"]), P(hpts)) == 0
    assert mock.b200_plonk_linearise(s_, P(sc["zeta"]), P(two), P(vals2)) == 0
    assert mock.b200_plonk_batch_open(s_, P(sc["v"]), P(bpt)) == 0
    assert mock.b200_plonk_end(s_) == 0
    staged = np.concatenate([lro, zpt[None], hpts, two[0:1], bpt[None], two[1:2]])
    assert np.array_equal(staged, pts) and np.array_equal(vals2, vals)
    # completeQk (prove.go:349-373): a key loaded with an INCOMPLETE Qk (public rows still zero) proves the same when the
    # proof's complete Qk is handed in - through the one-call entry point and through b200_plonk_set_qk
    qk_incomplete = keep["qk"].copy()
    qk_incomplete[:2] = 0
    d.qk = P(qk_incomplete).value
    h3 = ctypes.c_void_p(0)
    assert mock.b200_plonk_pk_load(0, c.curve_id, ctypes.byref(d), ctypes.byref(h3)) == 0, mock.b200_last_error()
    pts3, vals3 = np.zeros_like(pts), np.zeros_like(vals)
    assert mock.b200_plonk_prove(h3, P(L_), P(R_), P(O_), ctypes.byref(cs), P(pts3), P(vals3)) == 0
    assert not np.array_equal(pts3, pts)                       # the incomplete key alone proves something else
    cs.qk = P(keep["qk"]).value
    assert mock.b200_plonk_prove(h3, P(L_), P(R_), P(O_), ctypes.byref(cs), P(pts3), P(vals3)) == 0, mock.b200_last_error()
    assert np.array_equal(pts3, pts) and np.array_equal(vals3, vals)
    cs.qk = None
    mock.b200_plonk_set_qk.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    s3 = ctypes.c_void_p(0)
    assert mock.b200_plonk_set_qk(s3, P(keep["qk"])) != 0
    assert mock.b200_plonk_begin(h3, P(L_), P(R_), P(O_), P(sc["bl"]), P(sc["br"]), P(sc["bo"]), None, None,
                                 ctypes.byref(s3), P(lro)) == 0
    assert mock.b200_plonk_commit_z(s3, P(sc["beta"]), P(sc["gamma"]), P(sc["bz"]), P(zpt)) == 0
    assert mock.b200_plonk_set_qk(s3, P(keep["qk"])) == 0
    assert mock.b200_plonk_quotient(s3, P(sc["alpha"]), P(hpts)) == 0
    assert mock.b200_plonk_set_qk(s3, P(keep["qk"])) != 0 and b"before plonk_quotient" in mock.b200_last_error()
    assert mock.b200_plonk_linearise(s3, P(sc["zeta"]), P(two), P(vals2)) == 0
    assert mock.b200_plonk_batch_open(s3, P(sc["v"]), P(bpt)) == 0
    assert mock.b200_plonk_end(s3) == 0
    assert np.array_equal(np.concatenate([lro, zpt[None], hpts, two[0:1], bpt[None], two[1:2]]), pts)
    assert mock.b200_plonk_pk_free(h3) == 0
    d.qk = P(keep["qk"]).value
    # a permutation entry out of range is refused, not dereferenced
    bad = perm.copy(); bad[1] = 3 * n
    d.perm = P(bad).value
    h2 = ctypes.c_void_p(0)
    assert mock.b200_plonk_pk_load(0, c.curve_id, ctypes.byref(d), ctypes.byref(h2)) != 0
    assert b"out of range" in mock.b200_last_error()
    assert mock.b200_plonk_pk_free(h) == 0


@pytest.mark.parametrize("n_commit", (1, 2))
@pytest.mark.parametrize("coset_cache", ("1", "0"))
def test_plonk_host_bsb22_commitments(mock, monkeypatch, n_commit, coset_cache):
    """keys with BSB22 commitment gates: gate term on every coset, [PI2_j], sum_j Qcp_j(zeta) PI2_j(X) in the linearised
    polynomial, Qcp openings - b200_plonk_prove against the oracle prover (extended verifier equations hold)"""
    # GB200_PLONK_COSET_CACHE: key polynomials' coset evaluations kept from key load (default) or recomputed per proof
    monkeypatch.setenv("GB200_PLONK_COSET_CACHE", coset_cache)
    c = CURVES["bn254"]
    logn = 4
    rng = random.Random(4000 + n_commit)
    r, L = c.r, c.fr_limbs
    n = 1 << logn
    circ, l, rr, o, pi2 = pp.random_satisfied_instance(c, n, seed=61, n_commit=n_commit)
    rnd = lambda: rng.randrange(r)
    ch = pp.Challenges(gamma=rnd(), beta=rnd(), alpha=rnd(), zeta=rnd(), v=rnd(), bl=[rnd(), rnd()], br=[rnd(), rnd()],
                       bo=[rnd(), rnd()], bz=[rnd(), rnd(), rnd()])
    tau = rnd()
    want = pp.prove(c, circ, l, rr, o, ch, tau, pi2=pi2)
    assert pp.verify(c, circ, want, ch, tau)
    pe = lambda v: np.ascontiguousarray(ff.pack_elements(v, r, L))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    srs = np.ascontiguousarray(corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), pe([pow(tau, i, r) for i in range(n + 3)])))
    keep = {k: pe(getattr(circ, k)) for k in ("ql", "qr", "qm", "qo", "qk")}
    perm = np.ascontiguousarray(np.array(circ.perm, dtype=np.int64))
    qcp = [pe(v) for v in circ.qcp]
    qarr = (ctypes.c_void_p * n_commit)(*[a.ctypes.data for a in qcp])
    d = b200.PlonkPkDesc()
    d.log2n = logn
    for k, a in keep.items():
        setattr(d, k, P(a).value)
    d.perm, d.srs_canonical = P(perm).value, P(srs).value
    d.n_qcp, d.qcp = n_commit, ctypes.cast(qarr, ctypes.POINTER(ctypes.c_void_p))
    h = ctypes.c_void_p(0)
    assert mock.b200_plonk_pk_load(0, c.curve_id, ctypes.byref(d), ctypes.byref(h)) == 0, mock.b200_last_error()
    sc = {k: pe(v) for k, v in (("gamma", [ch.gamma]), ("beta", [ch.beta]), ("alpha", [ch.alpha]), ("zeta", [ch.zeta]),
                                ("v", [ch.v]), ("bl", ch.bl), ("br", ch.br), ("bo", ch.bo), ("bz", ch.bz))}
    cs = b200.PlonkChallenges()
    for k, a in sc.items():
        setattr(cs, k, P(a).value)
    pts = np.zeros((10, 3 * c.fp_limbs), dtype=np.uint64)
    vals = np.zeros((7 + n_commit, L), dtype=np.uint64)
    bsb = np.zeros((n_commit, 3 * c.fp_limbs), dtype=np.uint64)
    L_, R_, O_ = pe(l), pe(rr), pe(o)
    # without the committed polynomials the call is refused
    assert mock.b200_plonk_prove(h, P(L_), P(R_), P(O_), ctypes.byref(cs), P(pts), P(vals)) != 0
    pi2a = [pe(v) for v in pi2]
    parr = (ctypes.c_void_p * n_commit)(*[a.ctypes.data for a in pi2a])
    cs.pi2, cs.out_bsb22 = ctypes.cast(parr, ctypes.POINTER(ctypes.c_void_p)), P(bsb).value
    assert mock.b200_plonk_prove(h, P(L_), P(R_), P(O_), ctypes.byref(cs), P(pts), P(vals)) == 0, mock.b200_last_error()
    F = ff.Fp(c.p)
    dl = [want.L, want.R, want.O, want.Z, want.H[0], want.H[1], want.H[2], want.lin, want.batch_opening, want.z_opening]
    for k, name in enumerate(("L", "R", "O", "Z", "H1", "H2", "H3", "lin", "batch", "zopen")):
        assert jac_to_affine(c, 1, pts[k]) == ec.scalar_mul(F, dl[k], c.g1), name
    for j in range(n_commit):
        assert jac_to_affine(c, 1, bsb[j]) == ec.scalar_mul(F, want.bsb22[j], c.g1)
    got = ff.unpack_elements(vals, r, L)
    assert got[:6] + got[7:] == want.claimed and got[6] == want.zu
    if coset_cache == "1":      # Verify with real pairings, the BSB22 digests and Qcp openings included
        assert pp.verify_pairing(c, circ, [jac_to_affine(c, 1, pts[k]) for k in range(10)], got, ch, tau,
                                 bsb22_points=[jac_to_affine(c, 1, bsb[j]) for j in range(n_commit)])
    assert mock.b200_plonk_pk_free(h) == 0


@pytest.mark.parametrize("cname", ["bn254", "bls12-381"])
def test_plonk_host_reproduces_golden(mock, cname):
    """the committed PLONK known-answer vector (tests/golden/kat_plonk_v1.json: 8 rows, one BSB22 gate) through the C++
    orchestration on host stand-ins"""
    from test_golden import _plonk_case
    c, circ, l, rr, o, pi2, ch, tau, want = _plonk_case(cname)
    r, L = c.r, c.fr_limbs
    n, logn = circ.n, circ.n.bit_length() - 1
    pe = lambda v: np.ascontiguousarray(ff.pack_elements(v, r, L))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    srs = np.ascontiguousarray(corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), pe([pow(tau, i, r) for i in range(n + 3)])))
    keep = {k: pe(getattr(circ, k)) for k in ("ql", "qr", "qm", "qo", "qk")}
    perm = np.ascontiguousarray(np.array(circ.perm, dtype=np.int64))
    qcp = [pe(v) for v in circ.qcp]
    qarr = (ctypes.c_void_p * len(qcp))(*[a.ctypes.data for a in qcp])
    d = b200.PlonkPkDesc()
    d.log2n = logn
    for k, a in keep.items():
        setattr(d, k, P(a).value)
    d.perm, d.srs_canonical = P(perm).value, P(srs).value
    d.n_qcp, d.qcp = len(qcp), ctypes.cast(qarr, ctypes.POINTER(ctypes.c_void_p))
    h = ctypes.c_void_p(0)
    assert mock.b200_plonk_pk_load(0, c.curve_id, ctypes.byref(d), ctypes.byref(h)) == 0, mock.b200_last_error()
    sc = {k: pe(v) for k, v in (("gamma", [ch.gamma]), ("beta", [ch.beta]), ("alpha", [ch.alpha]), ("zeta", [ch.zeta]),
                                ("v", [ch.v]), ("bl", ch.bl), ("br", ch.br), ("bo", ch.bo), ("bz", ch.bz))}
    cs = b200.PlonkChallenges()
    for k, a in sc.items():
        setattr(cs, k, P(a).value)
    pi2a = [pe(v) for v in pi2]
    parr = (ctypes.c_void_p * len(pi2a))(*[a.ctypes.data for a in pi2a])
    bsb = np.zeros((len(pi2a), 3 * c.fp_limbs), dtype=np.uint64)
    cs.pi2, cs.out_bsb22 = ctypes.cast(parr, ctypes.POINTER(ctypes.c_void_p)), P(bsb).value
    pts = np.zeros((10, 3 * c.fp_limbs), dtype=np.uint64)
    vals = np.zeros((7 + len(pi2a), L), dtype=np.uint64)
    L_, R_, O_ = pe(l), pe(rr), pe(o)
    assert mock.b200_plonk_prove(h, P(L_), P(R_), P(O_), ctypes.byref(cs), P(pts), P(vals)) == 0, mock.b200_last_error()
    F = ff.Fp(c.p)
    Hx = lambda s_: int(s_, 16)
    dl = [Hx(want[k]) for k in ("L", "R", "O", "Z")] + [Hx(x) for x in want["H"]] + [Hx(want["lin"]), Hx(want["batch_opening"]),
                                                                                    Hx(want["z_opening"])]
    for k in range(10):
        assert jac_to_affine(c, 1, pts[k]) == ec.scalar_mul(F, dl[k], c.g1), k
    assert jac_to_affine(c, 1, bsb[0]) == ec.scalar_mul(F, Hx(want["bsb22"][0]), c.g1)
    got = ff.unpack_elements(vals, r, L)
    assert got[:6] + got[7:] == [Hx(x) for x in want["claimed"]] and got[6] == Hx(want["zu"])
    assert mock.b200_plonk_pk_free(h) == 0


def test_plonk_proof_over_the_ethereum_srs_verifies(mock):
    """No trapdoor anywhere: the C++ PLONK orchestration commits with the Ethereum KZG ceremony SRS the reference ships
    (tests/golden/eth_kzg_srs_v1.bin, tau unknown), and the proof is checked by the verifier's pairing equations with
    the fixture's own [tau]_2.  Every commitment is an MSM over externally produced points, every opening must be
    consistent with them - there is no discrete log to compare with."""
    from oracle import kzg_srs
    c = CURVES["bls12-381"]
    logn = 5
    n = 1 << logn
    r, L = c.r, c.fr_limbs
    rng = random.Random(77)
    mono, _, g2 = kzg_srs.load()
    circ, l, rr, o = pp.random_satisfied_instance(c, n, seed=123)
    rnd = lambda: rng.randrange(r)
    ch = pp.Challenges(gamma=rnd(), beta=rnd(), alpha=rnd(), zeta=rnd(), v=rnd(), bl=[rnd(), rnd()], br=[rnd(), rnd()],
                       bo=[rnd(), rnd()], bz=[rnd(), rnd(), rnd()])
    pe = lambda v: np.ascontiguousarray(ff.pack_elements(v, r, L))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    srs = np.ascontiguousarray(ec.pack_points(c, 1, mono[:n + 3]))
    keep = {k: pe(getattr(circ, k)) for k in ("ql", "qr", "qm", "qo", "qk")}
    perm = np.ascontiguousarray(np.array(circ.perm, dtype=np.int64))
    d = b200.PlonkPkDesc()
    d.log2n = logn
    for k, a in keep.items():
        setattr(d, k, P(a).value)
    d.perm, d.srs_canonical = P(perm).value, P(srs).value
    h = ctypes.c_void_p(0)
    assert mock.b200_plonk_pk_load(0, c.curve_id, ctypes.byref(d), ctypes.byref(h)) == 0, mock.b200_last_error()
    sc = {k: pe(v) for k, v in (("gamma", [ch.gamma]), ("beta", [ch.beta]), ("alpha", [ch.alpha]), ("zeta", [ch.zeta]),
                                ("v", [ch.v]), ("bl", ch.bl), ("br", ch.br), ("bo", ch.bo), ("bz", ch.bz))}
    cs = b200.PlonkChallenges()
    for k, a in sc.items():
        setattr(cs, k, P(a).value)
    pts = np.zeros((10, 3 * c.fp_limbs), dtype=np.uint64)
    vals = np.zeros((7, L), dtype=np.uint64)
    L_, R_, O_ = pe(l), pe(rr), pe(o)
    assert mock.b200_plonk_prove(h, P(L_), P(R_), P(O_), ctypes.byref(cs), P(pts), P(vals)) == 0, mock.b200_last_error()
    proof_pts = [jac_to_affine(c, 1, pts[k]) for k in range(10)]
    got = ff.unpack_elements(vals, r, L)
    assert pp.verify_pairing(c, circ, proof_pts, got, ch, srs_g1=mono, tau_g2=g2[1])
    assert not pp.verify_pairing(c, circ, proof_pts, got, ch, srs_g1=mono, tau_g2=g2[2])      # another [tau]_2
    wrong = list(got); wrong[1] = (wrong[1] + 1) % r
    assert not pp.verify_pairing(c, circ, proof_pts, wrong, ch, srs_g1=mono, tau_g2=g2[1])
    assert mock.b200_plonk_pk_free(h) == 0


def test_concurrent_callers_are_serialised_per_device(mock):
    """The reference lets callers invoke Prove concurrently and serialises them per device (icicle.go:53-60); here every
    entry point holds the device's lock from its lookup to its return.  Eight threads prove at once on the mocked C ABI -
    four on one shared key, four on their own keys - and every proof must equal the single-threaded one (ctypes releases
    the GIL during the calls, so the C++ orchestration really is entered concurrently)."""
    import threading
    c = CURVES["bn254"]
    logn = 4
    n = 1 << logn
    r, L = c.r, c.fr_limbs
    pe = lambda v: np.ascontiguousarray(ff.pack_elements(v, r, L))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)

    def make(seed):
        rng = random.Random(seed)
        circ, l, rr, o = pp.random_satisfied_instance(c, n, seed=seed)
        rnd = lambda: rng.randrange(r)
        ch = pp.Challenges(gamma=rnd(), beta=rnd(), alpha=rnd(), zeta=rnd(), v=rnd(), bl=[rnd(), rnd()], br=[rnd(), rnd()],
                           bo=[rnd(), rnd()], bz=[rnd(), rnd(), rnd()])
        tau = rnd()
        keep = {"srs": np.ascontiguousarray(corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), pe([pow(tau, i, r) for i in range(n + 3)]))),
                "perm": np.ascontiguousarray(np.array(circ.perm, dtype=np.int64)),
                **{k: pe(getattr(circ, k)) for k in ("ql", "qr", "qm", "qo", "qk")}}
        d = b200.PlonkPkDesc()
        d.log2n = logn
        for k in ("ql", "qr", "qm", "qo", "qk"):
            setattr(d, k, P(keep[k]).value)
        d.perm, d.srs_canonical = P(keep["perm"]).value, P(keep["srs"]).value
        h = ctypes.c_void_p(0)
        assert mock.b200_plonk_pk_load(0, c.curve_id, ctypes.byref(d), ctypes.byref(h)) == 0, mock.b200_last_error()
        sc = {k: pe(v) for k, v in (("gamma", [ch.gamma]), ("beta", [ch.beta]), ("alpha", [ch.alpha]), ("zeta", [ch.zeta]),
                                    ("v", [ch.v]), ("bl", ch.bl), ("br", ch.br), ("bo", ch.bo), ("bz", ch.bz))}
        cs = b200.PlonkChallenges()
        for k, a in sc.items():
            setattr(cs, k, P(a).value)
        return {"h": h, "cs": cs, "wit": (pe(l), pe(rr), pe(o)), "keep": (keep, sc)}

    def prove(inst):
        pts = np.zeros((10, 3 * c.fp_limbs), dtype=np.uint64)
        vals = np.zeros((7, L), dtype=np.uint64)
        rc = mock.b200_plonk_prove(inst["h"], P(inst["wit"][0]), P(inst["wit"][1]), P(inst["wit"][2]), ctypes.byref(inst["cs"]),
                                   P(pts), P(vals))
        return rc, pts, vals

    insts = [make(900 + i) for i in range(5)]
    want = [prove(i) for i in insts]
    assert all(w[0] == 0 for w in want)
    jobs = [insts[0]] * 4 + insts[1:]                     # four threads share one key, four have their own
    refs = [want[0]] * 4 + want[1:]
    got = [None] * len(jobs)

    def run(k):
        for _ in range(3):
            got[k] = prove(jobs[k])
    ts = [threading.Thread(target=run, args=(k,)) for k in range(len(jobs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for g, w in zip(got, refs):
        assert g[0] == 0 and np.array_equal(g[1], w[1]) and np.array_equal(g[2], w[2])
    for i in insts:
        assert mock.b200_plonk_pk_free(i["h"]) == 0


@pytest.mark.parametrize("cname,logn", [("bn254", 4), ("bls12-381", 3)])
def test_plonk_host_statistical_zk(mock, cname, logn):
    """backend.WithStatisticalZeroKnowledge (prove.go:239-242,689-722,1476-1481): with the two quotient-shard
    randomisers injected, [H1], [H2], [H3] and the linearised digest change as the oracle's do, the opened values do not,
    and the proof verifies (the verifier is the same) - one call and round by round."""
    c = CURVES[cname]
    rng = random.Random(4100 + logn)
    r, L = c.r, c.fr_limbs
    n = 1 << logn
    circ, l, rr, o = pp.random_satisfied_instance(c, n, seed=logn + 11)
    rnd = lambda: rng.randrange(r)
    base = dict(gamma=rnd(), beta=rnd(), alpha=rnd(), zeta=rnd(), v=rnd(), bl=[rnd(), rnd()], br=[rnd(), rnd()],
                bo=[rnd(), rnd()], bz=[rnd(), rnd(), rnd()])
    ch = pp.Challenges(**base, hr=[rnd(), rnd()])
    tau = rnd()
    want = pp.prove(c, circ, l, rr, o, ch, tau)
    plain = pp.prove(c, circ, l, rr, o, pp.Challenges(**base), tau)
    assert pp.verify(c, circ, want, ch, tau) and want.H != plain.H and want.lin != plain.lin
    pe = lambda v: np.ascontiguousarray(ff.pack_elements(v, r, L))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    srs = np.ascontiguousarray(corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), pe([pow(tau, i, r) for i in range(n + 3)])))
    keep = {k: pe(getattr(circ, k)) for k in ("ql", "qr", "qm", "qo", "qk")}
    perm = np.ascontiguousarray(np.array(circ.perm, dtype=np.int64))
    d = b200.PlonkPkDesc()
    d.log2n = logn
    for k, a in keep.items():
        setattr(d, k, P(a).value)
    d.perm, d.srs_canonical = P(perm).value, P(srs).value
    h = ctypes.c_void_p(0)
    assert mock.b200_plonk_pk_load(0, c.curve_id, ctypes.byref(d), ctypes.byref(h)) == 0, mock.b200_last_error()
    sc = {k: pe(v) for k, v in (("gamma", [ch.gamma]), ("beta", [ch.beta]), ("alpha", [ch.alpha]), ("zeta", [ch.zeta]),
                                ("v", [ch.v]), ("bl", ch.bl), ("br", ch.br), ("bo", ch.bo), ("bz", ch.bz), ("hr", ch.hr))}
    cs = b200.PlonkChallenges()
    for k, a in sc.items():
        setattr(cs, k, P(a).value)
    pts = np.zeros((10, 3 * c.fp_limbs), dtype=np.uint64)
    vals = np.zeros((7, L), dtype=np.uint64)
    L_, R_, O_ = pe(l), pe(rr), pe(o)
    assert mock.b200_plonk_prove(h, P(L_), P(R_), P(O_), ctypes.byref(cs), P(pts), P(vals)) == 0, mock.b200_last_error()
    F = ff.Fp(c.p)
    dl = [want.L, want.R, want.O, want.Z, want.H[0], want.H[1], want.H[2], want.lin, want.batch_opening, want.z_opening]
    for k, name in enumerate(("L", "R", "O", "Z", "H1", "H2", "H3", "lin", "batch", "zopen")):
        assert jac_to_affine(c, 1, pts[k]) == ec.scalar_mul(F, dl[k], c.g1), name
    got = ff.unpack_elements(vals, r, L)
    assert got[:6] == want.claimed == plain.claimed and got[6] == want.zu
    proof_pts = [jac_to_affine(c, 1, pts[k]) for k in range(10)]
    assert pp.verify_pairing(c, circ, proof_pts, got, ch, tau)
    # round by round; the randomisers are refused before begin and after the quotient
    mock.b200_plonk_set_quotient_randomizers.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    jl = 3 * c.fp_limbs
    s_ = ctypes.c_void_p(0)
    lro, zpt, hpts = np.zeros((3, jl), dtype=np.uint64), np.zeros(jl, dtype=np.uint64), np.zeros((3, jl), dtype=np.uint64)
    two, vals2, bpt = np.zeros((2, jl), dtype=np.uint64), np.zeros((7, L), dtype=np.uint64), np.zeros(jl, dtype=np.uint64)
    assert mock.b200_plonk_set_quotient_randomizers(s_, P(sc["hr"])) != 0
    assert mock.b200_plonk_begin(h, P(L_), P(R_), P(O_), P(sc["bl"]), P(sc["br"]), P(sc["bo"]), None, None,
                                 ctypes.byref(s_), P(lro)) == 0
    assert mock.b200_plonk_commit_z(s_, P(sc["beta"]), P(sc["gamma"]), P(sc["bz"]), P(zpt)) == 0
    assert mock.b200_plonk_set_quotient_randomizers(s_, P(sc["hr"])) == 0
    assert mock.b200_plonk_quotient(s_, P(sc["alpha"]), P(hpts)) == 0
    assert mock.b200_plonk_set_quotient_randomizers(s_, P(sc["hr"])) != 0 and b"before plonk_quotient" in mock.b200_last_error()
    assert mock.b200_plonk_linearise(s_, P(sc["zeta"]), P(two), P(vals2)) == 0
    assert mock.b200_plonk_batch_open(s_, P(sc["v"]), P(bpt)) == 0
    assert mock.b200_plonk_end(s_) == 0
    assert np.array_equal(np.concatenate([lro, zpt[None], hpts, two[0:1], bpt[None], two[1:2]]), pts)
    assert np.array_equal(vals2, vals)
    assert mock.b200_plonk_pk_free(h) == 0
