"""GPU tests of the round-2 entry points: the library's own communicator + device-side combine (comm.cu),
b200_msm_submit_dev, the PLONK prover at large sizes checked by the verifier's equations, concurrent Groth16 proofs on
one key, the 2^31-entry guard."""
import random
import threading

import numpy as np
import pytest

from oracle import ec, ff
from oracle.params import CURVES
from util import jac_to_affine, known_dlog_instance

pytestmark = pytest.mark.gpu


def test_points_allreduce_without_communicator_is_a_copy(gpu):
    import torch
    c = CURVES["bn254"]
    assert gpu.comm_info(0) == (1, 0)
    src = torch.arange(5 * 12, dtype=torch.int64, device="cuda")
    dst = torch.zeros_like(src)
    torch.cuda.synchronize()
    gpu.points_allreduce(0, c.curve_id, 1, src, 5, dst)
    gpu.sync(0)
    assert torch.equal(src, dst)


@pytest.mark.parametrize("cname,group", [("bn254", 1), ("bn254", 2), ("bls12-381", 1), ("bw6-761", 1)])
def test_fold_kernel_sums_partials(gpu, cname, group):
    """b200_points_fold (the kernel half of b200_points_allreduce) on one device: `world` rows of partial MSM results
    gathered by hand, empty ranges (the point at infinity) among them, against host-side additions"""
    import torch
    c = CURVES[cname]
    world, count, n = 3, 4, 300
    F, base, pts, sc, _ = known_dlog_instance(c, group, n, seed=5)
    t = gpu.Table(c.curve_id, group, pts, precomp=False)
    L = c.fr_limbs
    rows, want = [], []
    rng = random.Random(4)
    scl = ff.unpack_elements(sc, c.r, L)
    for r in range(world):
        row = []
        for k in range(count):
            lo = rng.randrange(0, n - 40)
            cnt = rng.randrange(0, 40)        # includes empty ranges: the point at infinity as a partial
            row.append(t.msm(np.ascontiguousarray(sc.reshape(n, L)[lo:lo + cnt]), off=lo, n=cnt))
        rows.append(row)
    g = torch.from_numpy(np.stack([np.stack(r_) for r_ in rows]).view(np.int64)).cuda()
    out = torch.zeros((count, rows[0][0].size), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    gpu.points_fold(0, c.curve_id, group, g, world, count, out)
    gpu.sync(0)
    for k in range(count):
        acc = rows[0][k].copy()
        for r in range(1, world):
            gpu.point_add_jac(c.curve_id, group, acc, rows[r][k])
        assert jac_to_affine(c, group, out[k].cpu().numpy().view(np.uint64)) == jac_to_affine(c, group, acc)
    t.free()


def test_msm_submit_dev_and_allreduce_single_device(gpu):
    import torch
    c = CURVES["bn254"]
    n = 5000
    F, base, pts, sc, expected = known_dlog_instance(c, 1, n, seed=31)
    t = gpu.Table(c.curve_id, 1, pts, precomp=True)
    h_sc = torch.from_numpy(sc.view(np.int64).copy()).pin_memory()
    K = 4
    d_out = torch.zeros((K, 12), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for i in range(K):
        t.msm_submit_dev(h_sc, d_out[i], n=n)
    gpu.sync(0)
    for i in range(K):
        assert jac_to_affine(c, 1, d_out[i].cpu().numpy().view(np.uint64)) == expected
    assert jac_to_affine(c, 1, t.msm_allreduce(sc, n=n)) == expected          # no communicator: b200_msm
    t.free()


def test_two_devices_one_process_allreduce(gpu):
    """b200_comm_init_all + b200_msm_allreduce from one thread per device (the Go shape: a goroutine per device):
    both devices end up with the sum of the two shards"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    c = CURVES["bn254"]
    n = 4096
    F, base, pts, sc, expected = known_dlog_instance(c, 1, n, seed=32)
    gpu.init([0, 1])
    gpu.comm_init_all([0, 1])
    per = 2 * c.fp_limbs
    P, S = pts.reshape(n, per), sc.reshape(n, c.fr_limbs)
    half = n // 2
    tabs = [gpu.Table(c.curve_id, 1, np.ascontiguousarray(P[d * half:(d + 1) * half]), dev=d, precomp=True) for d in (0, 1)]
    res, errs = [None, None], [None, None]

    def run(d):
        try:
            res[d] = tabs[d].msm_allreduce(np.ascontiguousarray(S[d * half:(d + 1) * half]), n=half)
        except Exception as e:
            errs[d] = e
    th = [threading.Thread(target=run, args=(d,)) for d in (0, 1)]
    [x.start() for x in th]
    [x.join(timeout=120) for x in th]
    assert errs == [None, None], errs
    for d in (0, 1):
        assert jac_to_affine(c, 1, res[d]) == expected
        tabs[d].free()
    gpu.comm_destroy(0)
    gpu.comm_destroy(1)


@pytest.mark.parametrize("cname,logn", [("bls12-381", 16), ("bls12-381", 20), ("bn254", 18)])
def test_plonk_prove_large_verifies(gpu, cname, logn):
    """b200_plonk_prove on a satisfied 2^logn-gate instance with a trapdoor SRS; the ten points and seven values pass the
    verifier's equations (oracle/plonk_fast.py: verifying key from the C++ oracle, openings by the trapdoor identity AND
    by real pairings); a proof of a DIFFERENT witness column does not.  bench.py runs the same at 2^22 (config 4)."""
    from oracle import plonk_fast
    c = CURVES[cname]
    inst = plonk_fast.satisfied_instance(c, logn, seed=logn)
    assert plonk_fast.check_satisfied(inst)
    srs = plonk_fast.trapdoor_srs_gpu(gpu, c, inst)
    key = gpu.PlonkKey(c.curve_id, logn, inst.ql, inst.qr, inst.qm, inst.qo, inst.qk, inst.perm, srs)
    pts, vals = key.prove(inst.l, inst.r, inst.o, *inst.challenges_packed())
    st = key.last_stage_ms()
    assert set(st) == set(gpu.PlonkKey.STAGES) and all(v > 0 for v in st.values())
    assert plonk_fast.verify(c, inst, pts, vals, with_pairing=(logn <= 16))
    bad_o = inst.o.copy()
    bad_o[7] = inst.o[8]
    pts2, vals2 = key.prove(inst.l, inst.r, bad_o, *inst.challenges_packed())
    assert not plonk_fast.verify(c, inst, pts2, vals2, with_pairing=False)
    key.free()


def test_plonk_rejects_tiny_domains(gpu):
    c = CURVES["bn254"]
    z = np.zeros((4, c.fr_limbs), dtype=np.uint64)
    with pytest.raises(gpu.B200Error):
        gpu.PlonkKey(c.curve_id, 2, z, z, z, z, z, np.arange(12, dtype=np.int64), np.zeros((7, 2 * c.fp_limbs), dtype=np.uint64))


def test_concurrent_groth16_proofs_on_one_key(gpu):
    """two host threads prove on the SAME device-resident key at once (goroutines in the Go shim): no per-key lock, calls
    are ordered by the device context's lock and every proof has its own stream-ordered buffers; both proofs equal the
    single-threaded ones"""
    from gnark_b200 import groth16 as b200
    from oracle import groth16 as g16
    from util import build_groth16_pk, pack_solution
    c = CURVES["bn254"]
    cs, W = g16.square_chain_r1cs(500), g16.square_chain_witness(c.r, 500)
    pk, pkd, _, _ = build_groth16_pk(c, cs, g16.random_toxic(c, 9), 9)
    sol = pack_solution(c, cs, W)
    rs = [[11, 12], [13, 14]]

    def prove(k):
        it = iter(rs[k])
        return b200.ProveSolution(pk, sol, b200.WithDeviceID(0), b200.WithRandomness(lambda q: next(it)))
    want = [prove(0), prove(1)]
    got, errs = [None, None], [None, None]

    def run(k):
        try:
            for _ in range(3):
                got[k] = prove(k)
        except Exception as e:
            errs[k] = e
    th = [threading.Thread(target=run, args=(k,)) for k in (0, 1)]
    [x.start() for x in th]
    [x.join(timeout=300) for x in th]
    assert errs == [None, None], errs
    for k in (0, 1):
        for f in ("Ar", "Bs", "Krs"):
            assert np.array_equal(getattr(got[k], f), getattr(want[k], f)), (k, f)
    pk.free_gpu_resources()


def test_msm_entry_count_guard(gpu):
    """a plain (not precomputed) table whose n * windows reaches 2^31 is refused with an error instead of overflowing the
    31-bit entry counts of the sort: 2^27 points x 16 windows, the check precedes every allocation"""
    import torch
    c = CURVES["bn254"]
    n = 1 << 27
    d_pts = torch.zeros((n, 8), dtype=torch.int64, device="cuda")          # 8 GiB of points at infinity
    torch.cuda.synchronize()
    t = gpu.Table(c.curve_id, 1, d_pts, precomp=False, n=n, on_device=True)
    del d_pts
    assert t.info()["n_windows"] * n >= 1 << 31
    d_sc = torch.zeros(4, dtype=torch.int64, device="cuda")
    with pytest.raises(gpu.B200Error, match="2\\^31"):
        t.msm(d_sc, n=n, on_device=True)
    # a sub-range below the limit still works (all bases are the point at infinity: the sum is infinity, Z = 0)
    d_sc = torch.zeros((1 << 20) * 4, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    out = t.msm(d_sc, n=1 << 20, on_device=True)
    assert not out[8:].any()
    t.free()


@pytest.mark.parametrize("cname", ["bn254", "bls12-381"])
def test_pedersen_commit_prove_knowledge_fold(gpu, cname):
    """backend/groth16/bn254/prove.go:72-129 through the ABI: per commitment Commit (:84) and ProveKnowledge (:114) as
    two MSMs over one upload, then Fold (:127).  Keys built like gnark-crypto's pedersen.Setup with a trapdoor
    (Basis_i = a_i G, BasisExpSigma_i = sigma a_i G): commitment = (sum v a) G, pok = sigma * commitment, and the
    verifier's relation of the folded proof  sum c^i pok_i = sum c^i sigma_i commitment_i.  Sizes as in
    backend/groth16/bn254/commitment_test.go (1 and 2 committed values) and two larger ones."""
    from oracle import corelib
    c = CURVES[cname]
    F = ff.Fp(c.p)
    rng = random.Random(91)
    r, L = c.r, c.fr_limbs
    pe = lambda v: ff.pack_elements(v, r, L)
    G = ec.pack_points(c, 1, [c.g1])
    sizes = (1, 2, 57, 20000)
    poks, want_fold, chal = [], None, rng.randrange(r)
    for i, n in enumerate(sizes):
        a = [rng.randrange(1, r) for _ in range(n)]
        sigma = rng.randrange(1, r)
        v = [rng.randrange(r) for _ in range(n)]
        if n > 2:
            v[3] = 0                                     # a zero value and a repeated basis element
            a[5] = a[4]
        basis = corelib.fixed_base(c, 1, G, pe(a))
        basis_sigma = corelib.fixed_base(c, 1, G, pe([x * sigma % r for x in a]))
        key = gpu.PedersenKey(c.curve_id, basis, basis_sigma)
        cm, pok = key.commit(pe(v))
        dl = sum(x * y for x, y in zip(v, a)) % r
        assert ec.unpack_points(c, 1, cm)[0] == ec.scalar_mul(F, dl, c.g1), n
        assert ec.unpack_points(c, 1, pok)[0] == ec.scalar_mul(F, dl * sigma % r, c.g1), n
        cm_only, none = key.commit(pe(v), want_pok=False)
        assert none is None and np.array_equal(cm_only, cm)
        with pytest.raises(gpu.B200Error):
            key.commit(pe(v + [1]))                      # length mismatch is an error, as in gnark-crypto
        key.free()
        poks.append(pok)
        want_fold = ec.affine_add(F, want_fold, ec.scalar_mul(F, pow(chal, i, r) * dl % r * sigma % r, c.g1))
    got = gpu.pedersen_fold(c.curve_id, np.concatenate(poks), pe([chal]))
    assert ec.unpack_points(c, 1, got)[0] == want_fold


@pytest.mark.parametrize("on_device", [False, True])
def test_msm_gather_is_the_filtered_msm(gpu, on_device):
    """b200_msm_gather: the wire vector is used as is, the table holds only the bases that are not infinity
    (prove.go:147-168): equals the MSM over the filtered copies, host and device operands, sub-range of the table"""
    import torch
    c = CURVES["bn254"]
    n_wires, n = 9000, 6000
    F, base, pts, sc, _ = known_dlog_instance(c, 1, n, seed=41)
    rng = random.Random(6)
    idx = np.array(sorted(rng.sample(range(n_wires), n)), dtype=np.uint32)
    wires = ff.pack_elements([rng.randrange(c.r) for _ in range(n_wires)], c.r, c.fr_limbs).reshape(n_wires, c.fr_limbs)
    t = gpu.Table(c.curve_id, 1, pts, precomp=True)
    want = t.msm(np.ascontiguousarray(wires[idx]), n=n)
    if on_device:
        d_idx = torch.from_numpy(idx.astype(np.int32)).cuda()
        d_w = torch.from_numpy(wires.view(np.int64).copy()).cuda()
        torch.cuda.synchronize()
        got = t.msm_gather(d_idx, d_w)
    else:
        got = t.msm_gather(idx, wires)
    assert jac_to_affine(c, 1, got) == jac_to_affine(c, 1, want)
    # a sub-range of the table with its own index list
    got2 = t.msm_gather(idx[100:1100], wires, off=100)
    assert jac_to_affine(c, 1, got2) == jac_to_affine(c, 1, t.msm(np.ascontiguousarray(wires[idx[100:1100]]), off=100, n=1000))
    with pytest.raises(gpu.B200Error):
        t.msm_gather(np.array([n_wires], dtype=np.uint32), wires)           # index out of range
    t.free()


def test_table_from_the_ethereum_srs_in_gnark_encoding(gpu):
    """b200_table_upload_encoded on EXTERNAL bytes: the 4096 compressed BLS12-381 G1 points of the Ethereum KZG ceremony
    file the reference ships (gnark-crypto's compressed encoding = the ZCash convention) are decoded on the device - one
    square root per point - and the table commits like the table built from the oracle-decoded points; the same points
    re-encoded uncompressed give the same table; a corrupted point is refused with its index."""
    from oracle import encoding, kzg_srs
    c = CURVES["bls12-381"]
    blob = open(kzg_srs.PATH, "rb").read()
    mono, _, _ = kzg_srs.load()
    n = len(mono)
    rng = random.Random(3)
    sc = ff.pack_elements([rng.randrange(c.r) for _ in range(n)], c.r, c.fr_limbs)
    want_t = gpu.Table(c.curve_id, 1, ec.pack_points(c, 1, mono), precomp=False)
    want = jac_to_affine(c, 1, want_t.msm(sc))
    want_t.free()
    t = gpu.Table.from_encoded(c.curve_id, 1, blob[:48 * n], n, gpu.POINTS_COMPRESSED, precomp=True)
    assert jac_to_affine(c, 1, t.msm(sc)) == want
    t.free()
    raw = b"".join(encoding.encode_g1(c, P_, False) for P_ in mono)
    t = gpu.Table.from_encoded(c.curve_id, 1, raw, n, gpu.POINTS_RAW, precomp=False)
    assert jac_to_affine(c, 1, t.msm(sc)) == want
    t.free()
    bad = bytearray(blob[:48 * n])
    x = int.from_bytes(blob[48 * 1234:48 * 1235], "big") & ((1 << 381) - 1)
    while pow((x ** 3 + 4) % c.p, (c.p - 1) // 2, c.p) == 1:        # the next x that is NOT on the curve
        x += 1
    bad[48 * 1234:48 * 1235] = (x | (0b100 << 381)).to_bytes(48, "big")
    with pytest.raises(gpu.B200Error, match="point 1234"):
        gpu.Table.from_encoded(c.curve_id, 1, bytes(bad), n, gpu.POINTS_COMPRESSED)


@pytest.mark.parametrize("cname", ["bn254", "bw6-761", "bls12-377"])
def test_table_from_encoded_points_other_curves(gpu, cname):
    """the decoder on the other curves against oracle/encoding.py: BN254's two-bit metadata, BW6-761's 96-byte
    coordinates, BLS12-377 compressed through Tonelli-Shanks; infinity inside the slice; G2 raw"""
    from oracle import encoding
    c = CURVES[cname]
    n = 300
    F, base, pts, sc, expected = known_dlog_instance(c, 1, n, seed=15)
    P1 = ec.unpack_points(c, 1, pts)
    on_curve = cname != "bw6-761"          # the oracle's BW6-761 base lives on another a = 0 curve (no b in the group law)
    if on_curve:
        for enc, compressed in ((gpu.POINTS_RAW, False), (gpu.POINTS_COMPRESSED, True)):
            data = b"".join(encoding.encode_g1(c, P_, compressed) for P_ in P1)
            t = gpu.Table.from_encoded(c.curve_id, 1, data, n, enc, precomp=False)
            assert jac_to_affine(c, 1, t.msm(sc)) == expected, (cname, enc)
            t.free()
    else:
        rng = random.Random(5)
        P1 = []
        while len(P1) < n:
            x = rng.randrange(c.p)
            y2 = (x ** 3 - 1) % c.p
            y = pow(y2, (c.p + 1) // 4, c.p)
            if y * y % c.p == y2:
                P1.append((x, y))
        P1[7] = None
        s1 = [rng.randrange(1 << 20) for _ in range(n)]
        want = ec.msm_naive(ff.Fp(c.p), P1, s1)
        for enc, compressed in ((gpu.POINTS_RAW, False), (gpu.POINTS_COMPRESSED, True)):
            data = b"".join(encoding.encode_g1(c, P_, compressed) for P_ in P1)
            t = gpu.Table.from_encoded(c.curve_id, 1, data, n, enc, precomp=False)
            assert jac_to_affine(c, 1, t.msm(ff.pack_elements(s1, c.r, c.fr_limbs))) == want, enc
            t.free()
    if c.fp2_nonresidue is not None:
        F2, base2, pts2, sc2, exp2 = known_dlog_instance(c, 2, 64, seed=16)
        data = b"".join(encoding.encode_g2_raw(c, Q) for Q in ec.unpack_points(c, 2, pts2))
        t = gpu.Table.from_encoded(c.curve_id, 2, data, 64, gpu.POINTS_RAW, precomp=False)
        assert jac_to_affine(c, 2, t.msm(sc2)) == exp2
        t.free()
        with pytest.raises(gpu.B200Error):
            gpu.Table.from_encoded(c.curve_id, 2, data, 64, gpu.POINTS_COMPRESSED)


@pytest.mark.parametrize("cname,logn,seed", [("bn254", 20, 2020), ("bn254", 20, 20), ("bls12-381", 16, 2020)])
def test_groth16_full_size_proof_verifies(gpu, cname, logn, seed):
    """BASELINE configs[2] at size: a SATISFIED 2^20 - 1 constraint circuit (oracle/groth16_fast.py; its small sizes are
    checked against the big-int oracle in tests/test_groth16_fast.py), trapdoor key built with b200_fixed_base_batch,
    one b200_groth16_prove with r, s injected.  Checked: each of the five MSM results and the three proof points equal
    dlog * generator (bit-exact affine points), and the reference's own acceptance test of a proof - Verify, the pairing
    equation on the proof points (backend/groth16/bn254/verify.go:38-140)."""
    from gnark_b200 import groth16 as b200
    from oracle import groth16_fast as gf
    c = CURVES[cname]
    # seed 20 at 2^20 is the instance whose G2.B table had one wrong point before the carry fix of mont_reduce_wide
    # (field.cuh; tests/test_emulation.py::test_g2_doubling_chain_known_failure)
    inst = gf.satisfied_instance(c, logn, seed=seed)
    assert gf.check_satisfied(inst)
    fb = lambda group, dl: gpu.fixed_base_batch(c.curve_id, group, ec.pack_points(c, group, [c.g1 if group == 1 else c.g2]),
                                                np.ascontiguousarray(dl))
    kp = gf.key_points(inst, fb)
    pk = b200.ProvingKey.from_arrays(c.curve_id, inst.n, kp["alpha"], kp["beta"], kp["delta"], kp["A"], kp["B"], kp["Z"],
                                     kp["K"], kp["beta2"], kp["delta2"], kp["B2"], inst.inf_a, inst.inf_b, inst.nb_public)
    a, b, cc = inst.solution_abc()
    sol = b200.R1CSSolution(W=inst.wires(), A=a, B=b, C=cc)
    rs = [0x1234567 << 100 | 0x89, 0xABCDEF << 90 | 0x77]
    it = iter(rs)
    proof = b200.ProveSolution(pk, sol, b200.WithDeviceID(0), b200.WithRandomness(lambda q: next(it)), keep_msm=True)
    e = gf.expected(inst, rs[0], rs[1])
    F1, F2 = ff.Fp(c.p), ff.base_field(c, 2)
    L = 3 * c.fp_limbs
    for k, dlog in enumerate((e.msm_a, e.msm_b, e.msm_z, e.msm_k)):
        got = ec.from_jac(F1, ec.unpack_points(c, 1, proof.msm[k * L:(k + 1) * L], ncoords=3)[0])
        assert got == ec.scalar_mul(F1, dlog, c.g1), ("msm", k)
    assert ec.from_jac(F2, ec.unpack_points(c, 2, proof.msm[4 * L:], ncoords=3)[0]) == ec.scalar_mul(F2, e.msm_b, c.g2)
    assert gf.verify_points(inst, ec.unpack_points(c, 1, proof.Ar)[0], ec.unpack_points(c, 2, proof.Bs)[0],
                            ec.unpack_points(c, 1, proof.Krs)[0], e, with_pairing=True)
    pk.free_gpu_resources()


@pytest.mark.parametrize("cname,logn", [("bn254", 6), ("bls12-381", 4), ("bn254", 12)])
def test_plonk_prove_statistical_zk(gpu, cname, logn):
    """backend.WithStatisticalZeroKnowledge through b200_plonk_prove (b200_plonk_challenges.hr: the two
    quotientShardsRandomizers, prove.go:239-242,689-722,1476-1481): at small sizes all ten digests and the opened values
    against the big-int oracle prover with the same randomisers, and Verify with real pairings; at 2^12 (no big-int
    prover) the verifier's equations of oracle/plonk_fast.py, plus: the randomisers change [H1..3] and the linearised
    digest and nothing else."""
    from oracle import corelib, plonk_prover as pp
    c = CURVES[cname]
    r, L = c.r, c.fr_limbs
    pe = lambda v: ff.pack_elements(v, r, L)
    rng = random.Random(7000 + logn)
    rnd = lambda: rng.randrange(r)
    if logn <= 8:
        n = 1 << logn
        circ, l, rr, o = pp.random_satisfied_instance(c, n, seed=logn + 60)
        ch = pp.Challenges(gamma=rnd(), beta=rnd(), alpha=rnd(), zeta=rnd(), v=rnd(), bl=[rnd(), rnd()], br=[rnd(), rnd()],
                           bo=[rnd(), rnd()], bz=[rnd(), rnd(), rnd()], hr=[rnd(), rnd()])
        tau = rnd()
        want = pp.prove(c, circ, l, rr, o, ch, tau)
        assert pp.verify(c, circ, want, ch, tau)
        srs = corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), pe([pow(tau, i, r) for i in range(n + 3)]))
        key = gpu.PlonkKey(c.curve_id, logn, pe(circ.ql), pe(circ.qr), pe(circ.qm), pe(circ.qo), pe(circ.qk),
                           np.array(circ.perm, dtype=np.int64), srs)
        pts, vals = key.prove(pe(l), pe(rr), pe(o), pe([ch.gamma]), pe([ch.beta]), pe([ch.alpha]), pe([ch.zeta]), pe([ch.v]),
                              pe(ch.bl), pe(ch.br), pe(ch.bo), pe(ch.bz), hr=pe(ch.hr))
        F = ff.Fp(c.p)
        dl = [want.L, want.R, want.O, want.Z, want.H[0], want.H[1], want.H[2], want.lin, want.batch_opening, want.z_opening]
        for k, name in enumerate(("L", "R", "O", "Z", "H1", "H2", "H3", "lin", "batch", "zopen")):
            assert jac_to_affine(c, 1, pts[k]) == ec.scalar_mul(F, dl[k], c.g1), name
        got_vals = ff.unpack_elements(vals, r, L)
        assert got_vals[:6] == want.claimed and got_vals[6] == want.zu
        assert pp.verify_pairing(c, circ, [jac_to_affine(c, 1, pts[k]) for k in range(10)], got_vals, ch, tau)
        key.free()
        return
    from oracle import plonk_fast
    inst = plonk_fast.satisfied_instance(c, logn, seed=logn)
    srs = plonk_fast.trapdoor_srs_gpu(gpu, c, inst)
    key = gpu.PlonkKey(c.curve_id, logn, inst.ql, inst.qr, inst.qm, inst.qo, inst.qk, inst.perm, srs)
    chp = inst.challenges_packed()
    plain_pts, plain_vals = key.prove(inst.l, inst.r, inst.o, *chp)
    pts, vals = key.prove(inst.l, inst.r, inst.o, *chp, hr=pe([rnd(), rnd()]))
    key.free()
    assert plonk_fast.verify(c, inst, plain_pts, plain_vals) and plonk_fast.verify(c, inst, pts, vals)
    assert np.array_equal(vals, plain_vals)
    same = [jac_to_affine(c, 1, pts[k]) == jac_to_affine(c, 1, plain_pts[k]) for k in range(10)]
    # L R O Z unchanged; H1 H2 H3 and the linearised digest differ; the batch opening differs (it folds the linearised
    # polynomial); the Z opening is unchanged
    assert same == [True, True, True, True, False, False, False, False, False, True]


def test_g2_table_point_known_failure(gpu):
    """the BN254 G2 point of tests/golden/bn254_g2_lost_carry_point.npy: its precomputed slabs 11..15 were wrong on the
    GPU before the carry fix of mont_reduce_wide (field.cuh) - every slab (s = 2^(16 w)) and the scalar it was paired
    with, on a precomputed and on a plain table, alone and behind other points"""
    import os
    c = CURVES["bn254"]
    F = ff.base_field(c, 2)
    here = os.path.dirname(os.path.abspath(__file__))
    pt = np.load(os.path.join(here, "golden", "bn254_g2_lost_carry_point.npy"))
    sc = np.load(os.path.join(here, "golden", "bn254_g2_lost_carry_scalar.npy"))
    P = ec.unpack_points(c, 2, pt)[0]
    s = ff.unpack_elements(sc, c.r, c.fr_limbs)[0]
    pe = lambda v: ff.pack_elements(v, c.r, c.fr_limbs)
    for pad in (0, 37):
        tab = np.ascontiguousarray(np.concatenate([ec.pack_points(c, 2, [c.g2] * pad), pt]) if pad else pt)
        for precomp in (True, False):
            t = gpu.Table(c.curve_id, 2, tab, precomp=precomp)
            assert jac_to_affine(c, 2, t.msm(sc, off=pad, n=1)) == ec.scalar_mul(F, s, P), (pad, precomp)
            for w in range(16):
                assert jac_to_affine(c, 2, t.msm(pe([1 << (16 * w)]), off=pad, n=1)) == ec.scalar_mul(F, 1 << (16 * w), P), (pad, precomp, w)
            t.free()


def test_table_from_compressed_g2(gpu):
    """compressed G2 slices decoded on the device (Fp2 square roots; points_decode.cuh): the 65 compressed BLS12-381 G2
    points of the Ethereum KZG ceremony file (EXTERNAL bytes; oracle/kzg_srs.py decodes them for the expected sum) and
    random multiples of the BN254 G2 generator, each as an MSM against the C++ oracle; a corrupted x is refused."""
    from oracle import corelib, encoding, kzg_srs
    rng = random.Random(31)
    c = CURVES["bls12-381"]
    blob = open(kzg_srs.PATH, "rb").read()
    off = 2 * kzg_srs.N * 48
    n = kzg_srs.N_G2
    data = blob[off:off + 96 * n]
    want_pts = kzg_srs.load()[2]
    sc = ff.pack_elements([rng.randrange(c.r) for _ in range(n)], c.r, c.fr_limbs)
    want = jac_to_affine(c, 2, corelib.msm(c, 2, ec.pack_points(c, 2, want_pts), sc))
    for precomp in (False, True):
        t = gpu.Table.from_encoded(c.curve_id, 2, data, n, gpu.POINTS_COMPRESSED, precomp=precomp)
        assert jac_to_affine(c, 2, t.msm(sc)) == want
        t.free()
    # an x that is not on the twist, at position 9: refused with its index
    F2 = ff.base_field(c, 2)
    bad = bytearray(data)
    x0 = int.from_bytes(data[96 * 9 + 48:96 * 10], "big")
    x1 = int.from_bytes(data[96 * 9:96 * 9 + 48], "big") & ((1 << 381) - 1)
    while True:
        x0 = (x0 + 1) % c.p
        if encoding.sqrt_fp2(c, F2.add(F2.mul(F2.sqr((x0, x1)), (x0, x1)), encoding.twist_b(c))) is None:
            break
    bad[96 * 9 + 48:96 * 10] = x0.to_bytes(48, "big")
    with pytest.raises(gpu.B200Error, match="point 9"):
        gpu.Table.from_encoded(c.curve_id, 2, bytes(bad), n, gpu.POINTS_COMPRESSED)
    c = CURVES["bn254"]
    F2 = ff.base_field(c, 2)
    m = 200
    pts = [ec.scalar_mul(F2, rng.randrange(1, c.r), c.g2) for _ in range(m)]
    pts[11] = None
    data = b"".join(encoding.encode_g2(c, Q, True) for Q in pts)
    sc = ff.pack_elements([rng.randrange(c.r) for _ in range(m)], c.r, c.fr_limbs)
    want = jac_to_affine(c, 2, corelib.msm(c, 2, ec.pack_points(c, 2, pts), sc))
    t = gpu.Table.from_encoded(c.curve_id, 2, data, m, gpu.POINTS_COMPRESSED, precomp=True)
    assert jac_to_affine(c, 2, t.msm(sc)) == want
    t.free()
