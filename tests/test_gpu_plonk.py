"""PLONK building blocks on the GPU vs the oracle: element-wise vector ops, batch inversion, the
fused constraint evaluation of computeNumerator (one call per coset, inputs moved to the coset with
b200_ntt exactly as backend/plonk/bn254/prove.go:1035-1057 does), divideByZH, and a KZG commitment
against a trapdoor SRS (test/unsafekzg pattern: commit(p) = p(tau) * G)."""
import random

import numpy as np
import pytest
import torch

from oracle import ec, ff, ntt, plonk
from oracle.params import CURVES
from util import jac_to_affine

pytestmark = pytest.mark.gpu
ALL = list(CURVES.values())


def dev(arr):
    return torch.from_numpy(arr.view(np.int64)).cuda()


def host(t, limbs):
    return t.cpu().numpy().view(np.uint64).reshape(-1, limbs)


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_vec_ops(gpu, c):
    rng = random.Random(3)
    n, L, r = 1000, c.fr_limbs, c.r
    a = [rng.randrange(r) for _ in range(n)]
    b = [rng.randrange(r) for _ in range(n)]
    da, db = dev(ff.pack_elements(a, r, L)), dev(ff.pack_elements(b, r, L))
    out = torch.zeros_like(da)
    for op, f in ((gpu.VEC_MUL, lambda x, y: x * y % r), (gpu.VEC_ADD, lambda x, y: (x + y) % r),
                  (gpu.VEC_SUB, lambda x, y: (x - y) % r)):
        gpu.vec_op(0, c.curve_id, op, out, da, db, n)
        gpu.sync(0)
        assert ff.unpack_elements(host(out, L), r, L) == [f(x, y) for x, y in zip(a, b)]
    # batch inversion (zeros stay zero)
    a[5] = 0
    da = dev(ff.pack_elements(a, r, L))
    gpu.vec_batch_invert(0, c.curve_id, da, n)
    gpu.sync(0)
    assert ff.unpack_elements(host(da, L), r, L) == [pow(x, -1, r) if x else 0 for x in a]
    # bit reverse + scale by powers
    m = 1 << 9
    v = [rng.randrange(r) for _ in range(m)]
    dv = dev(ff.pack_elements(v, r, L))
    gpu.vec_bit_reverse(0, c.curve_id, dv, 9)
    gpu.sync(0)
    assert ff.unpack_elements(host(dv, L), r, L) == ntt.bit_reverse(v)
    s, g = rng.randrange(r), rng.randrange(r)
    dv = dev(ff.pack_elements(v, r, L))
    gpu.vec_scale_powers(0, c.curve_id, dv, m, ff.pack_elements([s], r, L), ff.pack_elements([g], r, L))
    assert ff.unpack_elements(host(dv, L), r, L) == [x * s * pow(g, i, r) % r for i, x in enumerate(v)]


@pytest.mark.parametrize("c", [CURVES["bn254"], CURVES["bls12-381"]], ids=lambda c: c.name)
def test_quotient_pipeline(gpu, c):
    r, L = c.r, c.fr_limbs
    n, rho = 64, 4
    logn = 6
    rng = random.Random(17)
    dom0_o, dom1_o = ntt.Domain(c, n), ntt.Domain(c, rho * n)
    polys = {k: [rng.randrange(r) for _ in range(n)] for k in plonk.POLYS}
    alpha, beta, gamma = (rng.randrange(r) for _ in range(3))
    blind = {"l": [rng.randrange(r) for _ in range(2)], "r": [rng.randrange(r) for _ in range(2)],
             "o": [rng.randrange(r) for _ in range(2)], "z": [rng.randrange(r) for _ in range(3)]}
    want_cres = plonk.numerator(c, n, rho, polys, alpha, beta, gamma, blind)
    want_h = plonk.divide_by_zh(c, n, rho, want_cres)

    pe = lambda v: ff.pack_elements(v, r, L)
    g, w4 = dom1_o.coset_gen, dom1_o.generator
    out = torch.zeros(rho * n * L, dtype=torch.int64, device="cuda")
    for i in range(rho):
        coset = g * pow(w4, i, r) % r
        # the reference's per-coset move: iFFT (DIF) -> scale by coset^k -> FFT (DIT) = FFT on the coset
        d0 = gpu.Domain(c.curve_id, logn, coset_gen=pe([coset]))
        on_coset = {}
        for k in plonk.POLYS:
            t = dev(pe(polys[k]))
            d0.ntt_async(t, inverse=True, decimation=gpu.DIF)
            d0.ntt_async(t, inverse=False, decimation=gpu.DIT, on_coset=True)
            on_coset[k] = t
        gpu.plonk_constraints_coset(d0, pe([g]), pe([w4]), on_coset, pe([alpha]), pe([beta]), pe([gamma]),
                                    {k: pe(v) for k, v in blind.items()}, i, rho, out)
        gpu.sync(0)
        d0.free()
    assert ff.unpack_elements(host(out, L), r, L) == want_cres
    d1 = gpu.Domain(c.curve_id, logn + 2)
    gpu.plonk_divide_by_zh(d1, logn, out)
    assert ff.unpack_elements(host(out, L), r, L) == want_h
    d1.free()


def test_kzg_commit_trapdoor_srs(gpu):
    """kzg.Commit = MultiExp(pk.G1[:len(p)], p) (plonk/bn254/prove.go:300,532,788); with a
    known tau the digest must be p(tau) * G (test/unsafekzg/kzgsrs.go:142-172), BLS12-381 (config 4)."""
    from oracle import corelib
    c = CURVES["bls12-381"]
    rng = random.Random(23)
    n = 1 << 12
    tau = rng.randrange(c.r)
    srs_dlogs = [pow(tau, i, c.r) for i in range(n)]
    SRS = corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), ff.pack_elements(srs_dlogs, c.r, c.fr_limbs))
    p = [rng.randrange(c.r) for _ in range(n)]
    t = gpu.Table(c.curve_id, 1, SRS, precomp=True)
    digest = jac_to_affine(c, 1, t.msm(ff.pack_elements(p, c.r, c.fr_limbs)))
    assert digest == ec.scalar_mul(ff.Fp(c.p), ntt.poly_eval(c.r, p, tau), c.g1)
    t.free()


@pytest.mark.parametrize("c", [CURVES["bn254"], CURVES["bls12-381"], CURVES["bw6-761"]], ids=lambda c: c.name)
def test_scans_eval_division(gpu, c):
    rng = random.Random(41)
    r, L = c.r, c.fr_limbs
    pe = lambda v: ff.pack_elements(v, r, L)
    for n in (1, 5, 1024, 1025, 5000):          # below / at / across block boundaries (1024 per block)
        v = [rng.randrange(1, r) for _ in range(n)]
        for op, f, ident in ((gpu.SCAN_PRODUCT, lambda a, b: a * b % r, 1), (gpu.SCAN_SUM, lambda a, b: (a + b) % r, 0)):
            for excl in (False, True):
                d = dev(pe(v))
                gpu.vec_scan(0, c.curve_id, op, d, n, exclusive=excl)
                gpu.sync(0)
                acc, want = ident, []
                for x in v:
                    if excl:
                        want.append(acc)
                    acc = f(acc, x)
                    if not excl:
                        want.append(acc)
                assert ff.unpack_elements(host(d, L), r, L) == want, (c.name, n, op, excl)
        # Horner evaluation and division by (X - z)
        x = rng.randrange(r)
        d = dev(pe(v))
        assert ff.unpack_elements(gpu.poly_eval(0, c.curve_id, d, n, pe([x])), r, L)[0] == ntt.poly_eval(r, v, x)
        for z in (x, 0):
            d = dev(pe(v))
            rem = ff.unpack_elements(gpu.poly_div_by_linear(0, c.curve_id, d, n, pe([z])), r, L)[0]
            q, want_rem = plonk.div_by_linear(r, v, z)
            assert rem == want_rem
            assert ff.unpack_elements(host(d, L), r, L) == q + [0]


@pytest.mark.parametrize("c", [CURVES["bn254"], CURVES["bls12-381"]], ids=lambda c: c.name)
def test_build_z_grand_product(gpu, c):
    """iop.BuildRatioCopyConstraint on device vs the oracle; for a permutation that really links equal
    wires the product telescopes back to 1 (Z[n-1] * ratio[n-1] = 1)."""
    rng = random.Random(43)
    r, L = c.r, c.fr_limbs
    logn = 7
    n = 1 << logn
    dom0 = ntt.Domain(c, n)
    pe = lambda v: ff.pack_elements(v, r, L)
    # wires: 3n slots; build a random permutation made of cycles and assign equal values along each cycle
    slots = list(range(3 * n))
    rng.shuffle(slots)
    perm = list(range(3 * n))
    vals = [0] * (3 * n)
    i = 0
    while i < 3 * n:
        k = min(rng.choice((1, 1, 2, 3, 5)), 3 * n - i)
        cyc = slots[i:i + k]
        v = rng.randrange(r)
        for a, b in zip(cyc, cyc[1:] + cyc[:1]):
            perm[a] = b
            vals[a] = v
        i += k
    l, rr, o = vals[:n], vals[n:2 * n], vals[2 * n:]
    beta, gamma = rng.randrange(r), rng.randrange(r)
    want = plonk.build_ratio_copy_constraint(c, dom0, l, rr, o, perm, beta, gamma)
    d0 = gpu.Domain(c.curve_id, logn)
    dz = torch.zeros(n * L, dtype=torch.int64, device="cuda")
    dperm = torch.tensor(perm, dtype=torch.int64, device="cuda")
    gpu.plonk_build_z(d0, dev(pe(l)), dev(pe(rr)), dev(pe(o)), dperm, pe([beta]), pe([gamma]), dz)
    gpu.sync(0)
    got = ff.unpack_elements(host(dz, L), r, L)
    assert got == want
    # telescoping: the last ratio brings the product back to one
    supp = plonk.support_permutation(c, dom0)
    num = den = 1
    for j, f in enumerate((l, rr, o)):
        num = num * ((f[n - 1] + beta * supp[j * n + n - 1] + gamma) % r) % r
        den = den * ((f[n - 1] + beta * supp[perm[j * n + n - 1]] + gamma) % r) % r
    assert got[n - 1] * num % r * pow(den, -1, r) % r == 1
    d0.free()
