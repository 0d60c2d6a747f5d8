"""NTT / computeH parity through the C ABI on the GPU (bit-exact field vectors)."""
import random

import numpy as np
import pytest

from oracle import corelib, ff, ntt
from oracle import groth16 as g16
from oracle.params import CURVES

pytestmark = pytest.mark.gpu
ALL = list(CURVES.values())


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_all_modes_small(gpu, c):
    rng = random.Random(21)
    for logn in (0, 1, 2, 5, 10):
        n = 1 << logn
        dom_o = ntt.Domain(c, n)
        a = [rng.randrange(c.r) for _ in range(n)]
        A0 = ff.pack_elements(a, c.r, c.fr_limbs)
        d = gpu.Domain(c.curve_id, logn)
        for inv in (False, True):
            for dec in (gpu.DIF, gpu.DIT):
                for cos in (False, True):
                    got = d.ntt(A0.copy(), inverse=inv, decimation=dec, on_coset=cos)
                    exp = (dom_o.fft_inverse if inv else dom_o.fft)(a, dec, on_coset=cos)
                    assert ff.unpack_elements(got, c.r, c.fr_limbs) == exp, (c.name, logn, inv, dec, cos)
        d.free()


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
@pytest.mark.parametrize("logn", (11, 12, 13, 16, 20))
def test_vs_cpp_oracle(gpu, c, logn):
    """multi-pass plans (11: single tile; 12,13: tiny second pass; 16, 20: two full passes)."""
    if c.fr_limbs > 4 and logn > 16:
        logn = 16
    n = 1 << logn
    rs = np.random.RandomState(logn)
    # random Montgomery residues < r: take random 64-bit limbs and reduce via the oracle's ntt-free path
    a = rs.randint(0, 1 << 62, size=(n, c.fr_limbs), dtype=np.int64).astype(np.uint64)
    a[:, -1] &= np.uint64((1 << (c.r.bit_length() - 64 * (c.fr_limbs - 1) - 1)) - 1)   # < r
    d = gpu.Domain(c.curve_id, logn)
    for inv, dec, cos in ((False, gpu.DIF, False), (False, gpu.DIT, True), (True, gpu.DIF, True), (True, gpu.DIT, False)):
        want = corelib.ntt(c, a.copy(), logn, inv, dec, cos)
        got = d.ntt(a.copy(), inverse=inv, decimation=dec, on_coset=cos)
        assert np.array_equal(got, want), (c.name, logn, inv, dec, cos)
    # round trip: iFFT(DIF) then FFT(DIT) is the identity, no permutation needed (prove.go:362-368)
    x = d.ntt(a.copy(), inverse=True, decimation=gpu.DIF)
    x = d.ntt(x, inverse=False, decimation=gpu.DIT)
    assert np.array_equal(x, a)
    d.free()


def test_custom_generator_and_coset(gpu):
    """ICICLE's path uses w_2n as coset generator (icicle.go:96,127-132): the domain takes both."""
    c = CURVES["bn254"]
    logn = 8
    n = 1 << logn
    rng = random.Random(3)
    w2n = pow(c.root_of_unity, 1 << (c.two_adicity - logn - 1), c.r)
    gen = pow(w2n, 2, c.r)
    a = [rng.randrange(c.r) for _ in range(n)]
    dom_o = ntt.Domain(c, n, generator=gen, coset_gen=w2n)
    d = gpu.Domain(c.curve_id, logn, generator=ff.pack_elements([gen], c.r, 4), coset_gen=ff.pack_elements([w2n], c.r, 4))
    got = d.ntt(ff.pack_elements(a, c.r, 4), inverse=False, decimation=gpu.DIT, on_coset=True)
    assert ff.unpack_elements(got, c.r, 4) == dom_o.fft(a, ntt.DIT, on_coset=True)
    d.free()


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_compute_h(gpu, c):
    """computeH vs the oracle pipeline, incl. zero padding; satisfied instance => top coeff 0."""
    m = 300
    cs = g16.square_chain_r1cs(m)
    W = g16.square_chain_witness(c.r, m)
    A, B, C = g16.solve_abc(cs, W, c.r)
    dom_o = ntt.Domain(c, m)
    want = g16.compute_h(dom_o, A, B, C)
    d = gpu.Domain(c.curve_id, dom_o.logn)
    got = d.compute_h(*(ff.pack_elements(v, c.r, c.fr_limbs) for v in (A, B, C)))
    assert ff.unpack_elements(got, c.r, c.fr_limbs) == want
    assert want[dom_o.n - 1] == 0
    # unsatisfied random a,b,c: output depends on the coset, must still match the CPU convention
    rng = random.Random(1)
    a, b, cc = ([rng.randrange(c.r) for _ in range(dom_o.n)] for _ in range(3))
    want = g16.compute_h(dom_o, a, b, cc)
    got = d.compute_h(*(ff.pack_elements(v, c.r, c.fr_limbs) for v in (a, b, cc)))
    assert ff.unpack_elements(got, c.r, c.fr_limbs) == want
    d.free()


def test_compute_h_baseline_size(gpu):
    """n = 2^20 (BASELINE config 3): bit-exact vs the C++ oracle + quotient identity at a random point."""
    c = CURVES["bn254"]
    logn = 20
    n = 1 << logn
    rs = np.random.RandomState(7)
    def rnd():
        a = rs.randint(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64)
        a[:, -1] &= np.uint64((1 << 59) - 1)
        return a
    a, b, cc = rnd(), rnd(), rnd()
    d = gpu.Domain(c.curve_id, logn)
    got = d.compute_h(a, b, cc)
    want = corelib.compute_h(c, a.copy(), b.copy(), cc.copy(), logn)
    assert np.array_equal(got, want)
    d.free()
