"""Radix-2 NTT over Fr in gnark-crypto's fft.Domain conventions.
TEST INFRASTRUCTURE ONLY (see oracle/params.py header for what pins parity and what does not).

gnark-crypto v0.21.0 source is absent; the conventions restated here are the
ones the reference's call sites rely on:
  * fft.NewDomain(m): Cardinality = nextpow2(m), Generator = primitive root of
    that order, FrMultiplicativeGen = coset shift
    (backend/groth16/bn254/setup.go:101, backend/plonk/bn254/setup.go:120-123).
  * FFT/FFTInverse(a, DIF): natural in -> bit-reversed out; (a,b)->(a+b,(a-b)*tw)
    with twiddles[stage][i] = w^(i*2^stage); recursion shape is the in-tree
    group-element twin backend/groth16/bn254/mpcsetup/lagrange.go:132-169.
  * FFT(a, DIT): bit-reversed in -> natural out (prove.go:362-368 feeds DIF
    output to DIT with no permutation).
  * FFTInverse uses w^-1 and scales by 1/n
    (test/unsafekzg/kzgsrs.go:186-194 pins iFFT-DIF + BitReverse = Lagrange basis).
  * OnCoset: forward multiplies coefficient j by g^j before the transform,
    inverse multiplies coefficient j by g^-j after it (prove.go:366-386).
"""

DIF = 0
DIT = 1


def bitrev(i: int, logn: int) -> int:
    r = 0
    for _ in range(logn):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def bit_reverse(a):
    """fft.BitReverse — in-tree generic twin: backend/groth16/bn254/setup.go:671-682."""
    n = len(a)
    logn = n.bit_length() - 1
    out = list(a)
    for i in range(n):
        j = bitrev(i, logn)
        if j > i:
            out[i], out[j] = out[j], out[i]
    return out


class Domain:
    def __init__(self, curve, m: int, generator: int = None, coset_gen: int = None):
        self.r = curve.r
        n = 1
        while n < m:
            n <<= 1
        self.n = n
        self.logn = n.bit_length() - 1
        if self.logn > curve.two_adicity:
            raise ValueError("domain too large for the field's 2-adicity")
        if generator is None:
            generator = pow(curve.root_of_unity, 1 << (curve.two_adicity - self.logn), self.r)
        self.generator = generator
        self.generator_inv = pow(generator, -1, self.r)
        self.cardinality_inv = pow(n, -1, self.r)
        self.coset_gen = curve.mult_gen if coset_gen is None else coset_gen
        self.coset_gen_inv = pow(self.coset_gen, -1, self.r)

    # -- core in-place transforms on python lists of ints ------------------
    def _dif(self, a, w):
        r = self.r
        n = len(a)
        m = n >> 1
        stage_w = w
        while m >= 1:
            # twiddles for this stage: stage_w^i, i < m
            tw = [1] * m
            for i in range(1, m):
                tw[i] = tw[i - 1] * stage_w % r
            for start in range(0, n, 2 * m):
                for i in range(m):
                    x = a[start + i]
                    y = a[start + i + m]
                    a[start + i] = (x + y) % r
                    a[start + i + m] = (x - y) * tw[i] % r
            stage_w = stage_w * stage_w % r
            m >>= 1

    def _dit(self, a, w):
        r = self.r
        n = len(a)
        m = 1
        while m < n:
            stage_w = pow(w, n // (2 * m), r)
            tw = [1] * m
            for i in range(1, m):
                tw[i] = tw[i - 1] * stage_w % r
            for start in range(0, n, 2 * m):
                for i in range(m):
                    x = a[start + i]
                    y = a[start + i + m] * tw[i] % r
                    a[start + i] = (x + y) % r
                    a[start + i + m] = (x - y) % r
            m <<= 1

    def fft(self, a, decimation, on_coset=False):
        assert len(a) == self.n
        a = [x % self.r for x in a]
        r = self.r
        if on_coset:
            g = self.coset_gen
            if decimation == DIT:   # input is bit-reversed: position i holds coefficient bitrev(i)
                for i in range(self.n):
                    a[i] = a[i] * pow(g, bitrev(i, self.logn), r) % r
            else:
                for i in range(self.n):
                    a[i] = a[i] * pow(g, i, r) % r
        if decimation == DIF:
            self._dif(a, self.generator)
        else:
            self._dit(a, self.generator)
        return a

    def fft_inverse(self, a, decimation, on_coset=False):
        assert len(a) == self.n
        a = [x % self.r for x in a]
        r = self.r
        if decimation == DIF:
            self._dif(a, self.generator_inv)
        else:
            self._dit(a, self.generator_inv)
        ninv = self.cardinality_inv
        if on_coset:
            gi = self.coset_gen_inv
            for i in range(self.n):
                j = bitrev(i, self.logn) if decimation == DIF else i
                a[i] = a[i] * ninv % r * pow(gi, j, r) % r
        else:
            for i in range(self.n):
                a[i] = a[i] * ninv % r
        return a


def dft_naive(curve, coeffs, w, shift=1):
    """evaluations of sum_j c_j X^j at shift*w^k, k = 0..n-1 (natural order), O(n^2)."""
    r = curve.r
    n = len(coeffs)
    out = []
    for k in range(n):
        x = shift * pow(w, k, r) % r
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + c) % r
        out.append(acc)
    return out


def poly_eval(r, coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % r
    return acc
