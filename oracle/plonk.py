"""PLONK quotient pieces restated on big ints.  TEST INFRASTRUCTURE ONLY
(see oracle/params.py header for what pins parity and what does not).

Follows backend/plonk/bn254/prove.go:
  computeNumerator :841-1123  (gateConstraint :871-889, orderingConstraint :907-931,
                               localConstraint :933-941 with computeLagrangeOneOnCoset :380-385,
                               allConstraints :961-988, coset loop :996-1088, scatter :1070-1076)
  divideByZH :1287-1324, evaluateXnMinusOneDomainBigCoset :1327-1350, batchInvert :1130-1143.
BSB22 commitment gates (:881-884): gate += sum_i Qcp_i * PI2_i, passed as `bsb22` = [(qcp_lagrange, pi2_lagrange), ...].

The reference moves the 12 polynomials from coset to coset incrementally (iFFT, scale by the
shifter powers, FFT) and pre-scales the blinding polynomials' coefficients; mathematically every
coset iteration evaluates, at the points x_j = coset_i * w^j:
    P(x_j)                              for the 12 proving-key / witness polynomials
    P(x_j) + (x_j^n - 1) * b_P(x_j)     for the blinded L, R, O, Z
which is what is restated here (from canonical coefficients).
"""

from .ntt import DIF, DIT, Domain, bit_reverse, bitrev, poly_eval

POLYS = ("l", "r", "o", "z", "s1", "s2", "s3", "ql", "qr", "qm", "qo", "qk")


def all_constraints(r, n, ninv, u, x, x_next_blind, alpha, beta, gamma, cs, blind, xn_minus_one):
    """u: dict of the 12 values at x plus 'zs' (z at w*x); blind: dict name -> coeff list."""
    def b(name, at):
        return xn_minus_one * poly_eval(r, blind.get(name, []), at) % r
    L = (u["l"] + b("l", x)) % r
    R = (u["r"] + b("r", x)) % r
    O = (u["o"] + b("o", x)) % r
    Z = (u["z"] + b("z", x)) % r
    ZS = (u["zs"] + b("z", x_next_blind)) % r
    gate = (u["ql"] * L + u["qr"] * R + u["qm"] * L * R + u["qo"] * O + u["qk"]) % r
    idv = x * beta % r
    a = (gamma + L + idv) % r
    bb = (idv * cs + R + gamma) % r
    c = (idv * cs * cs + O + gamma) % r
    rr = a * bb * c * Z % r
    a = (u["s1"] * beta + L + gamma) % r
    bb = (u["s2"] * beta + R + gamma) % r
    c = (u["s3"] * beta + O + gamma) % r
    ordering = (a * bb * c * ZS - rr) % r
    lone = xn_minus_one * ninv % r * pow((x - 1) % r, -1, r) % r
    local = (Z - 1) * lone % r
    return ((local * alpha + ordering) * alpha + gate) % r


def coset_values(curve, dom0: Domain, lagrange, coset):
    """Lagrange/Regular values on <w>  ->  values on coset*<w> (natural order)."""
    r = curve.r
    coeffs = bit_reverse(dom0.fft_inverse(lagrange, DIF))        # canonical, regular
    return [poly_eval(r, coeffs, coset * pow(dom0.generator, j, r) % r) for j in range(dom0.n)]


def numerator(curve, n, rho, polys_lagrange, alpha, beta, gamma, blind, bsb22=()):
    """cres (LagrangeCoset on the big domain, BitReverse layout), length rho*n."""
    r = curve.r
    dom0 = Domain(curve, n)
    dom1 = Domain(curve, rho * n)
    g, w4, w = dom1.coset_gen, dom1.generator, dom0.generator
    logm = (rho * n).bit_length() - 1
    cres = [0] * (rho * n)
    for i in range(rho):
        coset = g * pow(w4, i, r) % r
        vals = {k: coset_values(curve, dom0, polys_lagrange[k], coset) for k in POLYS}
        extra = [(coset_values(curve, dom0, qcp, coset), coset_values(curve, dom0, pi2, coset)) for qcp, pi2 in bsb22]
        xn1 = (pow(coset, n, r) - 1) % r
        for j in range(n):
            x = coset * pow(w, j, r) % r
            x1 = coset * pow(w, (j + 1) % n, r) % r
            u = {k: vals[k][j] for k in POLYS}
            u["zs"] = vals["z"][(j + 1) % n]
            v = all_constraints(r, n, dom0.cardinality_inv, u, x, x1, alpha, beta, gamma, g, blind, xn1)
            for qv, pv in extra:                      # the gate term enters the sum with coefficient 1
                v = (v + qv[j] * pv[j]) % r
            cres[bitrev(rho * j + i, logm)] = v
    return cres


def divide_by_zh(curve, n, rho, cres):
    """:1287-1324 -> canonical, regular (length rho*n)."""
    r = curve.r
    dom1 = Domain(curve, rho * n)
    m = rho * n
    logm = m.bit_length() - 1
    gn = pow(dom1.coset_gen, n, r)
    wn = pow(dom1.generator, n, r)
    tab = [pow((gn * pow(wn, i, r) - 1) % r, -1, r) for i in range(rho)]
    scaled = [cres[i] * tab[bitrev(i, logm) % rho] % r for i in range(m)]
    return dom1.fft_inverse(scaled, DIT, on_coset=True)


def support_permutation(curve, dom0: Domain):
    """getSupportPermutation (backend/plonk/bn254/setup.go:377-392): <w> || g<w> || g^2<w>."""
    r, n, g = curve.r, dom0.n, dom0.coset_gen
    w = [pow(dom0.generator, i, r) for i in range(n)]
    return w + [g * x % r for x in w] + [g * g * x % r for x in w]


def build_ratio_copy_constraint(curve, dom0: Domain, l, rr, o, perm, beta, gamma):
    """iop.BuildRatioCopyConstraint as relied upon at plonk/bn254/prove.go:645-656:
    Z[0] = 1, Z[i+1] = Z[i] * prod_j (f_j[i] + beta*id_j[i] + gamma) / (f_j[i] + beta*supp[S[j n + i]] + gamma)."""
    r, n = curve.r, dom0.n
    supp = support_permutation(curve, dom0)
    f = (l, rr, o)
    z = [1] * n
    for i in range(n - 1):
        num = den = 1
        for j in range(3):
            num = num * ((f[j][i] + beta * supp[j * n + i] + gamma) % r) % r
            den = den * ((f[j][i] + beta * supp[perm[j * n + i]] + gamma) % r) % r
        z[i + 1] = z[i] * num % r * pow(den, -1, r) % r
    return z


def div_by_linear(r, coeffs, z):
    """(p(X) - p(z)) / (X - z) by synthetic division; returns (quotient (len n-1), p(z))."""
    n = len(coeffs)
    q = [0] * max(n - 1, 0)
    acc = 0
    for i in range(n - 1, 0, -1):
        acc = (coeffs[i] + z * acc) % r
        q[i - 1] = acc
    rem = (coeffs[0] + z * acc) % r if n else 0
    return q, rem
