"""gnark-crypto's serialised point encodings, restated.  TEST INFRASTRUCTURE ONLY (see oracle/params.py header).

What `G1Affine.Bytes()` (compressed) / `RawBytes()` (uncompressed) and `G2Affine.RawBytes()` write and `SetBytes` reads
(gnark-crypto v0.21.0, ecc/<curve>/marshal.go; the module is absent here, so this follows its published layout):
big-endian canonical coordinates, metadata in the most significant bits of the first byte -

    BN254 (2 bits):   00 uncompressed | 10 compressed, y smallest | 11 compressed, y largest | 01 compressed infinity
    BLS12-381 / BLS12-377 / BW6-761 (3 bits):
                      000 uncompressed | 010 uncompressed infinity | 100 / 101 compressed, y smallest / largest |
                      110 compressed infinity
    "largest" = y > (p - 1) / 2 on canonical values (an Fp2 element: decided by A1 unless A1 = 0, then by A0);
    Fp2 elements are written A1 || A0.  G2 lives on the twist y^2 = x^3 + b': BN254 3/(9+u), BLS12-381 4(1+u),
    BLS12-377 1/u (u^2 = -5), BW6-761 (over Fp) 4.

The reference consumes these encodings wherever it serialises keys: backend/plonk/bn254/marshal.go:96-129 (pk.Kzg,
pk.KzgLagrange), backend/groth16/bn254/marshal.go:136-214.  PINNED by reference-held bytes: the compressed G1 generators
inside gnark's serialised verifying keys (backend/solidity/testdata/blank_plonk_{bn254,bls12381}_*.vk: BN254 0x80..01,
BLS12-381 0x97f1d3a7...) and the 8192 compressed BLS12-381 G1 points of std/evmprecompiles/kzg_trusted_setup.json
(tests/test_golden_kzg.py::test_encoding_matches_reference_bytes); G2: the compressed G2 generators of the same verifying
keys and the 65 compressed BLS12-381 G2 points of the ceremony file (tests/test_emulation.py::
test_g2_compressed_decoding).  BLS12-377 / BW6-761: same scheme, unpinned.
"""

from . import ff


def _flag_bits(curve):
    return 2 if curve.name == "bn254" else 3


def _fp_bytes(curve):
    return 8 * curve.fp_limbs


def _b(curve):
    return {"bn254": 3, "bls12-381": 4, "bls12-377": 1, "bw6-761": curve.p - 1}[curve.name]


def encode_g1(curve, P, compressed: bool, check=True) -> bytes:
    nb, fb = _fp_bytes(curve), _flag_bits(curve)
    sh = 8 - fb
    if P is None:
        first = (0b01 if compressed else 0b00) if fb == 2 else (0b110 if compressed else 0b010)
        body = bytearray(nb if compressed else 2 * nb)
        body[0] |= first << sh
        return bytes(body)
    x, y = P
    if check:
        assert (y * y - x * x * x - _b(curve)) % curve.p == 0
    xb = bytearray(x.to_bytes(nb, "big"))
    assert xb[0] >> sh == 0
    if not compressed:
        return bytes(xb) + y.to_bytes(nb, "big")          # flag bits 0
    largest = y > (curve.p - 1) // 2
    flag = (0b11 if largest else 0b10) if fb == 2 else (0b101 if largest else 0b100)
    xb[0] |= flag << sh
    return bytes(xb)


def decode_g1(curve, b: bytes):
    nb, fb = _fp_bytes(curve), _flag_bits(curve)
    sh = 8 - fb
    flag = b[0] >> sh
    mask = (1 << (8 * nb - fb)) - 1
    if fb == 2:
        compressed, inf, largest = flag != 0, flag == 0b01, flag == 0b11
    else:
        assert flag in (0b000, 0b010, 0b100, 0b101, 0b110)
        compressed, inf, largest = bool(flag & 0b100), flag in (0b010, 0b110), flag == 0b101
    assert len(b) == (nb if compressed else 2 * nb)
    if inf:
        return None
    x = int.from_bytes(b[:nb], "big") & mask
    assert x < curve.p
    if not compressed:
        y = int.from_bytes(b[nb:], "big")
        if x == 0 and y == 0:
            return None                                    # BN254 writes infinity as zeros
        assert y < curve.p and (y * y - x * x * x - _b(curve)) % curve.p == 0
        return (x, y)
    y2 = (x * x * x + _b(curve)) % curve.p
    y = sqrt_fp(curve.p, y2)
    assert y is not None, "x is not on the curve"
    if (y > (curve.p - 1) // 2) != largest:
        y = curve.p - y
    return (x, y)


def encode_g2_raw(curve, Q) -> bytes:
    """uncompressed G2 over Fp2: X.A1 || X.A0 || Y.A1 || Y.A0"""
    nb, fb = _fp_bytes(curve), _flag_bits(curve)
    if Q is None:
        body = bytearray(4 * nb)
        if fb == 3:
            body[0] |= 0b010 << 5
        return bytes(body)
    (x0, x1), (y0, y1) = Q
    return x1.to_bytes(nb, "big") + x0.to_bytes(nb, "big") + y1.to_bytes(nb, "big") + y0.to_bytes(nb, "big")


def encode_g1_slice(curve, points, compressed: bool) -> bytes:
    """what the gnark-crypto encoder writes for a []G1Affine: uint32 big-endian length, then the points"""
    return len(points).to_bytes(4, "big") + b"".join(encode_g1(curve, P, compressed) for P in points)


# ---- square roots (Fp: exponentiation or Tonelli-Shanks; Fp2 = Fp[u]/(u^2 - nr) by the norm) and G2 ------------------
def sqrt_fp(p, a):
    """a square root of a mod p, or None"""
    a %= p
    if a == 0:
        return 0
    if pow(a, (p - 1) // 2, p) != 1:
        return None
    if p % 4 == 3:
        return pow(a, (p + 1) // 4, p)
    t, s = p - 1, 0
    while t % 2 == 0:
        t //= 2
        s += 1
    g = 2
    while pow(g, (p - 1) // 2, p) != p - 1:
        g += 1
    c, x, b, v = pow(g, t, p), pow(a, (t + 1) // 2, p), pow(a, t, p), s
    while b != 1:
        m, t2 = 0, b
        while t2 != 1:
            t2 = t2 * t2 % p
            m += 1
        cc = pow(c, 1 << (v - m - 1), p)
        x, c = x * cc % p, cc * cc % p
        b, v = b * c % p, m
    return x


def sqrt_fp2(curve, a):
    """a square root of a = (a0, a1) in Fp[u]/(u^2 - nr), nr = curve.fp2_nonresidue (-1 or -5), or None"""
    p, nr = curve.p, curve.fp2_nonresidue
    a0, a1 = a[0] % p, a[1] % p
    if a1 == 0:
        s = sqrt_fp(p, a0)
        if s is not None:
            return (s, 0)
        s = sqrt_fp(p, a0 * pow(nr, -1, p))
        return None if s is None else (0, s)
    al = sqrt_fp(p, a0 * a0 - nr * a1 * a1)          # the norm
    if al is None:
        return None
    inv2 = pow(2, -1, p)
    for cand in (a0 + al, a0 - al):
        x0 = sqrt_fp(p, cand * inv2)
        if x0 is None or x0 == 0:
            continue
        x1 = a1 * pow(2 * x0, -1, p) % p
        if ((x0 * x0 + nr * x1 * x1) % p, 2 * x0 * x1 % p) == (a0, a1):
            return (x0, x1)
    return None


def twist_b(curve):
    """b' of the G2 curve y^2 = x^3 + b' (an Fp2 pair, or an int for BW6-761 whose G2 is over Fp)"""
    p = curve.p
    from . import ff
    if curve.name == "bw6-761":
        return 4
    F2 = ff.Fp2(p, curve.fp2_nonresidue)
    if curve.name == "bn254":
        return F2.mul((3, 0), F2.inv((9, 1)))
    if curve.name == "bls12-381":
        return (4, 4)
    return F2.inv((0, 1))                              # BLS12-377: 1 / u


def _largest_fp2(p, y):
    return y[1] > (p - 1) // 2 if y[1] != 0 else y[0] > (p - 1) // 2


def encode_g2(curve, Q, compressed: bool) -> bytes:
    """G2Affine.Bytes() (compressed) / RawBytes(): over Fp2 X.A1 || X.A0 [|| Y.A1 || Y.A0]; BW6-761: the G1 layout"""
    if curve.fp2_nonresidue is None:
        nb, fb = _fp_bytes(curve), _flag_bits(curve)
        if Q is None or not compressed:
            return encode_g1(curve, Q, compressed, check=False)
        xb = bytearray(Q[0].to_bytes(nb, "big"))
        xb[0] |= (0b101 if Q[1] > (curve.p - 1) // 2 else 0b100) << (8 - fb)
        return bytes(xb)
    if not compressed:
        return encode_g2_raw(curve, Q)
    nb, fb = _fp_bytes(curve), _flag_bits(curve)
    sh = 8 - fb
    body = bytearray(2 * nb)
    if Q is None:
        body[0] |= (0b01 if fb == 2 else 0b110) << sh
        return bytes(body)
    (x0, x1), y = Q
    body[:nb] = x1.to_bytes(nb, "big")
    body[nb:] = x0.to_bytes(nb, "big")
    assert body[0] >> sh == 0
    big = _largest_fp2(curve.p, y)
    body[0] |= ((0b11 if big else 0b10) if fb == 2 else (0b101 if big else 0b100)) << sh
    return bytes(body)


def decode_g2(curve, b: bytes):
    """compressed G2 over Fp2 -> ((x0, x1), (y0, y1)) or None; BW6-761 -> (x, y)"""
    nb, fb = _fp_bytes(curve), _flag_bits(curve)
    sh = 8 - fb
    flag = b[0] >> sh
    if fb == 2:
        assert flag != 0
        inf, largest = flag == 0b01, flag == 0b11
    else:
        assert flag in (0b100, 0b101, 0b110)
        inf, largest = flag == 0b110, flag == 0b101
    if inf:
        return None
    p = curve.p
    if curve.fp2_nonresidue is None:
        x = int.from_bytes(b[:nb], "big") & ((1 << (8 * nb - fb)) - 1)
        y = sqrt_fp(p, x * x * x + twist_b(curve))
        assert y is not None, "x is not on the curve"
        return (x, y if (y > (p - 1) // 2) == largest else p - y)
    from . import ff
    F2 = ff.Fp2(p, curve.fp2_nonresidue)
    x1 = int.from_bytes(b[:nb], "big") & ((1 << (8 * nb - fb)) - 1)
    x0 = int.from_bytes(b[nb:2 * nb], "big")
    assert x0 < p and x1 < p
    x = (x0, x1)
    y = sqrt_fp2(curve, F2.add(F2.mul(F2.sqr(x), x), twist_b(curve)))
    assert y is not None, "x is not on the twist"
    if _largest_fp2(p, y) != largest:
        y = F2.neg(y)
    return (x, y)
