"""gnark-crypto's serialised point encodings, restated.  TEST INFRASTRUCTURE ONLY (see oracle/params.py header).

What `G1Affine.Bytes()` (compressed) / `RawBytes()` (uncompressed) and `G2Affine.RawBytes()` write and `SetBytes` reads
(gnark-crypto v0.21.0, ecc/<curve>/marshal.go; the module is absent here, so this follows its published layout):
big-endian canonical coordinates, metadata in the most significant bits of the first byte -

    BN254 (2 bits):   00 uncompressed | 10 compressed, y smallest | 11 compressed, y largest | 01 compressed infinity
    BLS12-381 / BLS12-377 / BW6-761 (3 bits):
                      000 uncompressed | 010 uncompressed infinity | 100 / 101 compressed, y smallest / largest |
                      110 compressed infinity
    "largest" = y > (p - 1) / 2 on canonical values; Fp2 elements are written A1 || A0.

The reference consumes these encodings wherever it serialises keys: backend/plonk/bn254/marshal.go:96-129 (pk.Kzg,
pk.KzgLagrange), backend/groth16/bn254/marshal.go:136-214.  PINNED by reference-held bytes: the compressed G1 generators
inside gnark's serialised verifying keys (backend/solidity/testdata/blank_plonk_{bn254,bls12381}_*.vk: BN254 0x80..01,
BLS12-381 0x97f1d3a7...) and the 8192 compressed BLS12-381 G1 points of std/evmprecompiles/kzg_trusted_setup.json
(tests/test_golden_kzg.py::test_encoding_matches_reference_bytes).  BLS12-377 / BW6-761: same scheme, unpinned.
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
    assert curve.p % 4 == 3, "square root by exponentiation needs p = 3 mod 4"
    y2 = (x * x * x + _b(curve)) % curve.p
    y = pow(y2, (curve.p + 1) // 4, curve.p)
    assert y * y % curve.p == y2, "x is not on the curve"
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
