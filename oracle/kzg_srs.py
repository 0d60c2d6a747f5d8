"""Decoder for the Ethereum KZG ceremony SRS fixture (tests/golden/eth_kzg_srs_v1.bin).  TEST INFRASTRUCTURE ONLY.

The fixture is the reference's own std/evmprecompiles/kzg_trusted_setup.json (loaded by its tests through
gnark-crypto's `G1Affine.SetBytes` / `G2Affine.SetBytes`, std/evmprecompiles/10-kzg_point_evaluation_test.go:853-903),
extracted by tests/golden/make_golden_kzg.py.  Point encoding = the compressed format gnark-crypto uses for BLS12-381
(the ZCash convention): big-endian x, top three bits of the first byte are flags
    bit 7: compressed, bit 6: infinity, bit 5: y is the lexicographically larger root;
G2: x = (x.A1 | x.A0), the "larger" comparison is on (A1, A0).
Everything it pins is listed in tests/test_golden_kzg.py.
"""
import hashlib
import os

from .params import BLS12_381 as C

N = 4096
LOGN = 12
N_G2 = 65
SHA256 = "753bd011b238fb9b63a35b9b526f7830057ced42b9a9c87368a67105bd8f0566"
PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "eth_kzg_srs_v1.bin")

_P = C.p
_HALF = (_P - 1) // 2


def _sqrt_fp(a):
    y = pow(a, (_P + 1) // 4, _P)        # p = 3 mod 4
    return y if y * y % _P == a % _P else None


def decode_g1(b: bytes):
    assert len(b) == 48 and b[0] & 0x80, "compressed G1 expected"
    if b[0] & 0x40:
        return None
    x = int.from_bytes(b, "big") & ((1 << 381) - 1)
    assert x < _P
    y = _sqrt_fp((x * x * x + C.b) % _P)
    assert y is not None, "x is not on the curve"
    if bool(b[0] & 0x20) != (y > _HALF):
        y = _P - y
    return (x, y)


def _sqrt_fp2(a):
    """square root in Fp[u]/(u^2+1), p = 3 mod 4 (complex method)"""
    a0, a1 = a
    if a1 == 0:
        s = _sqrt_fp(a0)
        if s is not None:
            return (s, 0)
        s = _sqrt_fp((-a0) % _P)
        return (0, s)
    norm = _sqrt_fp((a0 * a0 + a1 * a1) % _P)
    assert norm is not None
    inv2 = pow(2, _P - 2, _P)
    for nn in (norm, _P - norm):
        x0sq = (a0 + nn) * inv2 % _P
        x0 = _sqrt_fp(x0sq)
        if x0 is None or x0 == 0:
            continue
        x1 = a1 * pow(2 * x0, _P - 2, _P) % _P
        return (x0, x1)
    raise AssertionError("not a square in Fp2")


def decode_g2(b: bytes):
    assert len(b) == 96 and b[0] & 0x80, "compressed G2 expected"
    if b[0] & 0x40:
        return None
    x1 = int.from_bytes(b[:48], "big") & ((1 << 381) - 1)
    x0 = int.from_bytes(b[48:], "big")
    # y^2 = x^3 + 4(1+u)
    xx = ((x0 * x0 - x1 * x1) % _P, 2 * x0 * x1 % _P)
    x3 = ((xx[0] * x0 - xx[1] * x1) % _P, (xx[0] * x1 + xx[1] * x0) % _P)
    y = _sqrt_fp2(((x3[0] + 4) % _P, (x3[1] + 4) % _P))
    larger = y[1] > _HALF if y[1] != 0 else y[0] > _HALF
    if bool(b[0] & 0x20) != larger:
        y = ((-y[0]) % _P, (-y[1]) % _P)
    return ((x0, x1), y)


_cache = {}


def load():
    """-> (g1_monomial, g1_lagrange, g2_monomial) as lists of affine points (canonical integers)"""
    if "v" not in _cache:
        blob = open(PATH, "rb").read()
        assert hashlib.sha256(blob).hexdigest() == SHA256, "fixture corrupted"
        g1 = [decode_g1(blob[48 * i:48 * (i + 1)]) for i in range(2 * N)]
        off = 2 * N * 48
        g2 = [decode_g2(blob[off + 96 * i:off + 96 * (i + 1)]) for i in range(N_G2)]
        _cache["v"] = (g1[:N], g1[N:], g2)
    return _cache["v"]
