"""ctypes front end of the C++ oracle (oracle/c/oracle.cpp).  TEST INFRASTRUCTURE ONLY
(see oracle/params.py header for what pins parity and what does not).

Builds oracle/_build/liboracle.so on demand with g++ (no reference sources are
compiled: gnark is Go and its arithmetic is in an absent module, so there is no
oracle/_ref for this repository - DESIGN.md says so).
"""

import ctypes
import os
import subprocess

import numpy as np

from . import ff
from .params import CurveParams

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "c", "oracle.cpp")
OUT = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-madx", "-std=c++17", "-fPIC", "-shared", "-pthread", SRC, "-o", OUT + ".tmp"]
    subprocess.check_call(cmd)
    os.replace(OUT + ".tmp", OUT)
    return OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def _mods(curve: CurveParams):
    return ff.int_to_limbs(curve.p, curve.fp_limbs), ff.int_to_limbs(curve.r, curve.fr_limbs)


def _deg_beta(curve, group):
    if group == 1 or curve.fp2_nonresidue is None:
        return 1, -1
    return 2, curve.fp2_nonresidue


def default_threads():
    return os.cpu_count() or 1


def msm(curve, group, points: np.ndarray, scalars: np.ndarray, n=None, c=None, nthreads=None,
        batch_affine=False) -> np.ndarray:
    """Pippenger MSM; returns Jacobian {X,Y,Z} limbs (gnark layout).  batch_affine: affine buckets with batched
    additions (what gnark-crypto's MultiExp does for large windows) - the faster CPU arm; the default
    extended-Jacobian variant is the simpler one the parity tests check against."""
    deg, beta = _deg_beta(curve, group)
    pm, rm = _mods(curve)
    if n is None:
        n = scalars.size // curve.fr_limbs
    if c is None:
        c = 4 if n < 32 else min(16, max(4, int(np.log2(max(n, 2))) - 4))
    out = np.zeros(3 * deg * curve.fp_limbs, dtype=np.uint64)
    fn = lib().orc_msm_batch_affine if batch_affine else lib().orc_msm
    rc = fn(_p(pm), curve.fp_limbs, _p(rm), curve.fr_limbs, deg, beta, _p(points), _p(scalars),
            ctypes.c_size_t(n), c, nthreads or default_threads(), _p(out))
    assert rc == 0
    return out


def msm_naive(curve, group, points, scalars, n=None) -> np.ndarray:
    deg, beta = _deg_beta(curve, group)
    pm, rm = _mods(curve)
    if n is None:
        n = scalars.size // curve.fr_limbs
    out = np.zeros(3 * deg * curve.fp_limbs, dtype=np.uint64)
    rc = lib().orc_msm_naive(_p(pm), curve.fp_limbs, _p(rm), curve.fr_limbs, deg, beta, _p(points), _p(scalars),
                             ctypes.c_size_t(n), _p(out))
    assert rc == 0
    return out


def fixed_base(curve, group, base_affine: np.ndarray, scalars: np.ndarray, n=None, nthreads=None) -> np.ndarray:
    """[k_i * base] as affine points (BatchScalarMultiplicationG1/G2)."""
    deg, beta = _deg_beta(curve, group)
    pm, rm = _mods(curve)
    if n is None:
        n = scalars.size // curve.fr_limbs
    out = np.zeros((n, 2 * deg * curve.fp_limbs), dtype=np.uint64)
    rc = lib().orc_fixed_base(_p(pm), curve.fp_limbs, _p(rm), curve.fr_limbs, deg, beta, _p(base_affine),
                              _p(scalars), ctypes.c_size_t(n), _p(out), nthreads or default_threads())
    assert rc == 0
    return out


def _gen_coset(curve, logn, generator, coset_gen):
    r = curve.r
    if generator is None:
        generator = pow(curve.root_of_unity, 1 << (curve.two_adicity - logn), r)
    if coset_gen is None:
        coset_gen = curve.mult_gen
    g = ff.pack_elements([generator], r, curve.fr_limbs)
    c = ff.pack_elements([coset_gen], r, curve.fr_limbs)
    return g, c


def ntt(curve, data: np.ndarray, logn, inverse, decimation, on_coset, generator=None, coset_gen=None, nthreads=None):
    """in place on `data` ((n, fr_limbs) uint64, Montgomery)."""
    _, rm = _mods(curve)
    g, c = _gen_coset(curve, logn, generator, coset_gen)
    rc = lib().orc_ntt(_p(rm), curve.fr_limbs, _p(data), logn, int(inverse), decimation, int(on_coset), _p(g), _p(c),
                       nthreads or default_threads())
    assert rc == 0
    return data


def compute_h(curve, a, b, c, logn, generator=None, coset_gen=None, nthreads=None):
    """in place: a <- h (bit-reversed); b, c clobbered.  All (n, fr_limbs), zero padded by the caller."""
    _, rm = _mods(curve)
    g, cs = _gen_coset(curve, logn, generator, coset_gen)
    rc = lib().orc_compute_h(_p(rm), curve.fr_limbs, _p(a), _p(b), _p(c), logn, _p(g), _p(cs),
                             nthreads or default_threads())
    assert rc == 0
    return a


def fr_dot(curve, a: np.ndarray, b: np.ndarray, n=None) -> int:
    """sum a[i]*b[i] mod r as a canonical int (inputs Montgomery)."""
    _, rm = _mods(curve)
    if n is None:
        n = a.size // curve.fr_limbs
    out = np.zeros(curve.fr_limbs, dtype=np.uint64)
    rc = lib().orc_fr_dot(_p(rm), curve.fr_limbs, _p(a), _p(b), ctypes.c_size_t(n), _p(out))
    assert rc == 0
    # result is Montgomery form of (sum a_m*b_m*R^-1) = (sum a*b)*R  -> canonical
    return ff.unpack_elements(out, curve.r, curve.fr_limbs)[0]


# ---- Fr vector helpers (Montgomery limb arrays) for full-size fixtures / checks (oracle/plonk_fast.py) -------------
def fr_vec(curve, op, a: np.ndarray, b: np.ndarray, nthreads=None) -> np.ndarray:
    """elementwise a*b (op 0), a+b (1), a-b (2) on (n, fr_limbs) Montgomery arrays"""
    _, rm = _mods(curve)
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    assert a.shape == b.shape
    out = np.zeros_like(a)
    rc = lib().orc_fr_vec(_p(rm), curve.fr_limbs, op, _p(a), _p(b), _p(out), ctypes.c_size_t(a.size // curve.fr_limbs),
                          nthreads or default_threads())
    assert rc == 0
    return out


def fr_geometric(curve, start: int, ratio: int, n: int) -> np.ndarray:
    """[start * ratio^i for i < n] (canonical ints in, Montgomery array out)"""
    _, rm = _mods(curve)
    s = ff.pack_elements([start % curve.r], curve.r, curve.fr_limbs)
    q = ff.pack_elements([ratio % curve.r], curve.r, curve.fr_limbs)
    out = np.zeros((n, curve.fr_limbs), dtype=np.uint64)
    rc = lib().orc_fr_geometric(_p(rm), curve.fr_limbs, _p(s), _p(q), ctypes.c_size_t(n), _p(out))
    assert rc == 0
    return out


def fr_lagrange_at(curve, logn: int, x: int, generator=None) -> np.ndarray:
    """[L_i(x) for i < 2^logn] over the domain <generator> (default: gnark-crypto's), Montgomery array"""
    _, rm = _mods(curve)
    g, _c = _gen_coset(curve, logn, generator, None)
    xm = ff.pack_elements([x % curve.r], curve.r, curve.fr_limbs)
    out = np.zeros((1 << logn, curve.fr_limbs), dtype=np.uint64)
    rc = lib().orc_fr_lagrange_at(_p(rm), curve.fr_limbs, logn, _p(g), _p(xm), _p(out))
    assert rc == 0
    return out
