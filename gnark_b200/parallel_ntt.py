"""Multi-GPU NTT: one transform of size n = G * m sharded over G = 2^g ranks with ONE all-to-all
("six-step / four-step decomposition ... one all-to-all transpose"; the reference has no
multi-device FFT - its fft.Domain runs on the host, backend/groth16/bn254/prove.go:362-386).

Layouts (r = this rank, m = n / G, t < m / G):

  CYCLIC  (coefficient side)   local[j]             = x[j * G + r]                 j < m
  SLICED  (evaluation side)    local[k1 * m/G + t]  = X[k1 * m + r * m/G + t]      k1 < G

forward  (CYCLIC -> SLICED):   X[k] = sum_i x[i] w^(i k)          (on_coset: x[i] first scaled by g^i)
inverse  (SLICED -> CYCLIC):   x[i] = 1/n sum_k X[k] w^(-i k)     (on_coset: then scaled by g^-i)

    k = k1 * m + k2,  i = j * G + r:   w^(ik) = w_m^(j k2) * w^(r k2) * w_G^(r k1)
    forward = local NTT_m over j  ->  twiddle w^(r k2)  ->  all-to-all by k2 slice  ->  DFT_G over r
    inverse = the same steps backwards with inverse roots.

A prover keeps coefficients CYCLIC and evaluations SLICED throughout (PLONK: Lagrange values -> iNTT ->
coefficients -> coset NTT -> pointwise -> iNTT), so every transform costs one exchange of (G-1)/G of the local
vector over NVLink; pointwise stages are layout-agnostic; a KZG commitment of CYCLIC coefficients uses the
strided SRS shard {tau^(jG+r)}.  Everything on the device is a call of the validated single-GPU C ABI
(b200_ntt_async, b200_vec_bit_reverse, b200_vec_scale_powers, b200_vec_op); the exchange is
torch.distributed.all_to_all_single (NCCL on the GPUs, gloo in the CPU tests, tests/test_dist.py).
"""

from typing import List

import numpy as np

from . import lib as _lib
from . import plonk as _plonk
from .plonk import _FR_DOMAIN, _FR_MODULUS, _Fr


def _bitrev(i: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def cyclic_shard(x: np.ndarray, world: int, rank: int) -> np.ndarray:
    """host helper: full (n, limbs) vector -> this rank's CYCLIC shard"""
    return np.ascontiguousarray(x[rank::world])


def sliced_shard(x: np.ndarray, world: int, rank: int) -> np.ndarray:
    """host helper: full (n, limbs) vector -> this rank's SLICED shard"""
    n = x.shape[0]
    m = n // world
    s = m // world
    return np.ascontiguousarray(np.concatenate([x[k1 * m + rank * s:k1 * m + (rank + 1) * s] for k1 in range(world)]))


def cyclic_unshard(parts: List[np.ndarray]) -> np.ndarray:
    world = len(parts)
    out = np.empty((parts[0].shape[0] * world,) + parts[0].shape[1:], dtype=parts[0].dtype)
    for r, p in enumerate(parts):
        out[r::world] = p
    return out


def sliced_unshard(parts: List[np.ndarray]) -> np.ndarray:
    world = len(parts)
    m = parts[0].shape[0]
    s = m // world
    out = np.empty((m * world,) + parts[0].shape[1:], dtype=parts[0].dtype)
    for r, p in enumerate(parts):
        for k1 in range(world):
            out[k1 * m + r * s:k1 * m + (r + 1) * s] = p[k1 * s:(k1 + 1) * s]
    return out


class ShardedDomain:
    """fft.Domain of size 2^log2n spread over `world` ranks (power of two, world^2 <= n)."""

    def __init__(self, curve: int, log2n: int, rank: int, world: int, dev: int = 0, pg=None, coset_gen: int = None):
        import torch
        self.torch = torch
        g = world.bit_length() - 1
        if world != 1 << g or 2 * g > log2n:
            raise ValueError("world must be a power of two with world^2 <= n")
        self.curve, self.log2n, self.rank, self.world, self.dev, self.pg = curve, log2n, rank, world, dev, pg
        self.g = g
        self.n, self.m = 1 << log2n, 1 << (log2n - g)
        self.fr = _Fr(curve)
        q = self.fr.q = _FR_MODULUS[curve]
        s, root, mult_gen = _FR_DOMAIN[curve]
        self.w = pow(root, 1 << (s - log2n), q)              # generator of the size-n domain
        self.w_inv = pow(self.w, -1, q)
        self.wG = pow(self.w, self.m, q)                     # primitive G-th root
        self.coset = mult_gen if coset_gen is None else coset_gen
        # local domain of size m; its generator w^G is gnark-crypto's default generator for 2^(log2n-g)
        self.local = _lib.Domain(curve, log2n - g, dev=dev)
        self.L = self.fr.limbs
        # torch tensor ops, the NCCL exchange and the library's kernels are ordered on ONE stream (see plonk.py)
        self.stream = _plonk._new_stream(torch, dev)

    def _on_stream(self):
        dom = self

        class _Ctx:
            def __enter__(self_inner):
                self_inner.cm = _plonk._stream_ctx(dom.torch, dom.stream)
                self_inner.cm.__enter__()
                _lib.set_stream(dom.dev, dom.stream.cuda_stream)

            def __exit__(self_inner, *exc):
                dom.stream.synchronize()
                _lib.set_stream(dom.dev, 0)
                return self_inner.cm.__exit__(*exc)
        return _Ctx()

    def free(self):
        self.local.free()

    # ---- building blocks (each one validated C-ABI call) -------------------------------------------
    def _scale(self, d, count, s, gpow):
        _lib.vec_scale_powers(self.dev, self.curve, d, count, self.fr.enc(s), self.fr.enc(gpow))

    def _local_ntt(self, d, inverse):
        """size-m transform, natural -> natural (DIF leaves bit-reversed order; undo it)"""
        self.local.ntt_async(d, inverse=inverse, decimation=_lib.DIF)
        _lib.vec_bit_reverse(self.dev, self.curve, d, self.log2n - self.g)

    def _dft_rows(self, d, root):
        """in place on G rows of m/G elements: row[k1] = sum_r row[r] * root^(r k1)   (root: primitive G-th)"""
        t = self.torch
        G, cols = self.world, self.m // self.world
        if G == 1:
            return d
        rows = [d[r * cols * self.L:(r + 1) * cols * self.L] for r in range(G)]
        spare = [t.empty_like(rows[0]) for _ in range(2)]
        q = self.fr.q
        # radix-2 DIF over the row index: natural in, bit-reversed out
        span = G
        while span > 1:
            half = span // 2
            wstep = pow(root, G // span, q)
            for base in range(0, G, span):
                for k in range(half):
                    a, b = base + k, base + k + half
                    s_buf, d_buf = spare
                    _lib.vec_op(self.dev, self.curve, 1, s_buf, rows[a], rows[b], cols)      # a + b
                    _lib.vec_op(self.dev, self.curve, 2, d_buf, rows[a], rows[b], cols)      # a - b
                    tw = pow(wstep, k, q)
                    if tw != 1:
                        self._scale(d_buf, cols, tw, 1)
                    spare = [rows[a], rows[b]]
                    rows[a], rows[b] = s_buf, d_buf
            span = half
        out = t.empty_like(d)
        for p in range(G):
            k1 = _bitrev(p, self.g)
            out[k1 * cols * self.L:(k1 + 1) * cols * self.L] = rows[p]
        return out

    def _all_to_all(self, d):
        import torch.distributed as dist
        if self.world == 1:
            return d
        out = self.torch.empty_like(d)
        dist.all_to_all_single(out, d, group=self.pg)
        return out

    # ---- transforms -----------------------------------------------------------------------------
    def forward(self, d_cyclic, on_coset: bool = False):
        """CYCLIC coefficients (m elements, flat int64 tensor of m*limbs) -> SLICED evaluations (new tensor)"""
        with self._on_stream():
            return self._forward(d_cyclic, on_coset)

    def _forward(self, d_cyclic, on_coset):
        q, r, G = self.fr.q, self.rank, self.world
        d = d_cyclic
        if on_coset:      # x[jG + r] *= g^(jG + r)
            self._scale(d, self.m, pow(self.coset, r, q), pow(self.coset, G, q))
        self._local_ntt(d, inverse=False)                       # y_r[k2]
        if r:
            self._scale(d, self.m, 1, pow(self.w, r, q))      # * w^(r k2)
        z = self._all_to_all(d)                                 # rows r, columns k2 in my slice
        return self._dft_rows(z, self.wG)

    def inverse(self, d_sliced, on_coset: bool = False):
        """SLICED evaluations -> CYCLIC coefficients (new tensor); includes the 1/n factor"""
        with self._on_stream():
            return self._inverse(d_sliced, on_coset)

    def _inverse(self, d_sliced, on_coset):
        q, r, G = self.fr.q, self.rank, self.world
        z = self._dft_rows(d_sliced, pow(self.wG, -1, q))        # rows r (natural), my k2 slice
        d = self._all_to_all(z)                                 # y_r[k2], k2 < m
        # * w^(-r k2), and the 1/G of the inverse DFT_G (the local inverse transform brings 1/m)
        self._scale(d, self.m, pow(G, -1, q), pow(self.w_inv, r, q))
        self._local_ntt(d, inverse=True)
        if on_coset:      # x[jG + r] *= g^-(jG + r)
            gi = pow(self.coset, -1, q)
            self._scale(d, self.m, pow(gi, r, q), pow(gi, G, q))
        return d
