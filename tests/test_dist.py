"""N > 1 path on CPU: gloo, world_size 2.  Sharding + all_gather + host-side combine
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
are the product code; the per-rank partial MSM is supplied by the oracle here (no GPU)."""
import os
import random

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from gnark_b200 import parallel


def test_shard_range():
    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import corelib
    from oracle.params import BN254 as C
    from util import known_dlog_instance
    F, base, PTS, SC, expected = known_dlog_instance(C, 1, n, seed=99)
    off, cnt = parallel.shard_range(n, world, rank)
    local = lambda: corelib.msm(C, 1, PTS[off:off + cnt].copy(), SC[off:off + cnt].copy(), nthreads=1)
    full = parallel.sharded_msm(C.curve_id, 1, local)
    from util import jac_to_affine
    ret[rank] = jac_to_affine(C, 1, full) == expected
    dist.destroy_process_group()


def test_sharded_msm_gloo(b200lib):
    world, n = 2, 301
    port = 29500 + random.randrange(2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def _ntt_worker(rank, world, port, cname, logn, ret):
    """sharded NTT (gnark_b200/parallel_ntt.py) over gloo: the exchange, layouts, twiddles and the DFT over the rank
    index are the product code; each single-GPU C-ABI call is replaced by its oracle twin (no GPU here)."""
    import contextlib
    import types
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gnark_b200 import lib as real_lib, parallel_ntt as pn, plonk as b200_plonk
    from oracle import ff, ntt
    from oracle.params import CURVES
    from test_plonk_orchestration import make_mock_lib
    c = CURVES[cname]
    pn._lib = make_mock_lib(real_lib, [])
    b200_plonk._new_stream = lambda torch_, dev: types.SimpleNamespace(cuda_stream=1, synchronize=lambda: None)
    b200_plonk._stream_ctx = lambda torch_, s: contextlib.nullcontext()
    n = 1 << logn
    rng = random.Random(4242)                      # same data on every rank
    x = [rng.randrange(c.r) for _ in range(n)]
    X = ff.pack_elements(x, c.r, c.fr_limbs).reshape(n, c.fr_limbs)
    dom = ntt.Domain(c, n)
    ok = True
    sd = pn.ShardedDomain(c.curve_id, logn, rank, world)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1).copy())
    H = lambda t: t.numpy().view(np.uint64).reshape(-1, c.fr_limbs)
    for on_coset in (False, True):
        # forward: natural-order coefficients, CYCLIC shard in -> natural-order evaluations, SLICED shard out
        want = ntt.bit_reverse(dom.fft(list(x), ntt.DIF, on_coset=on_coset))        # DIF output is bit-reversed
        W = ff.pack_elements(want, c.r, c.fr_limbs).reshape(n, c.fr_limbs)
        got = H(sd.forward(T(pn.cyclic_shard(X, world, rank)), on_coset=on_coset))
        ok &= np.array_equal(got, pn.sliced_shard(W, world, rank))
        # inverse: SLICED evaluations in -> CYCLIC coefficients out
        want_i = ntt.bit_reverse(dom.fft_inverse(list(x), ntt.DIF, on_coset=on_coset))
        WI = ff.pack_elements(want_i, c.r, c.fr_limbs).reshape(n, c.fr_limbs)
        got_i = H(sd.inverse(T(pn.sliced_shard(X, world, rank)), on_coset=on_coset))
        ok &= np.array_equal(got_i, pn.cyclic_shard(WI, world, rank))
        # round trip through the exchange
        back = H(sd.inverse(sd.forward(T(pn.cyclic_shard(X, world, rank)), on_coset=on_coset), on_coset=on_coset))
        ok &= np.array_equal(back, pn.cyclic_shard(X, world, rank))
    # shard helpers are inverse to each other
    ok &= np.array_equal(pn.cyclic_unshard([pn.cyclic_shard(X, world, r) for r in range(world)]), X)
    ok &= np.array_equal(pn.sliced_unshard([pn.sliced_shard(X, world, r) for r in range(world)]), X)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,cname,logn", [(2, "bn254", 5), (4, "bls12-381", 6)])
def test_sharded_ntt_gloo(b200lib, world, cname, logn):
    port = 31500 + random.randrange(2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ntt_worker, args=(world, port, cname, logn, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def _plonk_worker(rank, world, port, logn, ret):
    """sharded PLONK prover (gnark_b200/plonk.py with shard=(rank, world, pg)) over gloo: point-range-sharded KZG
    commitments (all_gather + host adds) and coset-parallel quotient (all_reduce of disjoint quarters) are the
    product code; every single-GPU C-ABI call is replaced by its oracle twin (no GPU here)."""
    import contextlib
    import types
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gnark_b200 import lib as real_lib, plonk as b200_plonk
    from oracle import corelib, ec, ff, plonk_prover as pp
    from oracle.params import CURVES
    from test_plonk_orchestration import make_mock_lib
    from util import jac_to_affine
    c = CURVES["bn254"]
    mock = make_mock_lib(real_lib, [])
    mock.point_add_jac = real_lib.point_add_jac            # host-side group addition: the real (CPU) entry point
    b200_plonk._lib = mock
    b200_plonk._device = lambda dev: "cpu"
    b200_plonk._new_stream = lambda torch_, dev: types.SimpleNamespace(cuda_stream=1, synchronize=lambda: None)
    b200_plonk._stream_ctx = lambda torch_, s: contextlib.nullcontext()
    rng = random.Random(777)                               # same instance on every rank
    r, L = c.r, c.fr_limbs
    n = 1 << logn
    circ, l, rr, o = pp.random_satisfied_instance(c, n, seed=logn + 3)
    rnd = lambda: rng.randrange(r)
    ch = pp.Challenges(gamma=rnd(), beta=rnd(), alpha=rnd(), zeta=rnd(), v=rnd(), bl=[rnd(), rnd()], br=[rnd(), rnd()],
                       bo=[rnd(), rnd()], bz=[rnd(), rnd(), rnd()])
    tau = rnd()
    want = pp.prove(c, circ, l, rr, o, ch, tau)
    pe = lambda v: ff.pack_elements(v, r, L)
    srs = corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), pe([pow(tau, i, r) for i in range(n + 3)]))
    pk = b200_plonk.ProvingKey.from_trace(c.curve_id, logn, pe(circ.ql), pe(circ.qr), pe(circ.qm), pe(circ.qo),
                                         pe(circ.qk), np.array(circ.perm, dtype=np.int64), srs, shard=(rank, world, None))
    got = b200_plonk.Prove(pk, pe(l), pe(rr), pe(o),
                           b200_plonk.Challenges(gamma=ch.gamma, beta=ch.beta, alpha=ch.alpha, zeta=ch.zeta, v=ch.v,
                                                 bl=ch.bl, br=ch.br, bo=ch.bo, bz=ch.bz))
    F = ff.Fp(c.p)
    ok = True
    for g_, w_ in ((got.LRO[0], want.L), (got.LRO[1], want.R), (got.LRO[2], want.O), (got.Z, want.Z),
                   (got.H[0], want.H[0]), (got.H[1], want.H[1]), (got.H[2], want.H[2]), (got.LinearizedDigest, want.lin),
                   (got.BatchedProofH, want.batch_opening), (got.ZShiftedOpeningH, want.z_opening)):
        ok &= jac_to_affine(c, 1, g_) == ec.scalar_mul(F, w_, c.g1)
    ok &= got.BatchedClaimedValues == want.claimed and got.ZShiftedClaimedValue == want.zu
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,logn", [(2, 4), (3, 3)])
def test_sharded_plonk_gloo(b200lib, world, logn):
    port = 33500 + random.randrange(2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_plonk_worker, args=(world, port, logn, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
