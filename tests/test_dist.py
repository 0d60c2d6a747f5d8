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
