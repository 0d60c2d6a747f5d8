"""Multi-GPU MSM: point-range sharding, one process per GPU.

The (scalar, base) index range is cut into `world` contiguous shards - the same
shape as the reference's host-side chunk loop that sums partial Jacobian results
(backend/accelerated/icicle/groth16/bn254/icicle.go:383-411).  Each rank owns the
table shard for its range and produces one partial point; the only exchange is one
all_gather of `world` Jacobian points (<= 8 x 576 B) - group addition is not an
NCCL reduction operator - followed by world-1 host-side group additions
(b200_point_add_jac).  Backend-agnostic: "nccl" on the GPUs, "gloo" in the CPU tests.
"""

from typing import Callable, List, Tuple

import numpy as np

from . import lib as _lib


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """contiguous, balanced: sizes differ by at most one; returns (offset, count)."""
    base, rem = divmod(n, world)
    cnt = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, cnt


def combine_partials(curve: int, group: int, partials: List[np.ndarray]) -> np.ndarray:
    acc = np.ascontiguousarray(partials[0], dtype=np.uint64).copy()
    for p in partials[1:]:
        _lib.point_add_jac(curve, group, acc, np.ascontiguousarray(p, dtype=np.uint64))
    return acc


def sharded_msm(curve: int, group: int, local_msm: Callable[[], np.ndarray], pg=None, device=None) -> np.ndarray:
    """local_msm() -> this rank's partial Jacobian point (uint64 limbs).  Returns the full
    sum on every rank."""
    import torch
    import torch.distributed as dist
    part = np.ascontiguousarray(local_msm(), dtype=np.uint64)
    if not dist.is_initialized() or dist.get_world_size(pg) == 1:
        return part
    t = torch.from_numpy(part.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size(pg))]
    dist.all_gather(parts, t, group=pg)
    return combine_partials(curve, group, [p.cpu().numpy().view(np.uint64) for p in parts])
