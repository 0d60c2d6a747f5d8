"""Host logic of gnark_b200/groth16.py (the mirror of backend/accelerated/icicle/groth16) without a GPU: the C library is
replaced by a recorder, so what is checked is WHICH entry points are called, with which handles, in which order -
key caching per device (setupDevicePointers is once per key and device, icicle.go:88-264), release on a placement
change, the one-process multi-GPU path (WithDevices: one shard per device, concurrent device parts, one assembly)."""
import ctypes
import threading

import numpy as np
import pytest

from gnark_b200 import groth16 as b200
from gnark_b200 import lib as real_lib


class FakeC:
    def __init__(self):
        self.calls, self.next_handle, self.live = [], 100, {}
        self.lock = threading.Lock()

    def b200_groth16_pk_load(self, dev, desc_ref, out_ref):
        d = desc_ref._obj
        with self.lock:
            self.next_handle += 1
            h = self.next_handle
            self.live[h] = (dev, d.shard_rank, d.shard_world, d.flags)
            self.calls.append(("load", dev, d.shard_rank, d.shard_world))
        out_ref._obj.value = h
        return 0

    def b200_groth16_pk_free(self, h):
        with self.lock:
            self.calls.append(("free", self.live.pop(h.value)[0]))
        return 0

    def b200_groth16_prove(self, h, w, a, b, c, n, r, s, ar, bs, krs, msm):
        with self.lock:
            self.calls.append(("prove", self.live[h.value][0]))
        return 0

    def b200_groth16_msms(self, h, w, a, b, c, n, out):
        dev, rank, world, _ = self.live[h.value]
        arr = (ctypes.c_uint64 * 1).from_address(out.value)
        arr[0] = 1000 + rank                     # a recognisable "partial result"
        with self.lock:
            self.calls.append(("msms", dev, rank, world))
        return 0

    def b200_groth16_assemble(self, h, msm, r, s, ar, bs, krs):
        with self.lock:
            self.calls.append(("assemble", self.live[h.value][0], (ctypes.c_uint64 * 1).from_address(msm.value)[0]))
        return 0


class FakeLib:
    """stands in for gnark_b200.lib inside groth16.py"""
    CURVE_SHAPES = real_lib.CURVE_SHAPES
    TABLE_PRECOMP = real_lib.TABLE_PRECOMP
    Groth16PkDesc = real_lib.Groth16PkDesc
    BN254 = real_lib.BN254

    def __init__(self):
        self.c = FakeC()

    def load(self):
        return self.c

    @staticmethod
    def ptr(a):
        return real_lib.ptr(a)

    @staticmethod
    def check(rc):
        assert rc == 0

    comm = (1, 0)                                # (world, rank) of the library's communicator; (1, 0) = none

    def comm_info(self, dev):
        return self.comm

    @staticmethod
    def point_add_jac(curve, group, acc, q):
        acc[0] += q[0]                           # "group addition" of the recognisable partials
        return acc


@pytest.fixture
def fake(monkeypatch):
    f = FakeLib()
    monkeypatch.setattr(b200, "_lib", f)
    return f


def make_key_and_solution():
    frl, fpl, deg = real_lib.CURVE_SHAPES[real_lib.BN254]
    z = lambda n: np.zeros(n, dtype=np.uint64)
    nb_wires = 5
    pk = b200.ProvingKey.from_arrays(real_lib.BN254, 4, z(2 * fpl), z(2 * fpl), z(2 * fpl), z(2 * fpl * nb_wires),
                                     z(2 * fpl * nb_wires), z(2 * fpl * 3), z(2 * fpl * 3), z(2 * fpl * deg),
                                     z(2 * fpl * deg), z(2 * fpl * deg * nb_wires), [0] * nb_wires, [0] * nb_wires, 2)
    sol = b200.R1CSSolution(W=z(frl * nb_wires), A=z(frl * 3), B=z(frl * 3), C=z(frl * 3))
    return pk, sol


def test_single_device_key_is_loaded_once_and_released_on_a_placement_change(fake):
    pk, sol = make_key_and_solution()
    rnd = b200.WithRandomness(lambda q: 7)
    b200.ProveSolution(pk, sol, b200.WithDeviceID(0), rnd)
    b200.ProveSolution(pk, sol, b200.WithDeviceID(0), rnd)
    assert fake.c.calls == [("load", 0, 0, 1), ("prove", 0), ("prove", 0)]
    b200.ProveSolution(pk, sol, b200.WithDeviceID(1), rnd)                       # other device: old handle released first
    assert fake.c.calls[3:] == [("free", 0), ("load", 1, 0, 1), ("prove", 1)]
    b200.ProveSolution(pk, sol, b200.WithDeviceID(1), b200.WithPrecompute(False), rnd)   # other table mode: reloaded
    assert fake.c.calls[6:] == [("free", 1), ("load", 1, 0, 1), ("prove", 1)]
    pk.free_gpu_resources()
    pk.free_gpu_resources()                                                      # idempotent
    assert fake.c.calls[9:] == [("free", 1)] and not fake.c.live


def test_with_devices_runs_one_shard_per_device_and_assembles_once(fake):
    pk, sol = make_key_and_solution()
    rnd = b200.WithRandomness(lambda q: 7)
    proof = b200.ProveSolution(pk, sol, b200.WithDevices(2, 0, 3), rnd, keep_msm=True)
    calls = fake.c.calls
    assert sorted(calls[:3]) == [("load", 0, 1, 3), ("load", 2, 0, 3), ("load", 3, 2, 3)]
    assert sorted(calls[3:6]) == [("msms", 0, 1, 3), ("msms", 2, 0, 3), ("msms", 3, 2, 3)]   # concurrent: any order
    # assembled once, on the device that holds shard 0, from the SUM of the three partial results
    assert calls[6:] == [("assemble", 2, 1000 + 1001 + 1002)]
    assert proof.msm[0] == 3003
    fake.c.calls.clear()
    b200.ProveSolution(pk, sol, b200.WithDevices(2, 0, 3), rnd)                  # keys stay resident
    assert not [c for c in fake.c.calls if c[0] in ("load", "free")]
    fake.c.calls.clear()
    b200.ProveSolution(pk, sol, b200.WithDeviceID(0), rnd)                       # back to one device: shards released
    assert sorted(fake.c.calls[:3]) == [("free", 0), ("free", 2), ("free", 3)]
    assert fake.c.calls[3:] == [("load", 0, 0, 1), ("prove", 0)]
    with pytest.raises(ValueError):
        b200.NewConfig(b200.WithDevices(0, 0))
    with pytest.raises(ValueError):
        b200.NewConfig(b200.WithDevices())


def _sharded_worker(rank, world, port, ret):
    """WithSharding over gloo: each process's device part is the recorder's, the all_gather and the host-side fold of
    the five partial results are the product code"""
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = FakeLib()
    b200._lib = f
    pk, sol = make_key_and_solution()
    proof = b200.ProveSolution(pk, sol, b200.WithDeviceID(0), b200.WithSharding(rank, world),
                               b200.WithRandomness(lambda q: 7), keep_msm=True)
    want_sum = sum(1000 + r for r in range(world))
    ret[rank] = (f.c.calls == [("load", 0, rank, world), ("msms", 0, rank, world), ("assemble", 0, want_sum)]
                 and int(proof.msm[0]) == want_sum)
    dist.destroy_process_group()


def test_with_sharding_gathers_and_folds_over_gloo():
    import random
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + random.randrange(2000)
    ret = mp.Manager().dict()
    mp.spawn(_sharded_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_with_sharding_and_a_library_communicator_proves_in_one_call(fake):
    """b200_comm_init done (comm_info reports the shard world): the partial sums are combined on the device inside
    b200_groth16_prove - no process group, no host-side fold"""
    fake.comm = (4, 2)
    pk, sol = make_key_and_solution()
    b200.ProveSolution(pk, sol, b200.WithDeviceID(0), b200.WithSharding(2, 4), b200.WithRandomness(lambda q: 7))
    assert fake.c.calls == [("load", 0, 2, 4), ("prove", 0)]
