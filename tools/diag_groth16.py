#!/usr/bin/env python3
"""Diagnosis: the same verified Groth16 instance proved under the configurations bench.py uses (pinned host vectors,
torch-owned stream, after a stream of pipelined MSMs) - which of the five MSM results / proof points is wrong, if any."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from gnark_b200 import groth16 as g16, lib  # noqa: E402
from oracle import ec, ff, groth16_fast as gf  # noqa: E402
from oracle.params import CURVES  # noqa: E402


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    c = CURVES["bn254"]
    lib.load(); lib.init([0])
    inst = gf.satisfied_instance(c, logn, seed=20)
    fb = lambda group, dl: lib.fixed_base_batch(c.curve_id, group, ec.pack_points(c, group, [c.g1 if group == 1 else c.g2]),
                                                np.ascontiguousarray(dl))
    kp = gf.key_points(inst, fb)
    pk = g16.ProvingKey.from_arrays(c.curve_id, inst.n, kp["alpha"], kp["beta"], kp["delta"], kp["A"], kp["B"], kp["Z"], kp["K"],
                                    kp["beta2"], kp["delta2"], kp["B2"], inst.inf_a, inst.inf_b, inst.nb_public)
    a, b, cc = inst.solution_abc()
    pageable = g16.R1CSSolution(W=inst.wires(), A=a, B=b, C=cc)
    keep = [torch.from_numpy(v.view(np.int64)).pin_memory() for v in (pageable.W, pageable.A, pageable.B, pageable.C)]
    pinned = g16.R1CSSolution(*[k.numpy().view(np.uint64) for k in keep])
    rs = [0x5EED << 200 | 1, 0xFACE << 190 | 2]
    e = gf.expected(inst, *rs)
    F1, F2 = ff.Fp(c.p), ff.base_field(c, 2)
    want = [ec.scalar_mul(F1, d, c.g1) for d in (e.msm_a, e.msm_b, e.msm_z, e.msm_k)] + [ec.scalar_mul(F2, e.msm_b, c.g2)]
    L = 3 * c.fp_limbs

    def check(tag, sol, opts=()):
        it = iter(rs)
        proof = g16.ProveSolution(pk, sol, g16.WithDeviceID(0), *opts, g16.WithRandomness(lambda q: next(it)), keep_msm=True)
        ok = []
        for k in range(4):
            ok.append(ec.from_jac(F1, ec.unpack_points(c, 1, proof.msm[k * L:(k + 1) * L], ncoords=3)[0]) == want[k])
        ok.append(ec.from_jac(F2, ec.unpack_points(c, 2, proof.msm[4 * L:], ncoords=3)[0]) == want[4])
        v = gf.verify_points(inst, ec.unpack_points(c, 1, proof.Ar)[0], ec.unpack_points(c, 2, proof.Bs)[0],
                             ec.unpack_points(c, 1, proof.Krs)[0], e, with_pairing=False)
        print("%-46s msm A,B1,Z,K,B2 = %s  proof = %s" % (tag, ok, v), flush=True)
        return all(ok) and v

    good = True
    good &= check("pageable, library stream", pageable)
    good &= check("pinned, library stream", pinned)
    good &= check("pinned, library stream, again", pinned)
    good &= check("pinned, WithSharding(0, 1)", pinned, (g16.WithSharding(0, 1),))
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    lib.set_stream(0, st.cuda_stream)
    good &= check("pageable, torch stream", pageable)
    good &= check("pinned, torch stream", pinned)
    # a stream of pipelined MSMs + a profile run first, as bench.py's legs do
    n = 1 << 16
    rng = np.random.Generator(np.random.PCG64(1))
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    t = lib.Table(lib.BN254, 1, np.ascontiguousarray(kp["A"][:n]), precomp=True)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros((4, 12), dtype=torch.int64, device="cuda")
    for i in range(4):
        t.msm_pipelined(d_sc, d_out[i], n=n)
    t.join()
    t.msm_profile(d_sc, d_out[0], n=n)
    lib.sync(0)
    t.free()
    good &= check("pinned, torch stream, after pipelined MSMs", pinned)
    good &= check("pageable, torch stream, after pipelined MSMs", pageable)
    pk.free_gpu_resources()
    print("ALL OK" if good else "MISMATCH")
    sys.exit(0 if good else 1)


if __name__ == "__main__":
    main()
