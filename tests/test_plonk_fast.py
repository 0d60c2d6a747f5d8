"""oracle/plonk_fast.py (full-size PLONK fixtures + verifier from a verifying key) against the big-int oracle prover at a
small size: its satisfied instance is accepted by the restated prover / verifier, its verifying key equals the one the
list-based verifier derives, proofs verify and tampered proofs do not.  CPU only."""
import numpy as np
import pytest

from oracle import ec, ff, plonk_fast, plonk_prover as pp
from oracle.params import CURVES


def _as_circuit(c, inst):
    un = lambda a: ff.unpack_elements(np.ascontiguousarray(a), c.r, c.fr_limbs)
    circ = pp.Circuit(n=inst.n, ql=un(inst.ql), qr=un(inst.qr), qm=un(inst.qm), qo=un(inst.qo), qk=un(inst.qk),
                      perm=[int(x) for x in inst.perm])
    return circ, un(inst.l), un(inst.r), un(inst.o)


@pytest.mark.parametrize("cname", ["bls12-381", "bn254"])
def test_fast_fixture_and_verifier(cname):
    c = CURVES[cname]
    inst = plonk_fast.satisfied_instance(c, 5, seed=3)
    assert plonk_fast.check_satisfied(inst)
    assert (inst.perm != np.arange(3 * inst.n)).sum() >= inst.n // 2          # non-trivial copy constraints
    circ, l, r, o = _as_circuit(c, inst)
    proof = pp.prove(c, circ, l, r, o, inst.ch, inst.tau)
    assert pp.verify(c, circ, proof, inst.ch, inst.tau)
    F1 = ff.Fp(c.p)
    dl = [proof.L, proof.R, proof.O, proof.Z, proof.H[0], proof.H[1], proof.H[2], proof.lin, proof.batch_opening, proof.z_opening]
    one = ff.pack_elements([1], c.p, c.fp_limbs).reshape(-1)
    pts = np.stack([np.concatenate([ec.pack_points(c, 1, [ec.scalar_mul(F1, d, c.g1)]).reshape(-1), one]) for d in dl])
    vals = ff.pack_elements(proof.claimed[:6] + [proof.zu], c.r, c.fr_limbs)
    assert plonk_fast.verify(c, inst, pts, vals)
    # the verifying key derived with the C++ helpers is the one the list-based verifier derives
    affine = [ec.scalar_mul(F1, d, c.g1) for d in dl]
    assert pp.verify_pairing(c, circ, affine, proof.claimed[:6] + [proof.zu], inst.ch, inst.tau)
    bad = pts.copy()
    bad[3] = bad[2]                                                           # [Z] replaced
    assert not plonk_fast.verify(c, inst, bad, vals, with_pairing=False)
    badv = vals.copy()
    badv[2] = vals[1]                                                         # r(zeta) replaced
    assert not plonk_fast.verify(c, inst, pts, badv, with_pairing=False)
    # an unsatisfied trace is noticed by the fixture's own check
    inst.o[5] = inst.o[6]
    assert not plonk_fast.check_satisfied(inst)


def test_trapdoor_srs_matches_powers():
    c = CURVES["bn254"]
    inst = plonk_fast.satisfied_instance(c, 3, seed=9)
    srs = plonk_fast.trapdoor_srs_cpu(inst)
    F1 = ff.Fp(c.p)
    got = ec.unpack_points(c, 1, srs)
    assert got == [ec.scalar_mul(F1, pow(inst.tau, i, c.r), c.g1) for i in range(inst.n + 3)]
