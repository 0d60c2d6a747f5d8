"""The C-ABI library loads on a CPU-only box and exports every symbol include/gnark_b200.h
declares; product code fails loudly (no CPU fallback) when no device is present."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gnark_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_exports(b200lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(b200lib, s), f"{s} declared in include/gnark_b200.h but not exported"
    from gnark_b200 import lib
    assert set(lib.EXPORTS) == set(syms)
    assert b"sm_100a" in b200lib.b200_version()


def test_no_cpu_fallback(b200lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    from gnark_b200 import lib
    with pytest.raises(lib.B200Error):
        lib.init([0])
    with pytest.raises(lib.B200Error):
        lib.Table(lib.BN254, 1, np.zeros((4, 8), dtype=np.uint64), precomp=False)
    with pytest.raises(lib.B200Error):
        lib.Domain(lib.BN254, 4)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: no product source may import, include or link it."""
    pkg = os.path.join(ROOT, "gnark_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|#\s*include\s+[\"<][^\">]*oracle)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(txt), f
                assert "liboracle" not in txt, f


def test_host_point_helpers(b200lib):
    """b200_point_add_jac / b200_point_to_affine are pure CPU (used to combine shards)."""
    import random
    from gnark_b200 import lib
    from oracle import ec, ff
    from oracle.params import CURVES
    rng = random.Random(2)
    for c in CURVES.values():
        for group in (1, 2):
            F = ff.base_field(c, group)
            from util import pick_base
            _, base = pick_base(c, group, rng)
            p = ec.scalar_mul(F, 1234567, base)
            q = ec.scalar_mul(F, 7654321, base)
            J = lambda pt: np.concatenate([ec.pack_points(c, group, [pt]).reshape(-1),
                                           ff.pack_elements(F.coords(F.one), c.p, c.fp_limbs).reshape(-1)])
            acc = J(p)
            lib.point_add_jac(c.curve_id, group, acc, J(q))
            aff = lib.point_to_affine(c.curve_id, group, acc)
            assert ec.unpack_points(c, group, aff)[0] == ec.affine_add(F, p, q)
            acc = J(p)
            lib.point_add_jac(c.curve_id, group, acc, J(p))           # doubling path
            assert ec.unpack_points(c, group, lib.point_to_affine(c.curve_id, group, acc))[0] == ec.affine_add(F, p, p)
