#!/usr/bin/env python3
"""Extracts the elliptic-curve CONSTANTS that the reference spells out in its own sources (values computed by
gnark-crypto and pasted into gnark) into tests/golden/gnark_intree_points_v1.json.  They are known-answer vectors for
scalar multiplication on every curve of this path:

  * GLV endomorphism, G1 (std/algebra/emulated/sw_emulated/params.go:70-71 BN254, :88-89 BLS12-381, :157-158 BW6-761;
    std/algebra/native/sw_bls12377/inner.go:58-63 BLS12-377):  [lambda] P = (omega * x_P, y_P) for every P of order r -
    a scalar multiplication by a 64..190-bit scalar whose answer needs ONE field multiplication to state.
  * G2 (std/algebra/emulated/sw_bn254/g2.go:75-96, sw_bls12381/g2.go:89-110, sw_bw6761/g2.go:90-99): the G2 generator and
    [2^65] G2 (BN254, BLS12-381) / [2^96] G2 (BW6-761).
    python tests/golden/make_golden_intree_points.py
"""
import json
import os
import re

REF = "/root/reference/std/algebra"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gnark_intree_points_v1.json")


def func_body(src, name):
    i = src.index("func " + name)
    return src[i:src.index("\n}\n", i)]


def glv_emulated(fn):
    body = func_body(open(f"{REF}/emulated/sw_emulated/params.go").read(), fn)
    lam = re.search(r'lambda, _ := new\(big\.Int\)\.SetString\("(\d+)", 10\)', body).group(1)
    om = re.search(r'omega, _ := new\(big\.Int\)\.SetString\("(\d+)", 10\)', body).group(1)
    return {"lambda": lam, "omega": om}


def g2_points(path, fp2):
    src = open(path).read()
    out = {}
    for name in ("g2Gen", "g2GenNbits"):
        i = src.index(name + " := &g2AffP{")
        blk = src[i:src.index("\n\t}\n", i)]
        nums = re.findall(r'NewElement\("(\d+)"\)', blk)
        assert len(nums) == (4 if fp2 else 2), (path, name, len(nums))
        out[name] = nums          # X.A0, X.A1, Y.A0, Y.A1  |  X, Y
    k = re.search(r"this is \[2\^(\d+)\]G2", src).group(1)
    out["log2_multiple"] = int(k)
    return out


def main():
    out = {"version": 1, "curves": {}}
    out["curves"]["bn254"] = {"glv": glv_emulated("GetBN254Params"), "g2": g2_points(f"{REF}/emulated/sw_bn254/g2.go", True)}
    out["curves"]["bls12-381"] = {"glv": glv_emulated("GetBLS12381Params"), "g2": g2_points(f"{REF}/emulated/sw_bls12381/g2.go", True)}
    out["curves"]["bw6-761"] = {"glv": glv_emulated("GetBW6761Params"), "g2": g2_points(f"{REF}/emulated/sw_bw6761/g2.go", False)}
    src = open(f"{REF}/native/sw_bls12377/inner.go").read()
    lam = re.search(r"bls12377lambda := new\(big\.Int\)\.SetBytes\(\[\]byte\{([^}]*)\}\)", src).group(1)
    om = re.search(r"bls12377thirdRootOne1 := new\(big\.Int\)\.SetBytes\(\[\]byte\{([^}]*)\}\)", src, re.S).group(1)
    tobig = lambda s: str(int.from_bytes(bytes(int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]{2}", s)), "big"))
    out["curves"]["bls12-377"] = {"glv": {"lambda": tobig(lam), "omega": tobig(om)}}
    json.dump(out, open(OUT, "w"), indent=1)
    print(OUT)


if __name__ == "__main__":
    main()
