#!/usr/bin/env python3
"""Extracts the field / curve CONSTANTS that gnark itself wrote into its PLONK verifying-key fixtures
(/root/reference/backend/solidity/testdata/blank_plonk_{bn254,bls12381}_{nocommit,commit}.vk, written by
`plonk.Setup` in backend/solidity/solidity_test.go:244-268 and serialised by backend/plonk/bn254/marshal.go:177-212:
[marker, version,] Size, SizeInv, Generator, NbPublicVariables, CosetShift, S[3], Ql, Qr, Qm, Qo, Qk, Qcp, Kzg.G1,
Kzg.G2[0], Kzg.G2[1], ...) into tests/golden/gnark_vk_constants_v1.json.

These are values computed by gnark-crypto (absent here): the generator of the size-8 / size-16 fft.Domain, the coset
shift (FrMultiplicativeGen) and the G1 / G2 group generators in gnark-crypto's compressed encoding.  The SRS behind
the commitments used a random tau (test/unsafekzg/kzgsrs.go:147-151), so the commitments themselves pin nothing.
    python tests/golden/make_golden_vk_constants.py
"""
import json
import os

SRC = "/root/reference/backend/solidity/testdata"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gnark_vk_constants_v1.json")
SHAPES = {"bn254": (32, 32), "bls12381": (32, 48)}   # sizeof(fr), sizeof(fp)


def main():
    out = {"version": 1, "source": "backend/solidity/testdata/blank_plonk_*.vk", "keys": []}
    for cn, (frb, fpb) in SHAPES.items():
        for kind in ("nocommit", "commit"):
            b = open(f"{SRC}/blank_plonk_{cn}_{kind}.vk", "rb").read()
            off = 0
            if int.from_bytes(b[:8], "big") == 0:      # versioned encoding: marker + version in front (marshal.go:181-183)
                off = 16
            rd = lambda n: int.from_bytes(b[off:off + n], "big")
            size = rd(8); off += 8
            size_inv = rd(frb); off += frb
            gen = rd(frb); off += frb
            npub = rd(8); off += 8
            coset = rd(frb); off += frb
            off += 8 * fpb                               # S[3], Ql, Qr, Qm, Qo, Qk: compressed G1
            nqcp = rd(4); off += 4 + nqcp * fpb
            g1 = b[off:off + fpb]; off += fpb
            g2_0 = b[off:off + 2 * fpb]; off += 2 * fpb
            out["keys"].append({"curve": cn, "circuit": kind, "size": size, "size_inv": hex(size_inv), "generator": hex(gen),
                                "nb_public": npub, "coset_shift": hex(coset), "n_qcp": nqcp,
                                "kzg_g1_compressed": g1.hex(), "kzg_g2_0_compressed": g2_0.hex()})
    json.dump(out, open(OUT, "w"), indent=1)
    print(OUT)


if __name__ == "__main__":
    main()
