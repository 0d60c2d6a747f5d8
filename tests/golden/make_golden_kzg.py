#!/usr/bin/env python3
"""Extracts the reference's ONE externally produced fixture on this path into tests/golden/eth_kzg_srs_v1.bin.

Source: /root/reference/std/evmprecompiles/kzg_trusted_setup.json - the output of the Ethereum KZG ceremony
(BLS12-381), which the reference's own tests load into a gnark-crypto `kzg.ProvingKey` and commit / open with
(std/evmprecompiles/10-kzg_point_evaluation_test.go:50-71, 853-871).  It holds
    g1_monomial[k] = tau^k * G1      (k < 4096)
    g1_lagrange[i] = L_i(tau) * G1   (i < 4096, natural order, L_i over the 4096-th roots of unity w^i)
    g2_monomial[k] = tau^k * G2      (k < 65)
for a tau nobody knows.  Because both G1 bases are given, it is a set of known-answer vectors produced OUTSIDE this
repository for exactly the operations on the hot path:
    g1_monomial[k] = sum_i w^(i*k) * g1_lagrange[i]              (4096-point BLS12-381 G1 MSM, known answer)
    g1_lagrange[i] = (1/n) * sum_k w^(-i*k) * g1_monomial[k]     (idem)
    MSM(g1_monomial, c) = MSM(g1_lagrange, NTT(c))               (Fr NTT of size 2^12 in gnark-crypto's ordering and
                                                                  with gnark-crypto's root of unity, through the MSM)
The file written is the data only (no reference source): the compressed points exactly as the JSON spells them,
concatenated:  4096 x 48 B (monomial) | 4096 x 48 B (lagrange) | 65 x 96 B (G2 monomial)  = 399 456 bytes.
    python tests/golden/make_golden_kzg.py
"""
import hashlib
import json
import os

SRC = "/root/reference/std/evmprecompiles/kzg_trusted_setup.json"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "eth_kzg_srs_v1.bin")


def main():
    d = json.load(open(SRC))
    assert [len(d[k]) for k in ("g1_monomial", "g1_lagrange", "g2_monomial")] == [4096, 4096, 65]
    blob = b"".join(bytes.fromhex(h[2:]) for k in ("g1_monomial", "g1_lagrange", "g2_monomial") for h in d[k])
    assert len(blob) == 4096 * 48 * 2 + 65 * 96
    open(OUT, "wb").write(blob)
    print(OUT, len(blob), "bytes, sha256", hashlib.sha256(blob).hexdigest())


if __name__ == "__main__":
    main()
