// Package bn254 implements the B200-accelerated Groth16 prover for the BN254 curve.
package bn254
