// Package bn254 implements the B200-accelerated Groth16 prover for the BLS12-381 curve.
package bls12381
