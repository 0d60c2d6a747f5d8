// Package bn254 implements the B200-accelerated Groth16 prover for the BLS12-377 curve.
package bls12377
