// Package plonk implements the B200-accelerated PLONK prover for the BN254 curve.
package plonk
