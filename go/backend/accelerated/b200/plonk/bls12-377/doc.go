// Package plonk implements the B200-accelerated PLONK prover for the BLS12-377 curve.
package plonk
