// Package plonk implements the B200-accelerated PLONK prover for the BW6-761 curve.
package plonk
