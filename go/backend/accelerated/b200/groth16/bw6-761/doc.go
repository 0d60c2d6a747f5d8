// Package bn254 implements the B200-accelerated Groth16 prover for the BW6-761 curve.
package bw6761
