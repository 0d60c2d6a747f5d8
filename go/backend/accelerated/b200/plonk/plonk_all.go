// Package plonk implements PLONK proof system with B200 acceleration.
//
// The reference has no accelerated PLONK package; this one mirrors the function set of
// [github.com/consensys/gnark/backend/plonk] (backend/plonk/plonk.go:94-135) with the option type of the accelerated
// Groth16 packages.
package plonk

import (
	"github.com/consensys/gnark-crypto/ecc"
	"github.com/consensys/gnark-crypto/kzg"
	"github.com/consensys/gnark/backend"
	"github.com/consensys/gnark/backend/plonk"
	"github.com/consensys/gnark/backend/witness"
	"github.com/consensys/gnark/constraint"
)

// Verify verifies a PLONK proof. It wraps [plonk.Verify], provided for completeness.
func Verify(proof plonk.Proof, vk plonk.VerifyingKey, publicWitness witness.Witness, opts ...backend.VerifierOption) error {
	return plonk.Verify(proof, vk, publicWitness, opts...)
}

// NewVerifyingKey creates a new empty verifying key for deserializing into. It is compatible with [plonk.NewVerifyingKey].
func NewVerifyingKey(curveID ecc.ID) plonk.VerifyingKey {
	return plonk.NewVerifyingKey(curveID)
}

// NewProof creates a new empty proof for deserializing into. It is compatible with [plonk.NewProof].
func NewProof(curveID ecc.ID) plonk.Proof {
	return plonk.NewProof(curveID)
}

// NewCS creates a new typed SparseR1CS for the given curve. It is compatible with [plonk.NewCS].
func NewCS(curveID ecc.ID) constraint.ConstraintSystem {
	return plonk.NewCS(curveID)
}

// SRSSize returns the KZG SRS sizes a circuit needs. It wraps [plonk.SRSSize].
func SRSSize(ccs constraint.ConstraintSystem) (sizeCanonical, sizeLagrange int) {
	return plonk.SRSSize(ccs)
}

var _ kzg.SRS // the Setup signatures of the tagged files take kzg.SRS values
