//go:build !b200

package plonk

import (
	"github.com/consensys/gnark-crypto/ecc"
	"github.com/consensys/gnark-crypto/kzg"
	"github.com/consensys/gnark/backend/accelerated/b200"
	"github.com/consensys/gnark/backend/plonk"
	"github.com/consensys/gnark/backend/witness"
	"github.com/consensys/gnark/constraint"
)

// Prove generates a PLONK proof from a circuit, its B200 proving key and the full witness.
func Prove(ccs constraint.ConstraintSystem, pk plonk.ProvingKey, fullWitness witness.Witness, opts ...b200.Option) (plonk.Proof, error) {
	panic("b200 backend requested but program compiled without 'b200' build tag")
}

// Setup prepares the public data associated to a circuit. It wraps [plonk.Setup]; the returned proving key is a B200
// proving key.
func Setup(ccs constraint.ConstraintSystem, srs, srsLagrange kzg.SRS) (plonk.ProvingKey, plonk.VerifyingKey, error) {
	panic("b200 backend requested but program compiled without 'b200' build tag")
}

// NewProvingKey creates a new empty proving key for deserializing into. It is compatible with [plonk.NewProvingKey],
// but returns a B200 proving key.
func NewProvingKey(curveID ecc.ID) plonk.ProvingKey {
	panic("b200 backend requested but program compiled without 'b200' build tag")
}
