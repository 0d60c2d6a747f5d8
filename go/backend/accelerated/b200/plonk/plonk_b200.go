//go:build b200

package plonk

import (
	"fmt"

	"github.com/consensys/gnark-crypto/ecc"
	"github.com/consensys/gnark-crypto/kzg"
	"github.com/consensys/gnark/backend/plonk"
	plonk_bls12377 "github.com/consensys/gnark/backend/plonk/bls12-377"
	plonk_bls12381 "github.com/consensys/gnark/backend/plonk/bls12-381"
	plonk_bn254 "github.com/consensys/gnark/backend/plonk/bn254"
	plonk_bw6761 "github.com/consensys/gnark/backend/plonk/bw6-761"
	"github.com/consensys/gnark/backend/witness"
	"github.com/consensys/gnark/constraint"
	cs_bls12377 "github.com/consensys/gnark/constraint/bls12-377"
	cs_bls12381 "github.com/consensys/gnark/constraint/bls12-381"
	cs_bn254 "github.com/consensys/gnark/constraint/bn254"
	cs_bw6761 "github.com/consensys/gnark/constraint/bw6-761"

	kzg_bls12377 "github.com/consensys/gnark-crypto/ecc/bls12-377/kzg"
	kzg_bls12381 "github.com/consensys/gnark-crypto/ecc/bls12-381/kzg"
	kzg_bn254 "github.com/consensys/gnark-crypto/ecc/bn254/kzg"
	kzg_bw6761 "github.com/consensys/gnark-crypto/ecc/bw6-761/kzg"

	b200_bls12377 "github.com/consensys/gnark/backend/accelerated/b200/plonk/bls12-377"
	b200_bls12381 "github.com/consensys/gnark/backend/accelerated/b200/plonk/bls12-381"
	b200_bn254 "github.com/consensys/gnark/backend/accelerated/b200/plonk/bn254"
	b200_bw6761 "github.com/consensys/gnark/backend/accelerated/b200/plonk/bw6-761"

	"github.com/consensys/gnark/backend/accelerated/b200"
	"github.com/consensys/gnark/backend/accelerated/b200/internal/cabi"
)

// Prove generates a PLONK proof from a circuit, its B200 proving key and the full witness (backend/plonk/plonk.go:117).
//
// NB! the provided proving key must be a B200 proving key: use [Setup] or [NewProvingKey] + ReadFrom.
func Prove(ccs constraint.ConstraintSystem, pk plonk.ProvingKey, fullWitness witness.Witness, opts ...b200.Option) (plonk.Proof, error) {
	config, err := b200.NewConfig(opts...)
	if err != nil {
		return nil, fmt.Errorf("initializing config: %w", err)
	}
	if err := cabi.Init(config.Devices()); err != nil {
		panic(fmt.Sprintf("b200 device initialisation: %v", err))
	}
	switch tccs := ccs.(type) {
	case *cs_bn254.SparseR1CS:
		return b200_bn254.Prove(tccs, pk.(*b200_bn254.ProvingKey), fullWitness, config)
	case *cs_bls12381.SparseR1CS:
		return b200_bls12381.Prove(tccs, pk.(*b200_bls12381.ProvingKey), fullWitness, config)
	case *cs_bls12377.SparseR1CS:
		return b200_bls12377.Prove(tccs, pk.(*b200_bls12377.ProvingKey), fullWitness, config)
	case *cs_bw6761.SparseR1CS:
		return b200_bw6761.Prove(tccs, pk.(*b200_bw6761.ProvingKey), fullWitness, config)
	default:
		panic("b200 backend requested but SparseR1CS is not of a supported curve")
	}
}

// Setup prepares the public data associated to a circuit (backend/plonk/plonk.go:94). It wraps the per-curve Setup;
// the returned proving key is a B200 proving key embedding the native one.
func Setup(ccs constraint.ConstraintSystem, srs, srsLagrange kzg.SRS) (plonk.ProvingKey, plonk.VerifyingKey, error) {
	switch tccs := ccs.(type) {
	case *cs_bn254.SparseR1CS:
		pk, vk, err := plonk_bn254.Setup(tccs, *srs.(*kzg_bn254.SRS), *srsLagrange.(*kzg_bn254.SRS))
		if err != nil {
			return nil, nil, err
		}
		return &b200_bn254.ProvingKey{ProvingKey: *pk}, vk, nil
	case *cs_bls12381.SparseR1CS:
		pk, vk, err := plonk_bls12381.Setup(tccs, *srs.(*kzg_bls12381.SRS), *srsLagrange.(*kzg_bls12381.SRS))
		if err != nil {
			return nil, nil, err
		}
		return &b200_bls12381.ProvingKey{ProvingKey: *pk}, vk, nil
	case *cs_bls12377.SparseR1CS:
		pk, vk, err := plonk_bls12377.Setup(tccs, *srs.(*kzg_bls12377.SRS), *srsLagrange.(*kzg_bls12377.SRS))
		if err != nil {
			return nil, nil, err
		}
		return &b200_bls12377.ProvingKey{ProvingKey: *pk}, vk, nil
	case *cs_bw6761.SparseR1CS:
		pk, vk, err := plonk_bw6761.Setup(tccs, *srs.(*kzg_bw6761.SRS), *srsLagrange.(*kzg_bw6761.SRS))
		if err != nil {
			return nil, nil, err
		}
		return &b200_bw6761.ProvingKey{ProvingKey: *pk}, vk, nil
	default:
		panic("b200 backend requested but SparseR1CS is not of a supported curve")
	}
}

// NewProvingKey creates a new empty proving key for deserializing into. It is compatible with [plonk.NewProvingKey],
// but returns a B200 proving key.
func NewProvingKey(curveID ecc.ID) plonk.ProvingKey {
	switch curveID {
	case ecc.BN254:
		return &b200_bn254.ProvingKey{}
	case ecc.BLS12_381:
		return &b200_bls12381.ProvingKey{}
	case ecc.BLS12_377:
		return &b200_bls12377.ProvingKey{}
	case ecc.BW6_761:
		return &b200_bw6761.ProvingKey{}
	default:
		panic("b200 backend requested but curve is not supported")
	}
}
