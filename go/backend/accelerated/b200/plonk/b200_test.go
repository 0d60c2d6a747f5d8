//go:build b200

package plonk_test

import (
	"bytes"
	"fmt"
	"math/big"
	"testing"

	"github.com/consensys/gnark-crypto/ecc"
	"github.com/consensys/gnark/backend"
	"github.com/consensys/gnark/backend/accelerated/b200"
	b200_plonk "github.com/consensys/gnark/backend/accelerated/b200/plonk"
	native_plonk "github.com/consensys/gnark/backend/plonk"
	"github.com/consensys/gnark/frontend"
	"github.com/consensys/gnark/frontend/cs/scs"
	"github.com/consensys/gnark/test"
	"github.com/consensys/gnark/test/unsafekzg"
)

// x^3 + x + 5 == y chained so that the system has a few dozen rows (the device prover needs a domain of at least 8)
type circuit struct {
	X frontend.Variable
	Y frontend.Variable `gnark:",public"`
}

func (c *circuit) Define(api frontend.API) error {
	x := c.X
	for i := 0; i < 10; i++ {
		x3 := api.Mul(x, x, x)
		x = api.Add(x3, x, 5)
	}
	api.AssertIsEqual(c.Y, x)
	return nil
}

// witnessY runs the circuit's recurrence in the scalar field
func witnessY(x int64, modulus *big.Int) *big.Int {
	v := big.NewInt(x)
	for i := 0; i < 10; i++ {
		x3 := new(big.Int).Exp(v, big.NewInt(3), modulus)
		v = x3.Add(x3, v).Add(x3, big.NewInt(5)).Mod(x3, modulus)
	}
	return v
}

var curves = []ecc.ID{ecc.BLS12_377, ecc.BLS12_381, ecc.BN254, ecc.BW6_761}

// prove on the device, verify with the reference verifier; keys travel through their serialisation as in the
// accelerated Groth16 tests (backend/accelerated/icicle/groth16/marshal_test.go)
func TestProveVerify(t *testing.T) {
	for _, curve := range curves {
		t.Run(fmt.Sprintf("curve=%s", curve.String()), func(t *testing.T) {
			assert := test.NewAssert(t)
			ccs, err := frontend.Compile(curve.ScalarField(), scs.NewBuilder, &circuit{})
			assert.NoError(err)
			srs, srsLagrange, err := unsafekzg.NewSRS(ccs)
			assert.NoError(err)
			nativePK, vk, err := native_plonk.Setup(ccs, srs, srsLagrange)
			assert.NoError(err)
			accPK := b200_plonk.NewProvingKey(curve)
			buf := new(bytes.Buffer)
			_, err = nativePK.WriteTo(buf)
			assert.NoError(err)
			_, err = accPK.ReadFrom(buf)
			assert.NoError(err)

			assignment := circuit{X: 3, Y: witnessY(3, curve.ScalarField())}
			w, err := frontend.NewWitness(&assignment, curve.ScalarField())
			assert.NoError(err)
			pw, err := w.Public()
			assert.NoError(err)
			proofNative, err := native_plonk.Prove(ccs, nativePK, w)
			assert.NoError(err)
			proofAcc, err := b200_plonk.Prove(ccs, accPK, w)
			assert.NoError(err)
			assert.NoError(b200_plonk.Verify(proofNative, vk, pw))
			assert.NoError(b200_plonk.Verify(proofAcc, vk, pw))
		})
	}
}

// backend.WithStatisticalZeroKnowledge: the quotient shards are randomised (backend/plonk/bn254/prove.go:239-242,
// 689-722); the proof must verify with the unchanged verifier, and two proofs of the same witness must differ in H
func TestProveVerifyStatisticalZK(t *testing.T) {
	for _, curve := range curves {
		t.Run(fmt.Sprintf("curve=%s", curve.String()), func(t *testing.T) {
			assert := test.NewAssert(t)
			ccs, err := frontend.Compile(curve.ScalarField(), scs.NewBuilder, &circuit{})
			assert.NoError(err)
			srs, srsLagrange, err := unsafekzg.NewSRS(ccs)
			assert.NoError(err)
			pk, vk, err := b200_plonk.Setup(ccs, srs, srsLagrange)
			assert.NoError(err)
			assignment := circuit{X: 3, Y: witnessY(3, curve.ScalarField())}
			w, err := frontend.NewWitness(&assignment, curve.ScalarField())
			assert.NoError(err)
			pw, err := w.Public()
			assert.NoError(err)
			szk := b200.WithProverOptions(backend.WithStatisticalZeroKnowledge())
			proof1, err := b200_plonk.Prove(ccs, pk, w, szk)
			assert.NoError(err)
			proof2, err := b200_plonk.Prove(ccs, pk, w, szk)
			assert.NoError(err)
			assert.NoError(b200_plonk.Verify(proof1, vk, pw))
			assert.NoError(b200_plonk.Verify(proof2, vk, pw))
			var b1, b2 bytes.Buffer
			_, err = proof1.WriteTo(&b1)
			assert.NoError(err)
			_, err = proof2.WriteTo(&b2)
			assert.NoError(err)
			assert.False(bytes.Equal(b1.Bytes(), b2.Bytes()))
		})
	}
}
