//go:build b200

package groth16_test

import (
	"bytes"
	"fmt"
	"testing"

	"github.com/consensys/gnark-crypto/ecc"
	"github.com/consensys/gnark/backend/accelerated/b200"
	b200_groth16 "github.com/consensys/gnark/backend/accelerated/b200/groth16"
	native_groth16 "github.com/consensys/gnark/backend/groth16"
	"github.com/consensys/gnark/frontend"
	"github.com/consensys/gnark/frontend/cs/r1cs"
	"github.com/consensys/gnark/test"
)

// The tests of backend/accelerated/icicle/groth16/marshal_test.go with the B200 backend in the ICICLE backend's place:
// keys travel between the native and the accelerated package through their serialisation, both provers' proofs verify.

type circuit struct {
	A, B frontend.Variable `gnark:",public"`
	Res  frontend.Variable
}

func (c *circuit) Define(api frontend.API) error {
	api.AssertIsEqual(api.Mul(c.A, c.B), c.Res)
	return nil
}

// committing circuit of backend/groth16/bn254/commitment_test.go (one secret and one public value committed)
type commitCircuit struct {
	One frontend.Variable
	Two frontend.Variable `gnark:",public"`
}

func (c *commitCircuit) Define(api frontend.API) error {
	committer, ok := api.(frontend.Committer)
	if !ok {
		return fmt.Errorf("compiler does not commit")
	}
	commit, err := committer.Commit(c.One, c.Two)
	if err != nil {
		return err
	}
	api.AssertIsDifferent(commit, 0)
	api.AssertIsEqual(c.One, 1)
	api.AssertIsEqual(c.Two, 2)
	return nil
}

var curves = []ecc.ID{ecc.BLS12_377, ecc.BLS12_381, ecc.BN254, ecc.BW6_761}

func testMarshalNativeToB200(t *testing.T, curve ecc.ID) {
	assert := test.NewAssert(t)
	ccs, err := frontend.Compile(curve.ScalarField(), r1cs.NewBuilder, &circuit{})
	assert.NoError(err)
	nativePK, vk, err := native_groth16.Setup(ccs)
	assert.NoError(err)
	accPK := b200_groth16.NewProvingKey(curve)
	buf := new(bytes.Buffer)
	_, err = nativePK.WriteTo(buf)
	assert.NoError(err)
	_, err = accPK.ReadFrom(buf)
	assert.NoError(err)
	if accPK.IsDifferent(nativePK) {
		t.Error("marshal output difference")
	}
	assignment := circuit{A: 3, B: 5, Res: 15}
	w, err := frontend.NewWitness(&assignment, curve.ScalarField())
	assert.NoError(err)
	pw, err := w.Public()
	assert.NoError(err)
	proofNative, err := native_groth16.Prove(ccs, nativePK, w)
	assert.NoError(err)
	proofAcc, err := b200_groth16.Prove(ccs, accPK, w)
	assert.NoError(err)
	assert.NoError(b200_groth16.Verify(proofNative, vk, pw))
	assert.NoError(b200_groth16.Verify(proofAcc, vk, pw))
}

func testMarshalB200ToNative(t *testing.T, curve ecc.ID) {
	assert := test.NewAssert(t)
	ccs, err := frontend.Compile(curve.ScalarField(), r1cs.NewBuilder, &circuit{})
	assert.NoError(err)
	accPK, vk, err := b200_groth16.Setup(ccs)
	assert.NoError(err)
	nativePK := native_groth16.NewProvingKey(curve)
	buf := new(bytes.Buffer)
	_, err = accPK.WriteTo(buf)
	assert.NoError(err)
	_, err = nativePK.ReadFrom(buf)
	assert.NoError(err)
	if accPK.IsDifferent(nativePK) {
		t.Error("marshal output difference")
	}
	assignment := circuit{A: 3, B: 5, Res: 15}
	w, err := frontend.NewWitness(&assignment, curve.ScalarField())
	assert.NoError(err)
	pw, err := w.Public()
	assert.NoError(err)
	proofNative, err := native_groth16.Prove(ccs, nativePK, w)
	assert.NoError(err)
	proofAcc, err := b200_groth16.Prove(ccs, accPK, w, b200.WithDeviceID(0))
	assert.NoError(err)
	assert.NoError(b200_groth16.Verify(proofNative, vk, pw))
	assert.NoError(b200_groth16.Verify(proofAcc, vk, pw))
}

func TestMarshalNativeToB200(t *testing.T) {
	for _, curve := range curves {
		t.Run(fmt.Sprintf("curve=%s", curve.String()), func(t *testing.T) { testMarshalNativeToB200(t, curve) })
	}
}

func TestMarshalB200ToNative(t *testing.T) {
	for _, curve := range curves {
		t.Run(fmt.Sprintf("curve=%s", curve.String()), func(t *testing.T) { testMarshalB200ToNative(t, curve) })
	}
}

// Pedersen / BSB22 commitments through the device (backend/groth16/bn254/commitment_test.go's circuits)
func TestCommitment(t *testing.T) {
	for _, curve := range curves {
		t.Run(fmt.Sprintf("curve=%s", curve.String()), func(t *testing.T) {
			assert := test.NewAssert(t)
			ccs, err := frontend.Compile(curve.ScalarField(), r1cs.NewBuilder, &commitCircuit{})
			assert.NoError(err)
			pk, vk, err := b200_groth16.Setup(ccs)
			assert.NoError(err)
			w, err := frontend.NewWitness(&commitCircuit{One: 1, Two: 2}, curve.ScalarField())
			assert.NoError(err)
			pw, err := w.Public()
			assert.NoError(err)
			proof, err := b200_groth16.Prove(ccs, pk, w)
			assert.NoError(err)
			assert.NoError(b200_groth16.Verify(proof, vk, pw))
		})
	}
}

// the same proof sharded over every device of the box (WithDeviceIDs); skipped on a single-GPU machine by the library's
// error for a device that does not exist
func TestShardedOverTwoDevices(t *testing.T) {
	assert := test.NewAssert(t)
	ccs, err := frontend.Compile(ecc.BN254.ScalarField(), r1cs.NewBuilder, &circuit{})
	assert.NoError(err)
	pk, vk, err := b200_groth16.Setup(ccs)
	assert.NoError(err)
	w, err := frontend.NewWitness(&circuit{A: 3, B: 5, Res: 15}, ecc.BN254.ScalarField())
	assert.NoError(err)
	pw, err := w.Public()
	assert.NoError(err)
	proof, err := b200_groth16.Prove(ccs, pk, w, b200.WithDeviceIDs(0, 1))
	if err != nil {
		t.Skipf("two devices not available: %v", err)
	}
	assert.NoError(b200_groth16.Verify(proof, vk, pw))
}
