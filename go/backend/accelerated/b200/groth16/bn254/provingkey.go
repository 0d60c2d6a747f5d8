// Hand-written source: go/generate.py derives the bls12-377, bls12-381 and bw6-761 packages from this file.

//go:build b200

package bn254

import (
	"sync"

	groth16_bn254 "github.com/consensys/gnark/backend/groth16/bn254"

	"github.com/consensys/gnark/backend/accelerated/b200/internal/cabi"
)

// deviceInfo is what the key holds on the devices: one handle per (device, shard) and the Pedersen commitment keys
// (on the first device). Twin of backend/accelerated/icicle/groth16/bn254/provingkey.go:18-35.
type deviceInfo struct {
	// placement this info was built for
	devices    []int
	precompute bool
	keys       []*cabi.Groth16Key  // keys[i]: shard i of len(devices) on devices[i]
	commitment []*cabi.PedersenKey // CommitmentKeys[i] on devices[0]
}

// ProvingKey embeds the native key so that WriteTo / ReadFrom / WriteDump / ReadDump are inherited and keys are wire
// compatible with the CPU backend (same construction as the ICICLE key, provingkey.go:37-42).
type ProvingKey struct {
	groth16_bn254.ProvingKey
	*deviceInfo
	setupMu sync.Mutex // protects concurrent deviceInfo initialisation
}

// FreeGPUResources releases the device-resident copy of the key (FreeGPUResources, icicle.go:1493-1549). The key can
// be used again afterwards: the next Prove uploads it again. Safe to call more than once.
func (pk *ProvingKey) FreeGPUResources() {
	pk.setupMu.Lock()
	defer pk.setupMu.Unlock()
	pk.freeLocked()
}

func (pk *ProvingKey) freeLocked() {
	if pk.deviceInfo == nil {
		return
	}
	for _, k := range pk.deviceInfo.keys {
		k.Free()
	}
	for _, k := range pk.deviceInfo.commitment {
		k.Free()
	}
	pk.deviceInfo = nil
}
