package b200

import (
	"fmt"

	"github.com/consensys/gnark/backend"
)

// Config is the configuration for the B200 backend.
//
// It mirrors [github.com/consensys/gnark/backend/accelerated/icicle.Config] minus the fields that select an ICICLE
// backend library (there is one backend here) plus the multi-device placement.
type Config struct {
	// DeviceID is the CUDA device of a single-device proof. It is DeviceIDs[0] when DeviceIDs is set.
	DeviceID int
	// DeviceIDs lists the devices a proof is sharded over (point-range shards of every MSM table). Empty means
	// {DeviceID}.
	DeviceIDs  []int
	ProverOpts []backend.ProverOption
	// PinToGPU keeps the proving key's tables in device memory between proofs. The B200 backend always does that
	// (a key is uploaded once per device); the field is kept so that code written for the ICICLE backend compiles
	// unchanged. Use [ProvingKey.FreeGPUResources] to release a key.
	PinToGPU bool
	// Precompute builds the per-window table slabs at key load (16x the table memory on BN254, removes the serial
	// tail of every MSM). Default true.
	Precompute bool
}

// NewConfig creates a new Config with the given options. If no options are provided, it uses sensible defaults
// (device 0, precomputed tables).
func NewConfig(opts ...Option) (*Config, error) {
	cfg := Config{
		DeviceID:   0,
		Precompute: true,
	}
	for _, o := range opts {
		if o != nil {
			if err := o(&cfg); err != nil {
				return nil, err
			}
		}
	}
	return &cfg, nil
}

// Devices returns the devices of a proof in shard order.
func (c *Config) Devices() []int {
	if len(c.DeviceIDs) > 0 {
		return c.DeviceIDs
	}
	return []int{c.DeviceID}
}

// Option is an option for the B200 backend. If no options are set, then sensible defaults are used (device id 0).
type Option func(*Config) error

// WithDeviceID sets the device to be used by the backend. If this option is not set then device ID 0 is used.
func WithDeviceID(id int) Option {
	return func(c *Config) error {
		if id < 0 {
			return fmt.Errorf("invalid device id %d", id)
		}
		c.DeviceID = id
		c.DeviceIDs = nil
		return nil
	}
}

// WithDeviceIDs shards a proof over several devices of one box: device ids[i] holds point-range shard i of every MSM
// table. At least one id is required and ids must be distinct.
func WithDeviceIDs(ids ...int) Option {
	return func(c *Config) error {
		if len(ids) == 0 {
			return fmt.Errorf("no device ids provided")
		}
		seen := make(map[int]struct{}, len(ids))
		for _, id := range ids {
			if id < 0 {
				return fmt.Errorf("invalid device id %d", id)
			}
			if _, dup := seen[id]; dup {
				return fmt.Errorf("duplicate device id %d", id)
			}
			seen[id] = struct{}{}
		}
		c.DeviceIDs = append([]int(nil), ids...)
		c.DeviceID = ids[0]
		return nil
	}
}

// WithProverOptions sets prover options. See [backend.ProverOption] for details.
func WithProverOptions(opts ...backend.ProverOption) Option {
	return func(c *Config) error {
		if len(opts) == 0 {
			return fmt.Errorf("no prover options provided")
		}
		c.ProverOpts = opts
		return nil
	}
}

// WithPinKeysToGPU is accepted for compatibility with the ICICLE backend's option of the same name. Keys are always
// device resident here, so the option only records the caller's intent.
func WithPinKeysToGPU(pin bool) Option {
	return func(c *Config) error {
		c.PinToGPU = pin
		return nil
	}
}

// WithPrecompute switches the per-window table precomputation on or off (default on). Off keeps the tables at the
// size of the key's point slices at the price of a slower MSM tail.
func WithPrecompute(on bool) Option {
	return func(c *Config) error {
		c.Precompute = on
		return nil
	}
}
