// bls12_381: G2 MSM kernels over Fp2 (see inst.cuh)
#include "inst.cuh"
namespace gb200 {
GB200_REGISTER_MSM(bls12_381_g2, 1, 2, bls12_381_fr, bls12_381_fp2)
}  // namespace gb200
