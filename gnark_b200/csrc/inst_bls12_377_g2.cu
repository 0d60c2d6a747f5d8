// bls12_377: G2 MSM kernels over Fp2 (see inst.cuh)
#include "inst.cuh"
namespace gb200 {
GB200_REGISTER_MSM(bls12_377_g2, 2, 2, bls12_377_fr, bls12_377_fp2)
}  // namespace gb200
