// bn254: G2 MSM kernels over Fp2 (see inst.cuh)
#include "inst.cuh"
namespace gb200 {
GB200_REGISTER_MSM(bn254_g2, 0, 2, bn254_fr, bn254_fp2)
}  // namespace gb200
