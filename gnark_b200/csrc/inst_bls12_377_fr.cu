// bls12_377: scalar-field kernels (NTT, vector ops) + host group arithmetic (see inst.cuh)
#include "inst.cuh"
namespace gb200 {
using h_fr = HFp<bls12_377_fr_params>;
using h_fp = HFp<bls12_377_fp_params>;
using h_g2f = Fp2<HFp<bls12_377_fp_params>, 5>;
GB200_REGISTER_FR(bls12_377, 2, bls12_377_fr, h_fr, h_fp, h_g2f)
}  // namespace gb200
