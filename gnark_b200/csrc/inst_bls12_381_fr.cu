// bls12_381: scalar-field kernels (NTT, vector ops) + host group arithmetic (see inst.cuh)
#include "inst.cuh"
namespace gb200 {
using h_fr = HFp<bls12_381_fr_params>;
using h_fp = HFp<bls12_381_fp_params>;
using h_g2f = Fp2<HFp<bls12_381_fp_params>, 1>;
GB200_REGISTER_FR(bls12_381, 1, bls12_381_fr, h_fr, h_fp, h_g2f)
}  // namespace gb200
