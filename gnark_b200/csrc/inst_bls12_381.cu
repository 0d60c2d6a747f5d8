// bls12_381: kernel instantiations + registration (see inst.cuh)
#include "inst.cuh"
namespace gb200 {
using h_bls12_381_fr = HFp<bls12_381_fr_params>;
using h_bls12_381_fp = HFp<bls12_381_fp_params>;
using h_bls12_381_g2f = Fp2<HFp<bls12_381_fp_params>, 1>;
GB200_REGISTER_CURVE(1, bls12_381_fr, bls12_381_fp, bls12_381_fp2, h_bls12_381_fr, h_bls12_381_fp, h_bls12_381_g2f)
}  // namespace gb200
