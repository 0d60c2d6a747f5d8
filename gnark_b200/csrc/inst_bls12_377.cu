// bls12_377: kernel instantiations + registration (see inst.cuh)
#include "inst.cuh"
namespace gb200 {
using h_bls12_377_fr = HFp<bls12_377_fr_params>;
using h_bls12_377_fp = HFp<bls12_377_fp_params>;
using h_bls12_377_g2f = Fp2<HFp<bls12_377_fp_params>, 5>;
GB200_REGISTER_CURVE(2, bls12_377_fr, bls12_377_fp, bls12_377_fp2, h_bls12_377_fr, h_bls12_377_fp, h_bls12_377_g2f)
}  // namespace gb200
