// bn254: kernel instantiations + registration (see inst.cuh)
#include "inst.cuh"
namespace gb200 {
using h_bn254_fr = HFp<bn254_fr_params>;
using h_bn254_fp = HFp<bn254_fp_params>;
using h_bn254_g2f = Fp2<HFp<bn254_fp_params>, 1>;
GB200_REGISTER_CURVE(0, bn254_fr, bn254_fp, bn254_fp2, h_bn254_fr, h_bn254_fp, h_bn254_g2f)
}  // namespace gb200
