// bn254: scalar-field kernels (NTT, vector ops) + host group arithmetic (see inst.cuh)
#include "inst.cuh"
namespace gb200 {
using h_fr = HFp<bn254_fr_params>;
using h_fp = HFp<bn254_fp_params>;
using h_g2f = Fp2<HFp<bn254_fp_params>, 1>;
GB200_REGISTER_FR(bn254, 0, bn254_fr, h_fr, h_fp, h_g2f)
}  // namespace gb200
