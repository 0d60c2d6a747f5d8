// bw6_761: scalar-field kernels (NTT, vector ops) + host group arithmetic (see inst.cuh)
#include "inst.cuh"
namespace gb200 {
using h_fr = HFp<bw6_761_fr_params>;
using h_fp = HFp<bw6_761_fp_params>;
using h_g2f = HFp<bw6_761_fp_params>;
GB200_REGISTER_FR(bw6_761, 3, bw6_761_fr, h_fr, h_fp, h_g2f)
}  // namespace gb200
