// bw6_761: kernel instantiations + registration (see inst.cuh)
#include "inst.cuh"
namespace gb200 {
using h_bw6_761_fr = HFp<bw6_761_fr_params>;
using h_bw6_761_fp = HFp<bw6_761_fp_params>;
using h_bw6_761_g2f = HFp<bw6_761_fp_params>;
GB200_REGISTER_CURVE(3, bw6_761_fr, bw6_761_fp, bw6_761_fp, h_bw6_761_fr, h_bw6_761_fp, h_bw6_761_g2f)
}  // namespace gb200
