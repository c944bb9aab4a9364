// tcgen05 / TMA implicit-GEMM convolution (fp16 in, fp32 accumulate in TMEM).  Placeholder until the
// kernel lands: reports "unsupported" so SMOT_CONV_AUTO uses the SIMT member of the family.
#include "common.cuh"

namespace smot {
bool conv2d_tc_supported(const smot_conv_desc*) { return false; }
int conv2d_tc(const smot_conv_desc*, cudaStream_t) {
  set_error("smot_conv2d: tcgen05 path not available");
  return SMOT_ERR_UNSUPPORTED;
}
}  // namespace smot
