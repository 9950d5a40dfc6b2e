// tcgen05 / TMEM implicit-GEMM convolution engine (placeholder until the kernel lands).
#include "fn2_common.cuh"
namespace fn2 {
int conv_tc_eligible(const fn2_conv_desc*, const T4&, const T4&) { return 0; }
int conv_tc_forward(const fn2_conv_desc*, const T4&, const float*, const float*, const T4&, cudaStream_t) {
    set_error("conv_tc: not built");
    return FN2_ERR_INVALID;
}
int conv_tc_packed_floats(const fn2_conv_desc*, int, size_t* floats) { *floats = 0; return FN2_OK; }
int conv_tc_pack(const fn2_conv_desc*, int, const float*, float*, cudaStream_t) { return FN2_OK; }
}  // namespace fn2
