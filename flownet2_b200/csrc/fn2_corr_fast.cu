// Fast correlation path (placeholder until the TMA-staged register-tiled kernel lands).
#include "fn2_common.cuh"
namespace fn2 {
int corr_fast_eligible(const T4&, const T4&, const T4&, int, int, int, int, int, int) { return 0; }
int corr_fast_workspace(int, int, int, int, int, int, size_t* bytes) { *bytes = 0; return FN2_OK; }
int corr_fast_forward(const T4&, const T4&, const T4&, int, int, void*, size_t, cudaStream_t) {
    set_error("corr_fast: not built");
    return FN2_ERR_INVALID;
}
}  // namespace fn2
