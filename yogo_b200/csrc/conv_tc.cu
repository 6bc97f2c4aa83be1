// tcgen05 implicit-GEMM convolution engine (placeholder until the engine lands: reports
// "unsupported" for every shape so that dispatch goes to the SIMT kernels).
#include "common.cuh"
namespace yg {
bool tc_fwd_supported(int, int, int, int, int) { return false; }
bool tc_dgrad_supported(int, int, int, int, int) { return false; }
bool tc_wgrad_supported(int, int, int, int, int) { return false; }
int conv_fwd_tc(const void*, const float*, void*, int, int, int, int, int, int, int, const FwdEpi&, cudaStream_t) { set_error("tcgen05 conv not built"); return YG_ERR_INVALID; }
int conv_dgrad_tc(const void*, const float*, void*, int, int, int, int, int, int, int, const BwdEpi&, cudaStream_t) { set_error("tcgen05 conv not built"); return YG_ERR_INVALID; }
int conv_wgrad_tc(const void*, const void*, float*, float*, int, int, int, int, int, int, int, float, void*, size_t, cudaStream_t) { set_error("tcgen05 conv not built"); return YG_ERR_INVALID; }
size_t tc_wgrad_workspace(int, int, int, int, int, int, int) { return 0; }
}  // namespace yg
