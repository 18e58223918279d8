// tcgen05 (5th-gen tensor core) GEMM path -- placeholder until the 3xTF32 kernel lands; reports
// "unsupported" so MUGD_GEMM_AUTO falls through to the exact-fp32 FFMA kernel in gemm_simt.cu.
#include "common.cuh"

namespace mugd {

bool gemm_tc_supported(const mugd_gemm&) { return false; }

int launch_gemm_tc(const DeviceInfo&, const mugd_gemm&, cudaStream_t, int*) {
    set_error("gemm_tc: not built");
    return MUGD_ERR_INVALID;
}

}  // namespace mugd
