// tcgen05 / TMEM basis GEMMs (placeholder until the tensor-core kernels land:
// reports "no kernel" so api.cu routes every shape to the fp32 CUDA-core path).
#include "gemm.h"

namespace vcm {

bool tc_supported(int, int64_t, int, int, int, int) { return false; }
int tc_basis_fwd(const float *, const float *, const float *, int, const float *, float *, float *, int64_t, int, int,
                 int, int, int, float, float, int, cudaStream_t) {
    return fail(VCM_EUNSUPPORTED, "tensor-core forward GEMM not built");
}
int tc_basis_bwd_dz(const float *, const float *, float *, int64_t, int, int, int, int, int, cudaStream_t) {
    return fail(VCM_EUNSUPPORTED, "tensor-core dz GEMM not built");
}
size_t tc_basis_bwd_dw_workspace(int64_t, int, int, int, int) { return 0; }
int tc_basis_bwd_dw(const float *, const float *, float *, float *, int64_t, int, int, int, int, cudaStream_t) {
    return fail(VCM_EUNSUPPORTED, "tensor-core dW GEMM not built");
}

}  // namespace vcm
