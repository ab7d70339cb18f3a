// Internal interface between the C-ABI dispatch (api.cu) and the two GEMM back ends.
#pragma once
#include "common.cuh"

namespace vcm {

// fp32 CUDA-core kernels (gemm_simt.cu)
int simt_basis_fwd(const float *z, const float *W, const float *bias, int per_point, const float *res, float *y_act,
                   float *out, int64_t R, int P_out, int M, int I, int O, int act, float sy, float sr,
                   cudaStream_t st);
int simt_basis_bwd_dz(const float *ds, const float *W, float *dz, int64_t R, int M, int I, int O, int accumulate,
                      cudaStream_t st);
size_t simt_basis_bwd_dw_workspace(int64_t R, int M, int I, int O);
int simt_basis_bwd_dw(const float *ds, const float *z, float *dW, float *ws, int64_t R, int M, int I, int O,
                      cudaStream_t st);

// tcgen05 / TMEM kernels (gemm_tc.cu)
enum { TC_FWD = 0, TC_DZ = 1, TC_DW = 2 };
bool tc_supported(int which, int64_t R, int M, int I, int O, int precision);
size_t tc_basis_fwd_workspace(int64_t R, int M, int I, int O);
int tc_basis_fwd(const float *z, const float *W, const float *bias, int per_point, const float *res, float *y_act,
                 float *out, float *ws, int64_t R, int P_out, int M, int I, int O, int act, float sy, float sr,
                 int precision, cudaStream_t st);
size_t tc_basis_bwd_dz_workspace(int M, int I, int O);
int tc_stack_bases(const float *W, float *Wt, int M, int I, int O, cudaStream_t st);
int tc_basis_bwd_dz(const float *ds, const float *W, float *dz, float *ws, int64_t R, int M, int I, int O,
                    int accumulate, int precision, cudaStream_t st);
int reduce_dw_partials(const float *part, float *dW, int K, int O, int I, int n_split, cudaStream_t st);
size_t tc_basis_bwd_dw_workspace(int64_t R, int M, int I, int O, int precision);
int tc_basis_bwd_dw(const float *ds, const float *z, float *dW, float *ws, int64_t R, int M, int I, int O,
                    int precision, cudaStream_t st);

int splitk_fwd_epilogue(const float *part, int n_split, int64_t R, int O, int P_out, const float *bias, int per_point,
                        const float *res, float *y_act, float *out, int act, float sy, float sr, cudaStream_t st);

// K1 + K2 fused (fused_fwd.cu)
bool fused_fwd_supported(const Tables *t, int B, int M, int I, int O, int precision);
size_t fused_fwd_workspace(const Tables *t, int B, int M, int I, int O);
int fused_fwd(const Tables *t, const float *x, const float *A, const float *W, const float *bias, int per_point,
              const float *res, float *y_act, float *out, float *z, float *ws, int B, int M, int I, int O, int act,
              float sy, float sr, cudaStream_t st);

// K1 + K2 fused for the 3-channel layers, one warp per vertex (fused_fwd_small.cu)
bool fused_small_supported(const Tables *t, int B, int M, int I, int O, int precision);
int fused_small(const Tables *t, const float *x, const float *A, const float *W, const float *bias, int per_point,
                const float *res, float *y_act, float *out, float *z, int B, int M, int I, int O, int act, float sy,
                float sr, int precision, cudaStream_t st);

}  // namespace vcm
