// C-ABI entry points of the basis GEMMs: argument checks + back-end choice.
#include "gemm.h"

using namespace vcm;

static int check_gemm(int64_t R, int M, int I, int O, int precision) {
    VCM_CHECK(R >= 0 && M > 0 && I > 0 && O > 0, "bad GEMM shape rows=%lld M=%d I=%d O=%d", (long long)R, M, I, O);
    VCM_CHECK((int64_t)M * I < (1 << 24) && R < ((int64_t)1 << 31), "GEMM shape out of range");
    VCM_CHECK(precision >= VCM_PREC_FP32 && precision <= VCM_PREC_TF32X3, "unknown precision %d", precision);
    return VCM_OK;
}

extern "C" {

int vcm_basis_uses_tensor_cores(int64_t rows, int M, int I, int O, int precision) {
    return precision != VCM_PREC_FP32 && tc_supported(TC_FWD, rows, M, I, O, precision) ? 1 : 0;
}

size_t vcm_basis_fwd_workspace(int64_t rows, int M, int I, int O, int precision) {
    if (rows <= 0 || M <= 0 || I <= 0 || O <= 0) return 0;
    if (precision != VCM_PREC_FP32 && tc_supported(TC_FWD, rows, M, I, O, precision))
        return tc_basis_fwd_workspace(rows, M, I, O);
    return 0;
}

int vcm_basis_fwd(const float *z, const float *W, const float *bias, int bias_per_point, const float *res,
                  float *y_act, float *out, float *workspace, int B, int P_out, int M, int I, int O, int act,
                  float scale_y, float scale_res, int precision, void *stream) {
    VCM_CHECK(B > 0 && P_out >= 0, "bad batch / vertex count");
    const int64_t R = (int64_t)B * P_out;
    if (int e = check_gemm(R, M, I, O, precision)) return e;
    VCM_DEVPTR(z); VCM_DEVPTR(W); VCM_DEVPTR(out);
    VCM_DEVPTR_OPT(bias); VCM_DEVPTR_OPT(res); VCM_DEVPTR_OPT(y_act); VCM_DEVPTR_OPT(workspace);
    if (R == 0) return VCM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision != VCM_PREC_FP32 && tc_supported(TC_FWD, R, M, I, O, precision))
        return tc_basis_fwd(z, W, bias, bias_per_point, res, y_act, out, workspace, R, P_out, M, I, O, act, scale_y,
                            scale_res, precision, st);
    return simt_basis_fwd(z, W, bias, bias_per_point, res, y_act, out, R, P_out, M, I, O, act, scale_y, scale_res, st);
}

size_t vcm_basis_bwd_dz_workspace(int64_t rows, int M, int I, int O, int precision) {
    if (rows <= 0 || M <= 0 || I <= 0 || O <= 0) return 0;
    if (precision != VCM_PREC_FP32 && tc_supported(TC_DZ, rows, M, I, O, precision))
        return tc_basis_bwd_dz_workspace(M, I, O);
    return 0;
}

int vcm_basis_bwd_dz(const float *ds, const float *W, float *dz, float *workspace, int64_t rows, int M, int I, int O,
                     int accumulate, int precision, void *stream) {
    if (int e = check_gemm(rows, M, I, O, precision)) return e;
    VCM_DEVPTR(ds); VCM_DEVPTR(W); VCM_DEVPTR(dz); VCM_DEVPTR_OPT(workspace);
    if (rows == 0) return VCM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision != VCM_PREC_FP32 && !accumulate && tc_supported(TC_DZ, rows, M, I, O, precision)) {
        VCM_CHECK(workspace != nullptr, "workspace required for the tensor-core dz GEMM");
        return tc_basis_bwd_dz(ds, W, dz, workspace, rows, M, I, O, accumulate, precision, st);
    }
    return simt_basis_bwd_dz(ds, W, dz, rows, M, I, O, accumulate, st);
}

int vcm_basis_stack(const float *W, float *Wt, int M, int I, int O, void *stream) {
    VCM_CHECK(M > 0 && I > 0 && O > 0, "bad shape M=%d I=%d O=%d", M, I, O);
    VCM_DEVPTR(W); VCM_DEVPTR(Wt);
    return tc_stack_bases(W, Wt, M, I, O, (cudaStream_t)stream);
}

int vcm_basis_bwd_dz_stacked(const float *ds, const float *Wt, float *dz, int64_t rows, int M, int I, int O,
                             int precision, void *stream) {
    if (int e = check_gemm(rows, M, I, O, precision)) return e;
    VCM_DEVPTR(ds); VCM_DEVPTR(Wt); VCM_DEVPTR(dz);
    if (rows == 0) return VCM_OK;
    if (precision == VCM_PREC_FP32 || !tc_supported(TC_DZ, rows, M, I, O, precision))
        return fail(VCM_EUNSUPPORTED, "stacked bases are the tensor-core dz operand; this shape/precision runs "
                                      "vcm_basis_bwd_dz (vcm_basis_bwd_dz_workspace() == 0)");
    return tc_basis_bwd_dz(ds, nullptr, dz, const_cast<float *>(Wt), rows, M, I, O, 0, precision, (cudaStream_t)stream);
}

size_t vcm_basis_bwd_dw_workspace(int64_t rows, int M, int I, int O, int precision) {
    if (rows <= 0 || M <= 0 || I <= 0 || O <= 0) return 0;
    size_t a = simt_basis_bwd_dw_workspace(rows, M, I, O);
    if (precision != VCM_PREC_FP32 && tc_supported(TC_DW, rows, M, I, O, precision)) {
        const size_t b = tc_basis_bwd_dw_workspace(rows, M, I, O, precision);
        a = a > b ? a : b;
    }
    return a;
}

int vcm_basis_bwd_dw(const float *ds, const float *z, float *dW, float *workspace, int64_t rows, int M, int I, int O,
                     int precision, void *stream) {
    if (int e = check_gemm(rows, M, I, O, precision)) return e;
    VCM_DEVPTR(ds); VCM_DEVPTR(z); VCM_DEVPTR(dW); VCM_DEVPTR(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) {
        VCM_CUDA(cudaMemsetAsync(dW, 0, (size_t)M * I * O * sizeof(float), st));
        return VCM_OK;
    }
    if (precision != VCM_PREC_FP32 && tc_supported(TC_DW, rows, M, I, O, precision))
        return tc_basis_bwd_dw(ds, z, dW, workspace, rows, M, I, O, precision, st);
    return simt_basis_bwd_dw(ds, z, dW, workspace, rows, M, I, O, st);
}

int vcm_conv_fwd_fused_kind(const vcm_tables *tables, int B, int M, int I, int O, int precision) {
    const Tables *t = as_tables(tables);
    if (t == nullptr || B <= 0) return 0;
    if (fused_small_supported(t, B, M, I, O, precision)) return 2;
    return fused_fwd_supported(t, B, M, I, O, precision) ? 1 : 0;
}

int vcm_conv_fwd_fused_supported(const vcm_tables *tables, int B, int M, int I, int O, int precision) {
    return vcm_conv_fwd_fused_kind(tables, B, M, I, O, precision) != 0 ? 1 : 0;
}

size_t vcm_conv_fwd_fused_workspace(const vcm_tables *tables, int B, int M, int I, int O, int precision) {
    const Tables *t = as_tables(tables);
    if (t == nullptr || B <= 0 || fused_small_supported(t, B, M, I, O, precision) ||
        !fused_fwd_supported(t, B, M, I, O, precision))
        return 0;
    return fused_fwd_workspace(t, B, M, I, O);
}

int vcm_conv_fwd_fused(const vcm_tables *tables, const float *x, const float *A, const float *W, const float *bias,
                       int bias_per_point, const float *res, float *y_act, float *out, float *z, float *workspace,
                       int B, int M, int I, int O, int act, float scale_y, float scale_res, int precision,
                       void *stream) {
    const Tables *t = as_tables(tables);
    VCM_CHECK(t != nullptr, "tables handle is NULL");
    VCM_CHECK(B > 0, "bad batch");
    if (int e = check_gemm((int64_t)B * t->p_out, M, I, O, precision)) return e;
    VCM_DEVPTR(x); VCM_DEVPTR(A); VCM_DEVPTR(W); VCM_DEVPTR(out);
    VCM_DEVPTR_OPT(bias); VCM_DEVPTR_OPT(res); VCM_DEVPTR_OPT(y_act); VCM_DEVPTR_OPT(z); VCM_DEVPTR_OPT(workspace);
    if (fused_small_supported(t, B, M, I, O, precision)) {
        // scalar operand loads; only the epilogue's float2 accesses need alignment
        VCM_CHECK(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(y_act) |
                    reinterpret_cast<uintptr_t>(res)) & 7) == 0, "out, y_act and res must be 8-byte aligned");
        if (t->p_out == 0) return VCM_OK;
        return fused_small(t, x, A, W, bias, bias_per_point, res, y_act, out, z, B, M, I, O, act, scale_y, scale_res,
                           precision, (cudaStream_t)stream);
    }
    VCM_CHECK(aligned16(x) && aligned16(W) && aligned16(out) && (z == nullptr || aligned16(z)),
              "x, W, out and z must be 16-byte aligned");
    if (!fused_fwd_supported(t, B, M, I, O, precision))
        return fail(VCM_EUNSUPPORTED, "vcm_conv_fwd_fused: unsupported shape/precision (query vcm_conv_fwd_fused_supported)");
    if (t->p_out == 0) return VCM_OK;
    return fused_fwd(t, x, A, W, bias, bias_per_point, res, y_act, out, z, workspace, B, M, I, O, act, scale_y,
                     scale_res, (cudaStream_t)stream);
}

}  // extern "C"
