// One Adam update, shared by the fused optimizer kernels (gemm_simt.cu) and the reduce + Adam + broadcast kernel
// (collective.cu). Every operation is an explicit round-to-nearest intrinsic, so the compiler cannot contract the
// expression differently in different kernels: a single-GPU step and a multi-GPU step that sum the same gradients
// produce bit-identical parameters. Arithmetic of torch.optim.Adam (no weight decay, no amsgrad):
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#pragma once

namespace vcm {

__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, float b1, float b2, float eps,
                                         float step_size, float bc2_sqrt, float grad_scale) {
    const float gi = __fmul_rn(g, grad_scale);
    const float mi = __fmaf_rn(b1, m, __fmul_rn(1.f - b1, gi));
    const float vi = __fmaf_rn(b2, v, __fmul_rn(__fmul_rn(1.f - b2, gi), gi));
    m = mi;
    v = vi;
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), eps);
    p = __fsub_rn(p, __fmul_rn(step_size, __fdiv_rn(mi, denom)));
}

}  // namespace vcm
