// K5 (coefficient gradient dA only) for layers with a handful of channels: csrc/coeff_grad_small.cu.
#pragma once
#include "common.cuh"

namespace vcm {

// true = this shape is served (dA only, tensor-core modes, few columns B*I per vertex)
bool k5_small_ok(const Tables *t, int B, int I, int M, int precision, bool want_dA, bool want_dG);
int k5_small_launch(const Tables *t, const float *dz, const float *x, float *dA, int B, int I, int M, int precision,
                    cudaStream_t st);

}  // namespace vcm
