// K5 (coefficient gradient dA + per-edge input gradient dG) with TMA-staged dz tiles: csrc/coeff_grad_tma.cu.
#pragma once
#include "common.cuh"

namespace vcm {

struct K5TmaPlan {
    int nt16;            // 16-neighbour blocks of the widest vertex (table width)
    int cw;              // columns (b fixed, i..i+cw) per staged unit: 32 or 64
    int warps;           // warps per CTA, each with its own ring of `stages` units
    int stages;
    int gy;              // CTAs per vertex (split of its units; > 1 => dA partials in the workspace)
    int units_per_cta;
};

// false = this shape is not served (caller keeps the mma.sync / fp32 kernels)
bool k5_tma_plan(const Tables *t, int B, int I, int M, int precision, K5TmaPlan *pl);
size_t k5_tma_workspace(const Tables *t, const K5TmaPlan &pl, int M);
int k5_tma_launch(const Tables *t, const K5TmaPlan &pl, const float *dz, const float *x, const float *A, float *dA,
                  float *dG, float *ws, int B, int I, int M, cudaStream_t st, bool defer_sum = false);
// defer_sum: with gy > 1 leave the gy partial slabs of dA in the workspace (the caller sums them off the chain)

// sums `n_split` slabs of n floats in fixed order (gather.cu)
void launch_sum_splits(const float *part, float *out, size_t n, int n_split, cudaStream_t st);

}  // namespace vcm
