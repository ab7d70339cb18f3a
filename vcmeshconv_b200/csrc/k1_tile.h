// K1 (gather + coefficient contraction) with shared-memory staged neighbour rows on the tensor pipe: csrc/gather_fwd_tile.cu.
#pragma once
#include "common.cuh"

namespace vcm {

struct K1TilePlan {
    int nk;              // 8-neighbour k-steps of the widest vertex (table width)
    int cw;              // columns per staged unit: 32 or 64
    int stages;          // units in flight per warp
    int gy;              // CTAs per vertex (split of its units)
    int units_per_cta;
};

// false = this shape is not served (caller keeps the fp32 kernel)
bool k1_tile_plan(const Tables *t, int B, int I, int M, int precision, K1TilePlan *pl);
int k1_tile_launch(const Tables *t, const K1TilePlan &pl, const float *x, const float *A, float *z, int B, int I, int M,
                   cudaStream_t st);

}  // namespace vcm
