// K5 for the layers with a handful of channels (the 3-channel input layer, the residual paths that gather xyz):
//
//   dA[p,n,m] = sum_{b,i} dz[b,p,m,i] x[b,idx[p,n],i]          autograd of graphVAESSW.py:123 wrt the coefficients
//
// With I = 3 a vertex has only C = B*I columns (48 at B = 16): the whole per-vertex product is one small GEMM
// [n <= 64] x [C] x [M <= 24]. The generic fp32 kernel maps lanes to columns and pays a 17-row lane reduction through
// shared memory per neighbour (~60 dependent instructions each); here ONE WARP owns a vertex and the product runs on
// mma.sync m16n8k8 (tf32 operands, fp32 accumulate): M = neighbours, N = bases, K = columns. Operands are so small
// (dz of a vertex: B blocks of M*I contiguous floats; x rows: I floats each, L2 resident) that the fragments are
// loaded straight from global memory, every load of a k-step independent of the others; nothing is staged. The
// accumulators leave through a per-warp shared tile as coalesced rows of dA. Padding neighbours and bases >= M load
// zeros, so their dA entries are written as exact zeros, like the reference's masked gradient.
// X3 = fp32-grade (hi*hi + lo*hi + hi*lo with lo = v - trunc_tf32(v)), used by the tf32x3 mode.
// dG is not produced: the layers this serves either need no input gradient (first layer) or get it from
// vcm_scatter_rev_scaled (M = 1). No atomics, fixed summation order.
#include <stdlib.h>

#include "k5_small.h"

namespace vcm {
namespace {

constexpr int kWarpsS = 4;

__device__ __forceinline__ void mma_tf32_s(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lo_part(uint32_t a) {
    return __float_as_uint(__uint_as_float(a) - __uint_as_float(a & 0xFFFFE000u));
}

// NB: 16-neighbour blocks of the table width; MB: 8-basis blocks (1 for M <= 8, 3 for M <= 24)
template <int NB, int MB, bool X3>
__global__ void __launch_bounds__(kWarpsS * 32)
coeff_grad_small_kernel(const int32_t *__restrict__ idx_sorted, const int2 *__restrict__ meta_sorted,
                        const float *__restrict__ dz, const float *__restrict__ x, float *__restrict__ dA, int p_in,
                        int p_out, int n_max, int B, int I, int M) {
    constexpr int RS = MB * 8 + 1;
    __shared__ float tile_s[kWarpsS][NB * 16][RS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int slot = blockIdx.x * kWarpsS + warp;
    if (slot >= p_out) return;                                   // warp-uniform, no CTA-wide barrier below
    const int2 meta = __ldg(meta_sorted + slot);
    const int p = meta.x, l = meta.y;
    const int C = B * I;
    const int nbl = (l + 15) >> 4;                               // blocks that hold real neighbours

    // rows of this thread's A fragments: neighbours nb*16 + g and nb*16 + g + 8 (-1 = none)
    int q[NB][2];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = nb * 16 + g + 8 * h;
            const int v = n < l ? __ldg(idx_sorted + (size_t)slot * n_max + n) : p_in;
            q[nb][h] = v == p_in ? -1 : v;
        }

    float acc[NB][MB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[nb][mb][k] = 0.f;

#pragma unroll 2
    for (int c0 = 0; c0 < C; c0 += 8) {
        // this thread's two columns of the k-step: c0 + t and c0 + t + 4, column = (b, i)
        const float *xc[2], *zc[2];
        bool cv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = c0 + t + 4 * h;
            cv[h] = c < C;
            const int cc = cv[h] ? c : 0;
            const int b = cc / I, i = cc - b * I;
            xc[h] = x + (size_t)b * p_in * I + i;
            zc[h] = dz + ((size_t)b * p_out + p) * M * I + i;
        }
        // operand B: dz[m][c], m = mb*8 + g
        uint32_t zb[MB][2];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = mb * 8 + g;
                zb[mb][h] = (cv[h] && m < M) ? __float_as_uint(__ldg(zc[h] + (size_t)m * I)) : 0u;
            }
        // operand A: x[n][c]; fragment order a0 = (g, t), a1 = (g + 8, t), a2 = (g, t + 4), a3 = (g + 8, t + 4)
        uint32_t xa[NB][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int rh = k & 1, ch = k >> 1;
                xa[nb][k] = (nb < nbl && q[nb][rh] >= 0 && cv[ch])
                                ? __float_as_uint(__ldg(xc[ch] + (size_t)q[nb][rh] * I)) : 0u;
            }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            if (nb >= nbl) continue;                             // warp-uniform
            uint32_t xl[4];
            if (X3) {
#pragma unroll
                for (int k = 0; k < 4; ++k) xl[k] = lo_part(xa[nb][k]);
            }
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                mma_tf32_s(acc[nb][mb], xa[nb], zb[mb][0], zb[mb][1]);
                if (X3) {
                    mma_tf32_s(acc[nb][mb], xl, zb[mb][0], zb[mb][1]);
                    mma_tf32_s(acc[nb][mb], xa[nb], lo_part(zb[mb][0]), lo_part(zb[mb][1]));
                }
            }
        }
    }

    // accumulator fragments (rows g, g + 8; columns 2t, 2t + 1) -> the warp's tile -> coalesced rows of dA
    float(*tw)[RS] = tile_s[warp];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
            const int n = nb * 16 + g, m = mb * 8 + 2 * t;
            tw[n][m] = acc[nb][mb][0];
            tw[n][m + 1] = acc[nb][mb][1];
            tw[n + 8][m] = acc[nb][mb][2];
            tw[n + 8][m + 1] = acc[nb][mb][3];
        }
    __syncwarp();
    float *dst = dA + (size_t)p * n_max * M;
    const int total = n_max * M;
    for (int e = lane; e < total; e += 32) {
        const int n = e / M, m = e - n * M;
        dst[e] = tw[n][m];
    }
}

template <int NB, int MB>
int launch_small(const Tables *t, const float *dz, const float *x, float *dA, int B, int I, int M, bool x3,
                 cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div(t->p_out, kWarpsS);
    const int2 *meta = reinterpret_cast<const int2 *>(t->meta_sorted);
    if (x3)
        coeff_grad_small_kernel<NB, MB, true><<<grid, kWarpsS * 32, 0, st>>>(t->idx_sorted, meta, dz, x, dA, t->p_in,
                                                                             t->p_out, t->n_max, B, I, M);
    else
        coeff_grad_small_kernel<NB, MB, false><<<grid, kWarpsS * 32, 0, st>>>(t->idx_sorted, meta, dz, x, dA, t->p_in,
                                                                              t->p_out, t->n_max, B, I, M);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

}  // namespace

bool k5_small_ok(const Tables *t, int B, int I, int M, int precision, bool want_dA, bool want_dG) {
    static const int on = [] { const char *e = getenv("VCM_K5_SMALL"); return e ? atoi(e) : 1; }();
    if (!on || t == nullptr || precision == VCM_PREC_FP32 || !want_dA || want_dG) return false;
    if (B < 1 || I < 1 || I > 8 || (int64_t)B * I > 256) return false;
    if (M < 1 || M > 24 || t->n_max < 1 || t->n_max > 64 || t->p_out < 1) return false;
    return true;
}

int k5_small_launch(const Tables *t, const float *dz, const float *x, float *dA, int B, int I, int M, int precision,
                    cudaStream_t st) {
    const bool x3 = precision == VCM_PREC_TF32X3;
    const int nb = (t->n_max + 15) / 16;
#define VCM_K5S(NB_)                                                                   \
    (M <= 8 ? launch_small<NB_, 1>(t, dz, x, dA, B, I, M, x3, st)                      \
            : launch_small<NB_, 3>(t, dz, x, dA, B, I, M, x3, st))
    switch (nb) {
        case 1: return VCM_K5S(1);
        case 2: return VCM_K5S(2);
        case 3: return VCM_K5S(3);
        default: return VCM_K5S(4);
    }
#undef VCM_K5S
}

}  // namespace vcm
