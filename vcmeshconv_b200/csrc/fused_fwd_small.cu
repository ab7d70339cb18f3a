// K1 + K2 in one launch for the 3-channel layers (LASMConvssw.forward, graphVAESSW.py:111-161, at in_channel = 3):
//
//   z[b,p,m,i] = sum_n A[p,n,m] x[b,idx[p,n],i]                      (:123)
//   out[b,p,o] = epi( sum_{m,i} W[m,o,i] z[b,p,m,i] + bias )         (:125-161)
//
// With I = 3 a vertex carries 48 columns at B = 16 and 51 values of z per batch entry: both products are tiny GEMMs
// ([16 b x I] x [n <= 64] x [M <= 24], then [16 b] x [M*I <= 72] x [O]) and the two-kernel path pays two launches,
// a thread-per-row GEMM with 6 warps per SM and a round trip of z through memory on the forward chain. Here ONE WARP
// owns a vertex and 16 batch entries at a time:
//   stage 1  mma.sync m16n8k8 tf32, M = columns (b, i), N = bases, K = neighbours; x and coefficient fragments are
//            loaded straight from global memory (L2-resident, every load of a k-step independent);
//   re-layout the accumulators go to the warp's shared tile as z rows [b][m*I + i] - the layout of z in memory, so the
//            copy kept for the backward pass leaves as coalesced rows, and the layout stage 2 wants for its A operand;
//   stage 2  M = batch entries, N = output channels, K = (m, i); the stacked bases sit in shared memory once per CTA;
//   epilogue bias, ELU, residual blend exactly as K2's (csrc/gemm_tc.cu), from the accumulator fragments.
// X3 = fp32-grade arithmetic (hi*hi + lo*hi + hi*lo with lo = v - trunc_tf32(v)) for the tf32x3 mode. The fp32 mode
// keeps K1 + K2. Padding neighbours load x = 0 and coefficient = 0 (the reference's mask); no atomics.
#include <stdlib.h>

#include "gemm.h"

namespace vcm {
namespace {

constexpr int kWarpsF = 4;
constexpr int kZS = 76;            // z tile row stride (words): g*76 mod 32 are eight different multiples of 4
constexpr int kKKmax = 9;          // M*I <= 72

__device__ __forceinline__ void mma_tf32_f(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lo_of(uint32_t a) {
    return __float_as_uint(__uint_as_float(a) - __uint_as_float(a & 0xFFFFE000u));
}
template <bool X3>
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    mma_tf32_f(d, a, b0, b1);
    if (X3) {
        uint32_t al[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) al[k] = lo_of(a[k]);
        mma_tf32_f(d, al, b0, b1);
        mma_tf32_f(d, a, lo_of(b0), lo_of(b1));
    }
}
__device__ __forceinline__ float elu_f(float t) { return t > 0.f ? t : expm1f(t); }

struct FusedSmallParams {
    const int32_t *idx_sorted;
    const int2 *meta_sorted;
    const float *x, *A, *W, *bias, *res;
    float *y_act, *out, *z;
    int p_in, p_out, n_max, B, M, O, per_point, act;
    float sy, sr;
};

// I: channels (compile time: the (b, i) <-> column arithmetic); OB: 8-channel blocks of the output (O = 8*OB)
template <int I, int OB, bool X3>
__global__ void __launch_bounds__(kWarpsF * 32)
fused_fwd_small_kernel(const FusedSmallParams P) {
    constexpr int O = OB * 8, WS = O + 8;                       // (t*WS + g) mod 32 is a permutation
    __shared__ float w_s[kKKmax * 8][WS];                       // stacked bases Wst[k = m*I + i][o], zero rows >= M*I
    __shared__ float z_s[kWarpsF][16][kZS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int M = P.M, MI = M * I, KK = (MI + 7) >> 3;
    for (int e = threadIdx.x; e < KK * 8 * O; e += kWarpsF * 32) {
        const int k = e / O, o = e - k * O;
        const int m = k / I, i = k - m * I;
        w_s[k][o] = k < MI ? __ldg(P.W + ((size_t)m * O + o) * I + i) : 0.f;
    }
    float(*zw)[kZS] = z_s[warp];
    for (int e = lane; e < 16 * kZS; e += 32) zw[e / kZS][e % kZS] = 0.f;   // columns >= M*I stay zero
    __syncthreads();
    const int slot = blockIdx.x * kWarpsF + warp;
    if (slot >= P.p_out) return;                                 // warp-uniform; no CTA-wide barrier below
    const int2 meta = __ldg(P.meta_sorted + slot);
    const int p = meta.x, l = meta.y;
    const int NK = (l + 7) >> 3;
    // neighbour ids of the vertex, two per lane (n = lane, lane + 32); -1 = padding
    const int32_t *ip = P.idx_sorted + (size_t)slot * P.n_max;
    int id0 = lane < l ? __ldg(ip + lane) : P.p_in, id1 = lane + 32 < l ? __ldg(ip + lane + 32) : P.p_in;
    id0 = id0 == P.p_in ? -1 : id0;
    id1 = id1 == P.p_in ? -1 : id1;
    const float *Ap = P.A + (size_t)p * P.n_max * M;

    for (int bb = 0; bb < P.B; bb += 16) {
        // ---- stage 1: z[c][m] over the neighbours, c = b_local*I + i --------------------------------------------
        float zacc[I][3][4];
#pragma unroll
        for (int cb = 0; cb < I; ++cb)
#pragma unroll
            for (int mb = 0; mb < 3; ++mb)
#pragma unroll
                for (int k = 0; k < 4; ++k) zacc[cb][mb][k] = 0.f;
        // rows of this thread's x fragments: c = cb*16 + g (+ 8) -> pointer to x[b, 0, i]
        const float *xr[I][2];
        bool rv[I][2];
#pragma unroll
        for (int cb = 0; cb < I; ++cb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = cb * 16 + g + 8 * h;
                const int bl = c / I, i = c - bl * I;
                rv[cb][h] = bb + bl < P.B;
                xr[cb][h] = P.x + (size_t)(rv[cb][h] ? bb + bl : 0) * P.p_in * I + i;
            }
#pragma unroll 2
        for (int ks = 0; ks < NK; ++ks) {
            int q[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = ks * 8 + t + 4 * h;
                const int v0 = __shfl_sync(0xffffffffu, id0, n & 31), v1 = __shfl_sync(0xffffffffu, id1, n & 31);
                q[h] = n < 32 ? v0 : v1;
            }
            uint32_t cf[3][2];                                   // coefficients: (k = n, col = m = mb*8 + g)
#pragma unroll
            for (int mb = 0; mb < 3; ++mb)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int m = mb * 8 + g;
                    cf[mb][h] = (q[h] >= 0 && m < M) ? __float_as_uint(__ldg(Ap + (size_t)(ks * 8 + t + 4 * h) * M + m)) : 0u;
                }
            uint32_t xa[I][4];                                   // a0 (row g, n t), a1 (g + 8, t), a2 (g, t + 4), a3 (g + 8, t + 4)
#pragma unroll
            for (int cb = 0; cb < I; ++cb)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int rh = k & 1, nh = k >> 1;
                    xa[cb][k] = (rv[cb][rh] && q[nh] >= 0) ? __float_as_uint(__ldg(xr[cb][rh] + (size_t)q[nh] * I)) : 0u;
                }
#pragma unroll
            for (int cb = 0; cb < I; ++cb)
#pragma unroll
                for (int mb = 0; mb < 3; ++mb) mma3<X3>(zacc[cb][mb], xa[cb], cf[mb][0], cf[mb][1]);
        }
        // ---- accumulators -> z rows [b_local][m*I + i] -----------------------------------------------------------
        __syncwarp();                                            // the previous batch block's readers are done
#pragma unroll
        for (int cb = 0; cb < I; ++cb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = cb * 16 + g + 8 * h;
                const int bl = c / I, i = c - bl * I;
#pragma unroll
                for (int mb = 0; mb < 3; ++mb)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int m = mb * 8 + 2 * t + j;
                        if (m < M) zw[bl][m * I + i] = zacc[cb][mb][2 * h + j];
                    }
            }
        __syncwarp();
        if (P.z != nullptr) {
            for (int bl = 0; bl < 16 && bb + bl < P.B; ++bl) {
                float *zr = P.z + ((size_t)(bb + bl) * P.p_out + p) * MI;
                for (int k = lane; k < MI; k += 32) zr[k] = zw[bl][k];
            }
        }
        // ---- stage 2: y[b][o] = z[b][:] Wst[:][o] -----------------------------------------------------------------
        float yacc[OB][4];
#pragma unroll
        for (int ob = 0; ob < OB; ++ob)
#pragma unroll
            for (int k = 0; k < 4; ++k) yacc[ob][k] = 0.f;
        for (int kk = 0; kk < KK; ++kk) {
            uint32_t za[4];
            za[0] = __float_as_uint(zw[g][kk * 8 + t]);
            za[1] = __float_as_uint(zw[g + 8][kk * 8 + t]);
            za[2] = __float_as_uint(zw[g][kk * 8 + t + 4]);
            za[3] = __float_as_uint(zw[g + 8][kk * 8 + t + 4]);
#pragma unroll
            for (int ob = 0; ob < OB; ++ob)
                mma3<X3>(yacc[ob], za, __float_as_uint(w_s[kk * 8 + t][ob * 8 + g]),
                         __float_as_uint(w_s[kk * 8 + t + 4][ob * 8 + g]));
        }
        // ---- epilogue (K2's): rows b = bb + g (+ 8), channels ob*8 + 2t, + 1 ----------------------------------------
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b = bb + g + 8 * h;
            if (b >= P.B) continue;
            const size_t row = ((size_t)b * P.p_out + p) * O;
#pragma unroll
            for (int ob = 0; ob < OB; ++ob) {
                const int o = ob * 8 + 2 * t;
                float v0 = yacc[ob][2 * h], v1 = yacc[ob][2 * h + 1];
                if (P.bias != nullptr) {
                    const float *bp = P.bias + (P.per_point ? (size_t)p * O : 0) + o;
                    v0 += __ldg(bp); v1 += __ldg(bp + 1);
                }
                if (P.act) { v0 = elu_f(v0); v1 = elu_f(v1); }
                if (P.y_act != nullptr) *reinterpret_cast<float2 *>(P.y_act + row + o) = make_float2(v0, v1);
                float2 ov = make_float2(v0 * P.sy, v1 * P.sy);
                if (P.res != nullptr) {
                    const float2 rr = __ldg(reinterpret_cast<const float2 *>(P.res + row + o));
                    ov.x = fmaf(rr.x, P.sr, ov.x); ov.y = fmaf(rr.y, P.sr, ov.y);
                }
                *reinterpret_cast<float2 *>(P.out + row + o) = ov;
            }
        }
    }
}

template <int I, int OB>
int launch_fused_small(const FusedSmallParams &P, bool x3, cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div(P.p_out, kWarpsF);
    if (x3) fused_fwd_small_kernel<I, OB, true><<<grid, kWarpsF * 32, 0, st>>>(P);
    else fused_fwd_small_kernel<I, OB, false><<<grid, kWarpsF * 32, 0, st>>>(P);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

}  // namespace

bool fused_small_supported(const Tables *t, int B, int M, int I, int O, int precision) {
    static const int on = [] { const char *e = getenv("VCM_FUSED_SMALL"); return e ? atoi(e) : 1; }();
    if (!on || t == nullptr || precision == VCM_PREC_FP32 || B < 1) return false;
    if (I != 3 || (O != 32 && O != 64)) return false;
    if (M < 1 || M > 24 || M * I > kKKmax * 8 || t->n_max < 1 || t->n_max > 64 || t->p_out < 1) return false;
    return true;
}

int fused_small(const Tables *t, const float *x, const float *A, const float *W, const float *bias, int per_point,
                const float *res, float *y_act, float *out, float *z, int B, int M, int I, int O, int act, float sy,
                float sr, int precision, cudaStream_t st) {
    FusedSmallParams P;
    P.idx_sorted = t->idx_sorted;
    P.meta_sorted = reinterpret_cast<const int2 *>(t->meta_sorted);
    P.x = x; P.A = A; P.W = W; P.bias = bias; P.res = res; P.y_act = y_act; P.out = out; P.z = z;
    P.p_in = t->p_in; P.p_out = t->p_out; P.n_max = t->n_max; P.B = B; P.M = M; P.O = O; P.per_point = per_point;
    P.act = act; P.sy = sy; P.sr = sr;
    const bool x3 = precision == VCM_PREC_TF32X3;
    (void)I;
    return O == 32 ? launch_fused_small<3, 4>(P, x3, st) : launch_fused_small<3, 8>(P, x3, st);
}

}  // namespace vcm
