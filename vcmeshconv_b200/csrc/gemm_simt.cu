// fp32 CUDA-core versions of the basis GEMMs (K2, K3, K4) + the elementwise
// activation backward. These are the exact-fp32 arithmetic mode
// (VCM_PREC_FP32) and the fallback for shapes the tcgen05 kernels in
// gemm_tc.cu do not take (K = 17*3 = 51 on the first layer, O = 3 on the
// last, tiny row counts).
//
//   K2  t[r,o]  = sum_{m,i} z[r,m,i] W[m,o,i]            graphVAESSW.py:125-133
//       out     = ELU(t + bias) * sy + res * sr           graphVAESSW.py:135-138,161
//   K3  dz[r,m,i] = sum_o ds[r,o] W[m,o,i]                (autograd of :126)
//   K4  dW[m,o,i] = sum_r ds[r,o] z[r,m,i]                (autograd of :126)
//
// W is consumed in the reference's native parameter layout [M][O][I]
// (`weights.view(M, O, I)`, graphVAESSW.py:125): element (k = m*I + i, o) lives
// at (m*O + o)*I + i, so no transposed copy of the bases is ever made.
#include "adam.cuh"
#include "gemm.h"

namespace vcm {

__device__ __forceinline__ float elu1(float t) { return t > 0.f ? t : expm1f(t); }

// address of basis element (k, o), k = m*I + i
__device__ __forceinline__ size_t basis_addr(int k, int o, int I, int O) {
    const int m = k / I, i = k - m * I;
    return ((size_t)m * O + o) * I + i;
}

constexpr int BM = 128, BN = 64, BK = 16, GT = 256;

struct FwdEpilogue {
    const float *bias;   // may be null
    const float *res;    // may be null
    float *y_act;        // may be null
    float *out;
    int P_out, O, per_point, act;
    float sy, sr;
    __device__ __forceinline__ void operator()(int64_t r, int o, float acc) const { store(r, o, acc, load(r, o)); }
    // split form: the loads of several elements can be issued before the first store
    __device__ __forceinline__ float2 load(int64_t r, int o) const {
        float2 p = make_float2(0.f, 0.f);
        if (bias) p.x = __ldg(bias + (per_point ? (size_t)(r % P_out) * O : 0) + o);
        if (res) p.y = __ldg(res + (size_t)r * O + o);
        return p;
    }
    __device__ __forceinline__ void store(int64_t r, int o, float acc, float2 p) const {
        const float t = acc + p.x;
        const float y = act ? elu1(t) : t;
        const size_t at = (size_t)r * O + o;
        if (y_act) y_act[at] = y;
        out[at] = fmaf(p.y, sr, y * sy);
    }
};

struct DzEpilogue {
    float *dz;
    int K, accumulate;
    __device__ __forceinline__ void operator()(int64_t r, int k, float acc) const { store(r, k, acc, load(r, k)); }
    __device__ __forceinline__ float2 load(int64_t r, int k) const {
        return make_float2(accumulate ? dz[(size_t)r * K + k] : 0.f, 0.f);
    }
    __device__ __forceinline__ void store(int64_t r, int k, float acc, float2 p) const { dz[(size_t)r * K + k] = p.x + acc; }
};

// C[r, n] = sum_k A[r, k] * Bop(k, n);  A row-major [R, Kd].
//   B_KN = true : reduction index is the basis k = (m,i), n = o     (K2)
//   B_KN = false: reduction index is o, n = the basis k = (m,i)     (K3)
template <bool B_KN, typename Epi>
__global__ void __launch_bounds__(GT) gemm_rows_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                       int64_t R, int Kd, int Nd, int I, int O, Epi epi) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;          // 16 x 16 threads, 8 x 4 outputs each
    const int64_t r0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const bool a_vec = (Kd % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < Kd; k0 += BK) {
        // ---- A tile: BM rows x BK reduction, stored transposed ----
        if (a_vec) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int f = tid + GT * u;
                const int row = f >> 2, k4 = (f & 3) * 4;
                const int64_t r = r0 + row;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < R && k0 + k4 < Kd) v = ldg4(A + (size_t)r * Kd + k0 + k4);
                As[k4 + 0][row] = v.x; As[k4 + 1][row] = v.y; As[k4 + 2][row] = v.z; As[k4 + 3][row] = v.w;
            }
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = tid + GT * u;
                const int row = e >> 4, kk = e & 15;
                const int64_t r = r0 + row;
                As[kk][row] = (r < R && k0 + kk < Kd) ? __ldg(A + (size_t)r * Kd + k0 + kk) : 0.f;
            }
        }
        // ---- B tile: BK reduction x BN ----
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + GT * u;
            int kk, nn;
            if (B_KN) { kk = e & 15; nn = e >> 4; }     // contiguous along the basis index i
            else      { nn = e & 63; kk = e >> 6; }
            const int kg = k0 + kk, ng = n0 + nn;
            float v = 0.f;
            if (kg < Kd && ng < Nd) v = B_KN ? __ldg(W + basis_addr(kg, ng, I, O)) : __ldg(W + basis_addr(ng, kg, I, O));
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = r0 + ty * 8 + i;
        if (r >= R) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < Nd) epi(r, n, acc[i][j]);
        }
    }
}

// Skinny GEMM of the 3-channel layers (reduction <= 64, e.g. K = 17*3 = 51 -> O = 32, or O = 32 -> K = 51): the
// 128 x 64 tiled kernel above has 227 CTAs of four barrier rounds each for 28976 rows and is bound by its own
// latencies. Here one thread owns one row: a warp stages its 32 rows with coalesced loads in shared memory (odd row
// stride: conflict-free), the operand tile [Kd][32 outputs] is read as warp-wide broadcasts, 32 accumulators per thread,
// no barrier after the prologue. blockIdx.y walks 32-output chunks.
constexpr int SK_WARPS = 4;
template <bool B_KN, typename Epi>
__global__ void __launch_bounds__(32 * SK_WARPS) gemm_rows_skinny_kernel(const float *__restrict__ A,
                                                                         const float *__restrict__ W, int64_t R, int Kd,
                                                                         int Nd, int I, int O, Epi epi) {
    extern __shared__ __align__(16) float sk_smem[];
    const int ks = (Kd | 1) < 33 ? 33 : (Kd | 1);            // odd stride of the staged rows (>= 33: reused as [32][33])
    float *Ws = sk_smem;                                     // [Kd][32]
    float *As = Ws + Kd * 32 + (threadIdx.x >> 5) * 32 * ks; // this warp's [32][ks]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n0 = blockIdx.y * 32;
    for (int e = threadIdx.x; e < Kd * 32; e += blockDim.x) {
        int kk, nn;
        if (B_KN) { kk = e % Kd; nn = e / Kd; }              // contiguous along the basis index
        else      { nn = e & 31; kk = e >> 5; }
        const int ng = n0 + nn;
        float v = 0.f;
        if (ng < Nd) v = B_KN ? __ldg(W + basis_addr(kk, ng, I, O)) : __ldg(W + basis_addr(ng, kk, I, O));
        Ws[kk * 32 + nn] = v;
    }
    __syncthreads();
    const int64_t n_groups = (R + 31) / 32;
    for (int64_t grp = (int64_t)blockIdx.x * SK_WARPS + warp; grp < n_groups; grp += (int64_t)gridDim.x * SK_WARPS) {
        const int64_t r0 = grp * 32;
        const int n_rows = (int)(R - r0 < 32 ? R - r0 : 32);
        const float *src = A + (size_t)r0 * Kd;
        __syncwarp();
        // coalesced (lanes stride one row), as 4-byte cp.async: all copies of the tile are in flight at once, no
        // load -> store register dependency per element
        {
            const uint32_t as_u = (uint32_t)__cvta_generic_to_shared(As);
            for (int row = 0; row < n_rows; ++row)
                for (int k = lane; k < Kd; k += 32)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(as_u + (uint32_t)(row * ks + k) * 4u),
                                 "l"(src + row * Kd + k) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            for (int row = n_rows; row < 32; ++row)
                for (int k = lane; k < Kd; k += 32) As[row * ks + k] = 0.f;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
        const float *ar = As + lane * ks;
        for (int k = 0; k < Kd; ++k) {
            const float a = ar[k];
            const float4 *w4 = reinterpret_cast<const float4 *>(Ws + k * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const float4 w = w4[j4];
                acc[4 * j4 + 0] = fmaf(a, w.x, acc[4 * j4 + 0]);
                acc[4 * j4 + 1] = fmaf(a, w.y, acc[4 * j4 + 1]);
                acc[4 * j4 + 2] = fmaf(a, w.z, acc[4 * j4 + 2]);
                acc[4 * j4 + 3] = fmaf(a, w.w, acc[4 * j4 + 3]);
            }
        }
        // transpose through the (now free) row tile: lane = output column, so that the epilogue's loads and stores
        // are coalesced 128-byte rows
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) As[lane * 33 + j] = acc[j];
        __syncwarp();
        if (n0 + lane < Nd) {
            int i = 0;
            for (; i + 8 <= n_rows; i += 8) {                // the epilogue's own loads (bias, residual) 8 rows ahead
                float2 pre[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) pre[u] = epi.load(r0 + i + u, n0 + lane);
#pragma unroll
                for (int u = 0; u < 8; ++u) epi.store(r0 + i + u, n0 + lane, As[(i + u) * 33 + lane], pre[u]);
            }
            for (; i < n_rows; ++i) epi(r0 + i, n0 + lane, As[i * 33 + lane]);
        }
    }
}

template <bool B_KN, typename Epi>
int launch_skinny(const float *A, const float *W, int64_t R, int Kd, int Nd, int I, int O, const Epi &epi, cudaStream_t st) {
    const size_t smem = ((size_t)Kd * 32 + (size_t)SK_WARPS * 32 * ((Kd | 1) < 33 ? 33 : (Kd | 1))) * sizeof(float);
    int64_t bx = ceil_div(ceil_div(R, 32), SK_WARPS);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, (unsigned)ceil_div(Nd, 32));
    gemm_rows_skinny_kernel<B_KN, Epi><<<grid, 32 * SK_WARPS, smem, st>>>(A, W, R, Kd, Nd, I, O, epi);
    return VCM_OK;
}
static bool skinny_ok(int Kd, int Nd) {
    static const int on = [] { const char *e = getenv("VCM_SKINNY_GEMM"); return e ? atoi(e) : 1; }();
    return on && Kd <= 64 && Nd <= 128;
}

// K2 for a handful of output channels (last decoder layer, O = 3): one warp per
// row, lanes stride the reduction; the bases sit in shared memory as [o][k].
template <int OT>
__global__ void __launch_bounds__(256) gemm_rows_small_o_kernel(const float *__restrict__ z,
                                                                const float *__restrict__ W, int64_t R, int K, int I,
                                                                int O, FwdEpilogue epi) {
    extern __shared__ float Ws[];   // [O][K]
    for (int e = threadIdx.x; e < O * K; e += blockDim.x) {
        const int o = e / K, k = e - o * K;
        Ws[e] = __ldg(W + basis_addr(k, o, I, O));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int64_t r = (int64_t)blockIdx.x * nw + warp; r < R; r += (int64_t)gridDim.x * nw) {
        const float *zr = z + (size_t)r * K;
        float acc[OT];
#pragma unroll
        for (int o = 0; o < OT; ++o) acc[o] = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float v = __ldg(zr + k);
#pragma unroll
            for (int o = 0; o < OT; ++o)
                if (o < O) acc[o] = fmaf(v, Ws[o * K + k], acc[o]);
        }
#pragma unroll
        for (int o = 0; o < OT; ++o)
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
#pragma unroll
        for (int o = 0; o < OT; ++o)
            if (o < O && lane == o) epi(r, o, acc[o]);
    }
}

// K4: part[s][k][o] = sum_{r in split s} z[r,k] ds[r,o]; 64 x 64 tile, rows in steps of 16.
constexpr int DK = 64, DO = 64, DR = 16;
__global__ void __launch_bounds__(256) gemm_tn_kernel(const float *__restrict__ z, const float *__restrict__ ds,
                                                      float *__restrict__ part, int64_t R, int K, int O,
                                                      int64_t rows_per_split) {
    __shared__ __align__(16) float Zs[DR][DK + 4];
    __shared__ __align__(16) float Ds[DR][DO + 4];
    const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
    const int k0 = blockIdx.x * DK, o0 = blockIdx.y * DO;
    const int64_t rs = (int64_t)blockIdx.z * rows_per_split;
    const int64_t re = rs + rows_per_split < R ? rs + rows_per_split : R;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // register prefetch of the next 16-row slab while the current one is multiplied: the slabs are tiny, the
    // kernel is a chain of dependent global-load latencies otherwise
    float zr[4], dr[4];
    auto fetch = [&](int64_t r0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + 256 * u;
            const int rr = e >> 6, cc = e & 63;
            const int64_t r = r0 + rr;
            zr[u] = (r < re && k0 + cc < K) ? __ldg(z + (size_t)r * K + k0 + cc) : 0.f;
            dr[u] = (r < re && o0 + cc < O) ? __ldg(ds + (size_t)r * O + o0 + cc) : 0.f;
        }
    };
    fetch(rs);
    for (int64_t r0 = rs; r0 < re; r0 += DR) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + 256 * u;
            Zs[e >> 6][e & 63] = zr[u];
            Ds[e >> 6][e & 63] = dr[u];
        }
        __syncthreads();
        if (r0 + DR < re) fetch(r0 + DR);
#pragma unroll
        for (int rr = 0; rr < DR; ++rr) {
            const float4 a = *reinterpret_cast<const float4 *>(&Zs[rr][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Ds[rr][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *dst = part + (size_t)blockIdx.z * K * O;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty * 4 + i;
        if (k >= K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + tx * 4 + j;
            if (o < O) dst[(size_t)k * O + o] = acc[i][j];
        }
    }
}

// dW[m][o][i] = sum_s part[s][k = m*I+i][o], fixed order. Threads walk the partial layout (coalesced reads,
// n_split of them per element); the single write per element goes to dW's native layout.
// 256 threads = 32 elements x 8 groups of splits: every group sums its contiguous share of the splits in order, the
// eight group sums are added in order through shared memory (the chain of dependent loads is n_split / 8 long).
__global__ void reduce_dw_kernel(const float *__restrict__ part, float *__restrict__ dW, int K, int O, int I,
                                 int n_split) {
    __shared__ float red[8][32];
    const int le = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const size_t e = (size_t)blockIdx.x * 32 + le;                    // index into part[s]: k * O + o
    const bool in = e < (size_t)K * O;
    const int per = (n_split + 7) / 8, j0 = grp * per, j1 = min(n_split, j0 + per);
    float s = 0.f;
    if (in)
        for (int j = j0; j < j1; ++j) s += part[(size_t)j * K * O + e];
    red[grp][le] = s;
    __syncthreads();
    if (grp == 0 && in) {
        float tot = red[0][le];
#pragma unroll
        for (int g2 = 1; g2 < 8; ++g2) tot += red[g2][le];
        const int o = (int)(e % O);
        const int k = (int)(e / O);
        const int m = k / I, i = k - m * I;
        dW[((size_t)m * O + o) * I + i] = tot;
    }
}

// ds / dbias / dres in one pass; one thread per (p, o), loop over the batch. The batch loop is blocked by 8
// with all loads issued first: these tensors are small, the kernel is latency-bound, not bandwidth-bound.
__global__ void act_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ y, float *__restrict__ ds,
                               float *__restrict__ dbias, float *__restrict__ dres, int B, int P_out, int O, int act,
                               float sy, float sr) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // p * O + o
    const size_t PO = (size_t)P_out * O;
    if (e >= PO) return;
    float sum = 0.f;
    for (int b0 = 0; b0 < B; b0 += 8) {
        float g[8], yy[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool in = b0 + u < B;
            const size_t at = (size_t)(b0 + u) * PO + e;
            g[u] = in ? __ldg(dout + at) : 0.f;
            yy[u] = (in && act) ? __ldg(y + at) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (b0 + u < B) {
                const size_t at = (size_t)(b0 + u) * PO + e;
                float d = g[u] * sy;
                if (act) d = yy[u] > 0.f ? d : d * (yy[u] + 1.f);
                ds[at] = d;
                sum += d;                                   // fixed order b = 0..B-1
                if (dres) dres[at] = g[u] * sr;
            }
        }
    }
    if (dbias) dbias[e] = sum;
}

// out = (act ? ELU : id)(s + bias) * sy + res * sr   (epilogue of K2 as a stand-alone pass, used when the
// basis projection runs BEFORE the gather: graphVAESSW.py:135-138,161)
__global__ void bias_act_fwd_kernel(const float *__restrict__ s, const float *__restrict__ bias,
                                    const float *__restrict__ res, float *__restrict__ y_act, float *__restrict__ out,
                                    size_t n, int P_out, int O, int per_point, int act, float sy, float sr) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float t = __ldg(s + e);
    if (bias) {
        const size_t po = (size_t)P_out * O;
        t += __ldg(bias + (per_point ? e % po : e % O));
    }
    const float y = act ? elu1(t) : t;
    if (y_act) y_act[e] = y;
    float v = y * sy;
    if (res) v = fmaf(__ldg(res + e), sr, v);
    out[e] = v;
}

// dst[c] = sum_r src[r, c]: one block per 32 columns, fixed-order tree
__global__ void colsum_kernel(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols) {
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int w = threadIdx.x >> 5;
    float s = 0.f;
    if (c < cols)
        for (int r = w; r < rows; r += 8) s += __ldg(src + (size_t)r * cols + c);
    red[w][threadIdx.x & 31] = s;
    __syncthreads();
    if (w == 0 && c < cols) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
        dst[c] = t;
    }
}

// ---- host-side planners shared with gemm_tc.cu ------------------------------
struct DwPlan { int n_split; int64_t rows_per_split; };
static DwPlan plan_dw(int64_t R, int K, int O) {
    const int64_t tiles = ceil_div(K, DK) * ceil_div(O, DO);
    int64_t split = ceil_div(4 * (int64_t)sm_count(), tiles);
    const int64_t max_split = ceil_div(R, 8 * DR);
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    DwPlan p;
    p.rows_per_split = ceil_div(ceil_div(R, split), DR) * DR;
    p.n_split = (int)ceil_div(R, p.rows_per_split);
    if (p.n_split < 1) p.n_split = 1;
    return p;
}

int simt_basis_fwd(const float *z, const float *W, const float *bias, int per_point, const float *res, float *y_act,
                   float *out, int64_t R, int P_out, int M, int I, int O, int act, float sy, float sr,
                   cudaStream_t st) {
    const int K = M * I;
    FwdEpilogue epi{bias, res, y_act, out, P_out, O, per_point, act, sy, sr};
    if (O <= 8 && (size_t)O * K * sizeof(float) <= 96 * 1024) {
        const size_t smem = (size_t)O * K * sizeof(float);
        auto k = O <= 4 ? gemm_rows_small_o_kernel<4> : gemm_rows_small_o_kernel<8>;
        if (smem > 48 * 1024) VCM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t blocks = ceil_div(R, 8);
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        k<<<(unsigned)blocks, 256, smem, st>>>(z, W, R, K, I, O, epi);
    } else if (skinny_ok(K, O)) {
        launch_skinny<true, FwdEpilogue>(z, W, R, K, O, I, O, epi, st);
    } else {
        dim3 grid((unsigned)ceil_div(R, BM), (unsigned)ceil_div(O, BN));
        gemm_rows_kernel<true, FwdEpilogue><<<grid, GT, 0, st>>>(z, W, R, K, O, I, O, epi);
    }
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

int simt_basis_bwd_dz(const float *ds, const float *W, float *dz, int64_t R, int M, int I, int O, int accumulate,
                      cudaStream_t st) {
    const int K = M * I;
    DzEpilogue epi{dz, K, accumulate};
    if (skinny_ok(O, K)) {
        launch_skinny<false, DzEpilogue>(ds, W, R, O, K, I, O, epi, st);
    } else {
        dim3 grid((unsigned)ceil_div(R, BM), (unsigned)ceil_div(K, BN));
        gemm_rows_kernel<false, DzEpilogue><<<grid, GT, 0, st>>>(ds, W, R, O, K, I, O, epi);
    }
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

int reduce_dw_partials(const float *part, float *dW, int K, int O, int I, int n_split, cudaStream_t st) {
    reduce_dw_kernel<<<(unsigned)ceil_div((int64_t)K * O, 32), 256, 0, st>>>(part, dW, K, O, I, n_split);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

size_t simt_basis_bwd_dw_workspace(int64_t R, int M, int I, int O) {
    const DwPlan p = plan_dw(R, M * I, O);
    return (size_t)p.n_split * M * I * O;
}

int simt_basis_bwd_dw(const float *ds, const float *z, float *dW, float *ws, int64_t R, int M, int I, int O,
                      cudaStream_t st) {
    const int K = M * I;
    const DwPlan p = plan_dw(R, K, O);
    dim3 grid((unsigned)ceil_div(K, DK), (unsigned)ceil_div(O, DO), (unsigned)p.n_split);
    gemm_tn_kernel<<<grid, 256, 0, st>>>(z, ds, ws, R, K, O, p.rows_per_split);
    VCM_LAUNCH_CHECK();
    reduce_dw_kernel<<<(unsigned)ceil_div((int64_t)K * O, 32), 256, 0, st>>>(ws, dW, K, O, I, p.n_split);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

}  // namespace vcm

using namespace vcm;

extern "C" {

int vcm_act_bwd(const float *dout, const float *y_act, float *ds, float *dbias, float *dres, int B, int P_out, int O,
                int act, float scale_y, float scale_res, void *stream) {
    VCM_CHECK(B > 0 && P_out >= 0 && O > 0, "bad shape");
    VCM_DEVPTR(dout); VCM_DEVPTR(ds); VCM_DEVPTR_OPT(dbias); VCM_DEVPTR_OPT(dres);
    if (act) VCM_DEVPTR(y_act);
    if (P_out == 0) return VCM_OK;
    act_bwd_kernel<<<(unsigned)ceil_div((int64_t)P_out * O, 256), 256, 0, (cudaStream_t)stream>>>(
        dout, y_act, ds, dbias, dres, B, P_out, O, act, scale_y, scale_res);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

int vcm_bias_act_fwd(const float *s, const float *bias, int bias_per_point, const float *res, float *y_act,
                     float *out, int B, int P_out, int O, int act, float scale_y, float scale_res, void *stream) {
    VCM_CHECK(B > 0 && P_out >= 0 && O > 0, "bad shape");
    VCM_DEVPTR(s); VCM_DEVPTR(out); VCM_DEVPTR_OPT(bias); VCM_DEVPTR_OPT(res); VCM_DEVPTR_OPT(y_act);
    const size_t n = (size_t)B * P_out * O;
    if (n == 0) return VCM_OK;
    bias_act_fwd_kernel<<<(unsigned)ceil_div((int64_t)n, 256), 256, 0, (cudaStream_t)stream>>>(
        s, bias, res, y_act, out, n, P_out, O, bias_per_point, act, scale_y, scale_res);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

int vcm_colsum(const float *src, float *dst, int rows, int cols, void *stream) {
    VCM_CHECK(rows >= 0 && cols > 0, "bad shape");
    VCM_DEVPTR(src); VCM_DEVPTR(dst);
    colsum_kernel<<<(unsigned)ceil_div(cols, 32), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, cols);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Fused Adam over one flat parameter/gradient buffer: the
// reference runs torch.optim.Adam(lr, betas=(0.9, 0.999)) over ~60 tensors
// (graphVAE_train.py:252-262); here all parameters and gradients live in one
// contiguous fp32 buffer (the same one NCCL reduces in place), so the whole
// optimizer step is one launch. Arithmetic follows torch.optim.Adam (no weight
// decay, no amsgrad): p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
namespace vcm {
__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                            float *__restrict__ v, size_t n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt, float grad_scale) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i] * grad_scale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}
}  // namespace vcm

namespace vcm {
// Device-resident optimizer state for CUDA-graph replay: state = {step, lr, 1-b1^t, sqrt(1-b2^t)}.
__global__ void adam_advance_kernel(float *state, float b1, float b2) {
    const float t = state[0] + 1.f;
    state[0] = t;
    state[2] = 1.f - powf(b1, t);
    state[3] = sqrtf(1.f - powf(b2, t));
}
// 16-byte accesses over the aligned body (four streams of 128-bit loads in flight per thread), scalar tail.
__global__ void adam_dev_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                float *__restrict__ v, size_t n, const float *__restrict__ state, float b1, float b2,
                                float eps, float grad_scale) {
    const float lr = state[1], bc1 = state[2], bc2_sqrt = state[3];
    const float step_size = lr / bc1;
    const size_t stride = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    const size_t n4 = vec ? n / 4 : 0;
    for (size_t i = tid; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4 *>(p)[i], mm = reinterpret_cast<float4 *>(m)[i],
               vv = reinterpret_cast<float4 *>(v)[i];
        const float4 gg = reinterpret_cast<const float4 *>(g)[i];
        adam_one(pp.x, gg.x, mm.x, vv.x, b1, b2, eps, step_size, bc2_sqrt, grad_scale);
        adam_one(pp.y, gg.y, mm.y, vv.y, b1, b2, eps, step_size, bc2_sqrt, grad_scale);
        adam_one(pp.z, gg.z, mm.z, vv.z, b1, b2, eps, step_size, bc2_sqrt, grad_scale);
        adam_one(pp.w, gg.w, mm.w, vv.w, b1, b2, eps, step_size, bc2_sqrt, grad_scale);
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
        reinterpret_cast<float4 *>(p)[i] = pp;
    }
    for (size_t i = 4 * n4 + tid; i < n; i += stride) adam_one(p[i], g[i], m[i], v[i], b1, b2, eps, step_size, bc2_sqrt, grad_scale);
}
}  // namespace vcm

extern "C" int vcm_adam_advance(float *state, float beta1, float beta2, void *stream) {
    VCM_DEVPTR(state);
    vcm::adam_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, beta1, beta2);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

extern "C" int vcm_adam_step_graph(float *p, const float *g, float *m, float *v, size_t n, float *state, float beta1,
                                   float beta2, float eps, float grad_scale, void *stream) {
    if (int e = vcm_adam_advance(state, beta1, beta2, stream)) return e;
    return vcm_adam_apply(p, g, m, v, n, state, beta1, beta2, eps, grad_scale, stream);
}

extern "C" int vcm_adam_apply(float *p, const float *g, float *m, float *v, size_t n, const float *state, float beta1,
                              float beta2, float eps, float grad_scale, void *stream) {
    VCM_DEVPTR(p); VCM_DEVPTR(g); VCM_DEVPTR(m); VCM_DEVPTR(v); VCM_DEVPTR(state);
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return VCM_OK;
    int64_t blocks = vcm::ceil_div((int64_t)n, 256 * 4);
    const int64_t cap = (int64_t)vcm::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    vcm::adam_dev_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, state, beta1, beta2, eps, grad_scale);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

extern "C" int vcm_adam_step(float *p, const float *g, float *m, float *v, size_t n, float lr, float beta1,
                             float beta2, float eps, int step, float grad_scale, void *stream) {
    VCM_CHECK(step >= 1, "Adam step counter starts at 1");
    VCM_DEVPTR(p); VCM_DEVPTR(g); VCM_DEVPTR(m); VCM_DEVPTR(v);
    if (n == 0) return VCM_OK;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    int64_t blocks = vcm::ceil_div((int64_t)n, 256 * 4);
    const int64_t cap = (int64_t)vcm::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    vcm::adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1,
                                                                         bc2_sqrt, grad_scale);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}
