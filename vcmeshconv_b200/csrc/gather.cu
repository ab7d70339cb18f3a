// Gather-side kernels of the vc-conv hot path (K1, K5, K6 + residual edge weights).
//
//   K1  z[b,p,m,i]  = sum_n A[p,n,m] x[b,idx[p,n],i]            graphVAESSW.py:111-123
//   K5  dA[p,n,m]   = sum_{b,i} dz[b,p,m,i] x[b,idx[p,n],i]      (autograd of :123 wrt A)
//       dG[b,p,n,i] = sum_m A[p,n,m] dz[b,p,m,i]                 (autograd of :123 wrt G)
//   K6  dx[b,q,i]   = sum_{slot in rev(q)} dG[b,slot,i]          (autograd of :122, index_add_)
//
// Work decomposition (all three): a "column" is one (b, i..i+VEC) strip of a
// vertex row, VEC = 4 floats (128-bit loads/stores) whenever I % 4 == 0. A CTA
// of 256 threads owns TP vertices x CB = 256/TP columns; vertices are taken in
// the degree-bucketed order `perm`, so the TP rows of a CTA (and the warps of
// neighbouring CTAs) have near-equal trip counts. The per-vertex coefficient
// rows A[p,:,:] and neighbour ids are staged once per CTA in shared memory and
// read back as warp-wide broadcasts; neighbour rows of x come straight from
// L2 (x of a layer is a few MB and stays resident) as coalesced >=128 B
// segments. The sentinel id P_in is tested in the loop: no mask tensor, no
// zero-padded copy of x. All reductions have a fixed order: results are
// bitwise reproducible and there are no atomics.
#include "common.cuh"
#include "k1_tile.h"
#include "k5_small.h"
#include "k5_tma.h"

namespace vcm {

constexpr int kThreads = 256;

__host__ __device__ constexpr int round4(int m) { return (m + 3) / 4 * 4; }

struct ColumnGeom {
    int C;      // columns per vertex = B * I / VEC
    int TP;     // vertices per CTA
    int CB;     // columns per CTA pass = kThreads / TP
    int IV;     // I / VEC
};

static ColumnGeom column_geom(int B, int I, int vec, int max_tp, int threads = kThreads) {
    ColumnGeom g;
    g.IV = I / vec;
    g.C = B * g.IV;
    int tp = 1;
    while (tp < max_tp && threads / (tp * 2) >= g.C && threads / (tp * 2) >= 32) tp *= 2;   // CB stays a warp multiple
    g.TP = tp;
    g.CB = threads / tp;
    return g;
}

// ------------------------------------------------------------------ K1 ------
// K1 runs 128-thread CTAs, four per SM: the staging round trip of one CTA (a full L2 latency before its
// barrier) then idles a quarter of the SM's warps instead of half.
constexpr int kFwdThreads = 128;
template <int MT, int VEC>
__global__ void __launch_bounds__(kFwdThreads, (MT > 17 ? 2 : 4))
gather_contract_fwd_kernel(const int32_t *__restrict__ idx_sorted, const int2 *__restrict__ meta_sorted,
                           const float *__restrict__ x, const float *__restrict__ A, float *__restrict__ z, int p_in,
                           int p_out, int n_max, int I, int M, ColumnGeom g) {
    constexpr int MP = round4(MT);
    extern __shared__ __align__(16) float smem[];
    float *A_s = smem;                                               // [TP][n_max][MP]
    int32_t *idx_s = reinterpret_cast<int32_t *>(smem + (size_t)g.TP * n_max * MP);  // [TP][n_max]

    // Work item s = blockIdx.x*TP + v: {vertex, trip count} and the neighbour ids sit at addresses that depend
    // on s only (processing-order tables), so the staging below is ONE round of independent loads + one barrier.
    const int tid = threadIdx.x;
    for (int v = 0; v < g.TP; ++v) {
        const int slot = blockIdx.x * g.TP + v;
        if (slot >= p_out) break;
        const int2 meta = __ldg(meta_sorted + slot);                 // block-uniform broadcast
        const int p = meta.x, l = meta.y;
        const float *Ap = A + (size_t)p * n_max * M;
        for (int e = tid; e < l * MP; e += kFwdThreads) {
            const int n = e / MP, m = e - n * MP;
            A_s[((size_t)v * n_max + n) * MP + m] = m < M ? __ldg(Ap + n * M + m) : 0.f;
        }
        for (int n = tid; n < l; n += kFwdThreads) idx_s[v * n_max + n] = __ldg(idx_sorted + (size_t)slot * n_max + n);
    }
    __syncthreads();

    const int v = tid / g.CB;
    const int c = blockIdx.y * g.CB + (tid - v * g.CB);
    const int slot = blockIdx.x * g.TP + v;
    if (slot >= p_out || c >= g.C) return;
    const int2 meta = __ldg(meta_sorted + slot);
    const int p = meta.x, l = meta.y;
    const int b = c / g.IV, i0 = (c - b * g.IV) * VEC;
    const float *xb = x + (size_t)b * p_in * I + i0;
    const float *Av = A_s + (size_t)v * n_max * MP;
    const int32_t *iv = idx_s + v * n_max;

    // accumulators live as float2 pairs (aligned register pairs) so the contraction is FFMA2 with no repacking
    constexpr int V2 = VEC == 4 ? 2 : 1;
    float2 acc[MT][V2];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < V2; ++j) acc[m][j] = make_float2(0.f, 0.f);

    // No control flow in the loop body (a sentinel neighbour loads nothing and contributes an exact 0), so
    // the compiler can hoist the gathers of several neighbours ahead of the FMAs: the kernel is bound by
    // the latency of these L2 gathers, not by their bandwidth.
#pragma unroll 4
    for (int n = 0; n < l; ++n) {
        const int q = iv[n];
        float2 g2[V2];
        if (VEC == 4) {
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q != p_in) t4 = ldg4(xb + (size_t)q * I);
            g2[0] = make_float2(t4.x, t4.y);
            g2[V2 - 1] = make_float2(t4.z, t4.w);
        } else {
            g2[0] = make_float2(q != p_in ? __ldg(xb + (size_t)q * I) : 0.f, 0.f);
        }
        const float4 *a4 = reinterpret_cast<const float4 *>(Av + n * MP);
        float a[MP];
#pragma unroll
        for (int k = 0; k < MP / 4; ++k) {
            const float4 t = a4[k];
            a[4 * k] = t.x; a[4 * k + 1] = t.y; a[4 * k + 2] = t.z; a[4 * k + 3] = t.w;
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (VEC == 4) {
#pragma unroll
                for (int j = 0; j < V2; ++j) acc[m][j] = fma2(a[m], g2[j], acc[m][j]);
            } else {
                acc[m][0].x = fmaf(a[m], g2[0].x, acc[m][0].x);
            }
        }
    }

    float *zr = z + ((size_t)b * p_out + p) * M * I + i0;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        if (m < M) {
            if (VEC == 4) st4(zr + (size_t)m * I, make_float4(acc[m][0].x, acc[m][0].y, acc[m][V2 - 1].x, acc[m][V2 - 1].y));
            else zr[(size_t)m * I] = acc[m][0].x;
        }
    }
}

// ------------------------------------------------------------------ K5 ------
template <int MT, int VEC, bool HAS_DA, bool HAS_DG>
__global__ void __launch_bounds__(kThreads, (MT > 17 ? 1 : 2))
coeff_grad_kernel(const int32_t *__restrict__ idx_sorted, const int2 *__restrict__ meta_sorted,
                  const float *__restrict__ dz, const float *__restrict__ x, const float *__restrict__ A,
                  float *__restrict__ dA_out, float *__restrict__ dG, int p_in, int p_out, int n_max, int I, int M,
                  ColumnGeom g, int passes_per_split, int n_split) {
    constexpr int MP = round4(MT);
    extern __shared__ __align__(16) float smem[];
    float *A_s = smem;                                                 // [TP][n_max][MP]
    float *part_s = A_s + (size_t)g.TP * n_max * MP;                   // [8 warps][n_max][MT]
    int32_t *idx_s = reinterpret_cast<int32_t *>(part_s + (size_t)8 * n_max * MT);   // [TP][n_max]
    // per-warp scratch of the lane reduction: 2 buffers x [MT rows][36] floats (16-byte aligned rows)
    constexpr int RS = 36;
    float *red_s = reinterpret_cast<float *>(idx_s + ((size_t)g.TP * n_max + 3) / 4 * 4);
    __shared__ int32_t vert_s[8], last_s[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *red_w = red_s + (size_t)warp * 2 * MT * RS;
    for (int e = tid; e < 8 * n_max * MT; e += kThreads) part_s[e] = 0.f;
    for (int v = 0; v < g.TP; ++v) {
        const int slot = blockIdx.x * g.TP + v;
        int p = -1, l = 0;
        if (slot < p_out) {
            const int2 meta = __ldg(meta_sorted + slot);               // block-uniform broadcast
            p = meta.x;
            l = meta.y;
        }
        if (tid == 0) {
            vert_s[v] = p;
            last_s[v] = l;
        }
        if (p < 0) continue;
        const float *Ap = A + (size_t)p * n_max * M;
        for (int e = tid; e < l * MP; e += kThreads) {
            const int n = e / MP, m = e - n * MP;
            A_s[((size_t)v * n_max + n) * MP + m] = m < M ? __ldg(Ap + n * M + m) : 0.f;
        }
        for (int n = tid; n < l; n += kThreads) idx_s[v * n_max + n] = __ldg(idx_sorted + (size_t)slot * n_max + n);
    }
    __syncthreads();

    // CB is a multiple of 32, so a warp never straddles two vertices.
    const int v = tid / g.CB;
    const int p = vert_s[v];
    const int l = last_s[v];
    const float *Av = A_s + (size_t)v * n_max * MP;
    const int32_t *iv = idx_s + v * n_max;
    float *pw = part_s + (size_t)warp * n_max * MT;

    for (int pass = 0; pass < passes_per_split; ++pass) {
        const int c = (blockIdx.y * passes_per_split + pass) * g.CB + (tid - v * g.CB);
        const bool active = p >= 0 && c < g.C;
        // warp-uniform early-out: nothing in this warp to do for this pass
        if (__ballot_sync(0xffffffffu, active) == 0u) continue;
        const int b = active ? c / g.IV : 0, i0 = active ? (c - b * g.IV) * VEC : 0;
        const float *xb = x + (size_t)b * p_in * I + i0;

        constexpr int V2 = VEC == 4 ? 2 : 1;
        float2 d[MT][V2];                                // dz strip as aligned register pairs (FFMA2 operands)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (VEC == 4) {
                float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active && m < M) t4 = ldg4(dz + (((size_t)b * p_out + p) * M + m) * I + i0);
                d[m][0] = make_float2(t4.x, t4.y);
                d[m][V2 - 1] = make_float2(t4.z, t4.w);
            } else {
                d[m][0] = make_float2((active && m < M) ? __ldg(dz + (((size_t)b * p_out + p) * M + m) * I + i0) : 0.f, 0.f);
            }
        }
        // ---- phase A: dG[b,p,n,:] = sum_m A[p,n,m] * dz strip (coefficient row broadcast from smem) ----
        if (HAS_DG) {
            float *dGn = dG + (((size_t)b * p_out + p) * n_max) * I + i0;
            const float *Arow = Av;
#pragma unroll 2
            for (int n = 0; n < l; ++n, dGn += I, Arow += MP) {
                if (iv[n] == p_in) continue;             // warp-uniform; padding never sits below `last` in practice
                const float4 *a4 = reinterpret_cast<const float4 *>(Arow);
                float a[MP];
#pragma unroll
                for (int k = 0; k < MP / 4; ++k) {
                    const float4 t = a4[k];
                    a[4 * k] = t.x; a[4 * k + 1] = t.y; a[4 * k + 2] = t.z; a[4 * k + 3] = t.w;
                }
                if (VEC == 4) {
                    float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        lo = fma2(a[m], d[m][0], lo);
                        hi = fma2(a[m], d[m][V2 - 1], hi);
                    }
                    if (active) st4(dGn, make_float4(lo.x, lo.y, hi.x, hi.y));
                } else {
                    float og = 0.f;
#pragma unroll
                    for (int m = 0; m < MT; ++m) og = fmaf(a[m], d[m][0].x, og);
                    if (active) *dGn = og;
                }
            }
        }
        // ---- phase B: dA[p,n,m] += sum over the strip of dz * gathered x, reduced over the warp's lanes ----
        if (HAS_DA) {
            float *pwn = pw + lane;                      // lane m < MT owns column m of the warp-private slot
            float *buf0 = red_w + lane, *buf1 = buf0 + MT * RS;
            const float4 *row0 = reinterpret_cast<const float4 *>(red_w + lane * RS);
            const float4 *row1 = reinterpret_cast<const float4 *>(red_w + MT * RS + lane * RS);
            auto gather = [&](int q, float2 (&gq)[V2]) {
                if (VEC == 4) {
                    float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (active && q != p_in) t4 = ldg4(xb + (size_t)q * I);
                    gq[0] = make_float2(t4.x, t4.y);
                    gq[V2 - 1] = make_float2(t4.z, t4.w);
                } else {
                    gq[0] = make_float2((active && q != p_in) ? __ldg(xb + (size_t)q * I) : 0.f, 0.f);
                }
            };
            float2 gnext[V2];
            gather(l > 0 ? iv[0] : p_in, gnext);
            for (int n = 0; n < l; ++n, pwn += MT) {
                float2 gv[V2];
#pragma unroll
                for (int j = 0; j < V2; ++j) gv[j] = gnext[j];
                gather(n + 1 < l ? iv[n + 1] : p_in, gnext);     // software pipeline: next row already in flight
                if (MT == 1) {
                    float v = d[0][0].x * gv[0].x;
                    if (VEC == 4) v += d[0][0].y * gv[0].y + d[0][V2 - 1].x * gv[V2 - 1].x + d[0][V2 - 1].y * gv[V2 - 1].y;
#pragma unroll
                    for (int s2 = 16; s2 >= 1; s2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s2);
                    if (lane == 0) *pwn += v;
                } else {
                    // through shared memory: lane l stores its MT partials as column l of an [MT][36] tile
                    // (conflict-free), then lane m sums row m with 8 LDS.128 (quarter-warps hit disjoint banks).
                    // Two buffers alternate so one __syncwarp per neighbour is enough. Fixed order: reproducible.
                    float *buf = (n & 1) ? buf1 : buf0;
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        float sm;
                        if (VEC == 4) {
                            float2 s2 = fma2(d[m][0], gv[0], make_float2(0.f, 0.f));
                            s2 = fma2(d[m][V2 - 1], gv[V2 - 1], s2);
                            sm = s2.x + s2.y;
                        } else {
                            sm = d[m][0].x * gv[0].x;
                        }
                        buf[m * RS] = sm;
                    }
                    __syncwarp();
                    if (lane < MT) {
                        const float4 *row = (n & 1) ? row1 : row0;
                        float tot = 0.f;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float4 t4 = row[k];
                            tot += (t4.x + t4.y) + (t4.z + t4.w);
                        }
                        *pwn += tot;                     // warp-private slot: no race
                    }
                }
            }
        }
    }
    __syncthreads();

    if (!HAS_DA) return;
    // cross-warp reduction in fixed warp order; padding slots get exactly 0
    const int warps_per_v = g.CB / 32;
    for (int e = tid; e < g.TP * n_max * M; e += kThreads) {
        const int vv = e / (n_max * M);
        const int r = e - vv * (n_max * M);
        const int n = r / M, m = r - n * M;
        const int pp = vert_s[vv];
        if (pp < 0) continue;
        float s = 0.f;
        if (n < last_s[vv] && idx_s[vv * n_max + n] != p_in) {
            for (int w = 0; w < warps_per_v; ++w) s += part_s[((size_t)(vv * warps_per_v + w) * n_max + n) * MT + m];
        }
        dA_out[((size_t)blockIdx.y * p_out + pp) * n_max * M + (size_t)n * M + m] = s;
    }
}

__global__ void sum_splits_kernel(const float *__restrict__ part, float *__restrict__ out, size_t n, int n_split) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < n_split; ++k) s += part[(size_t)k * n + i];
    out[i] = s;
}

void launch_sum_splits(const float *part, float *out, size_t n, int n_split, cudaStream_t st) {
    sum_splits_kernel<<<(unsigned)ceil_div((int64_t)n, 256), 256, 0, st>>>(part, out, n, n_split);
    count_launch();
}

// ------------------------------------------------------------------ K6 ------
template <int VEC>
__global__ void __launch_bounds__(kThreads)
scatter_rev_kernel(const int32_t *__restrict__ rev_ptr, const int32_t *__restrict__ rev_slot,
                   const float *__restrict__ dG, float *__restrict__ dx, int p_in, int p_out, int n_max, int I,
                   ColumnGeom g, int accumulate) {
    const int tid = threadIdx.x;
    const int v = tid / g.CB;
    const int q = blockIdx.x * g.TP + v;
    const int c = blockIdx.y * g.CB + (tid - v * g.CB);
    if (q >= p_in || c >= g.C) return;
    const int b = c / g.IV, i0 = (c - b * g.IV) * VEC;
    const float *src = dG + (size_t)b * p_out * n_max * I + i0;
    float *dst = dx + ((size_t)b * p_in + q) * I + i0;
    const int e0 = __ldg(rev_ptr + q), e1 = __ldg(rev_ptr + q + 1);

    float acc[VEC];
    if (accumulate) {
        Vec<VEC> t;
        t.load(dst);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = t.v[j];
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    }
    int e = e0;
    for (; e + 4 <= e1; e += 4) {                         // 4 independent gathers in flight
        Vec<VEC> t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u].load(src + (size_t)__ldg(rev_slot + e + u) * I);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] += t[u].v[j];
    }
    for (; e < e1; ++e) {
        Vec<VEC> t;
        t.load(src + (size_t)__ldg(rev_slot + e) * I);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += t.v[j];
    }
    Vec<VEC> o;
#pragma unroll
    for (int j = 0; j < VEC; ++j) o.v[j] = acc[j];
    o.store(dst);
}

// K6 for one coefficient per edge (M = 1: the residual pool / un-pool of graphVAESSW.py:151-159, the Laplacian and the
// face gathers of the losses): dx[b,q,:] = sum_{slot in rev(q)} coef[slot] * dy[b, slot / N, :]. The per-edge gradient
// dG[b,p,n,:] = coef[p,n] * dy[b,p,:] is never written: the rows of dy (a few MB, L2-resident) are read through the
// reverse list directly - for a pooling layer with ~47 neighbours that is 1/47 of the bytes of dG, twice.
template <int VEC>
__global__ void __launch_bounds__(kThreads)
scatter_rev_scaled_kernel(const int32_t *__restrict__ rev_ptr, const int32_t *__restrict__ rev_slot,
                          const float *__restrict__ coef, const float *__restrict__ dy, float *__restrict__ dx, int p_in,
                          int p_out, int n_max, int I, ColumnGeom g, int accumulate) {
    const int tid = threadIdx.x;
    const int v = tid / g.CB;
    const int q = blockIdx.x * g.TP + v;
    const int c = blockIdx.y * g.CB + (tid - v * g.CB);
    if (q >= p_in || c >= g.C) return;
    const int b = c / g.IV, i0 = (c - b * g.IV) * VEC;
    const float *src = dy + (size_t)b * p_out * I + i0;
    float *dst = dx + ((size_t)b * p_in + q) * I + i0;
    const int e0 = __ldg(rev_ptr + q), e1 = __ldg(rev_ptr + q + 1);
    float acc[VEC];
    if (accumulate) {
        Vec<VEC> t;
        t.load(dst);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = t.v[j];
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    }
    int e = e0;
    for (; e + 4 <= e1; e += 4) {                         // 4 independent gathers in flight
        Vec<VEC> t[4];
        float w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int slot = __ldg(rev_slot + e + u);
            w[u] = __ldg(coef + slot);
            t[u].load(src + (size_t)(slot / n_max) * I);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = fmaf(w[u], t[u].v[j], acc[j]);
    }
    for (; e < e1; ++e) {
        const int slot = __ldg(rev_slot + e);
        const float w = __ldg(coef + slot);
        Vec<VEC> t;
        t.load(src + (size_t)(slot / n_max) * I);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = fmaf(w, t.v[j], acc[j]);
    }
    Vec<VEC> o;
#pragma unroll
    for (int j = 0; j < VEC; ++j) o.v[j] = acc[j];
    o.store(dst);
}

// ------------------------------------------------- residual edge weights ----
// one warp per output vertex, lanes over the neighbour slots (fixed-order butterfly sums)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

__global__ void edge_weights_fwd_kernel(const int32_t *__restrict__ idx, const float *__restrict__ p_nb,
                                        float *__restrict__ pn, int p_in, int p_out, int n_max) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= p_out) return;
    const int32_t *ir = idx + (size_t)p * n_max;
    const float *pr = p_nb + (size_t)p * n_max;
    float s = 0.f;
    for (int n = lane; n < n_max; n += 32) s += __ldg(ir + n) != p_in ? fabsf(__ldg(pr + n)) : 0.f;
    s = warp_sum(s) + 1e-8f;
    for (int n = lane; n < n_max; n += 32)
        pn[(size_t)p * n_max + n] = __ldg(ir + n) != p_in ? fabsf(__ldg(pr + n)) / s : 0.f;
}

__global__ void edge_weights_bwd_kernel(const int32_t *__restrict__ idx, const float *__restrict__ p_nb,
                                        const float *__restrict__ dpn, float *__restrict__ dp, int p_in, int p_out,
                                        int n_max) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= p_out) return;
    const int32_t *ir = idx + (size_t)p * n_max;
    const float *pr = p_nb + (size_t)p * n_max;
    const float *gr = dpn + (size_t)p * n_max;
    float s = 0.f, dot = 0.f;
    for (int n = lane; n < n_max; n += 32)
        if (__ldg(ir + n) != p_in) {
            const float a = fabsf(__ldg(pr + n));
            s += a;
            dot += __ldg(gr + n) * a;
        }
    s = warp_sum(s) + 1e-8f;
    dot = warp_sum(dot) / s;                               // sum_k dpn[k] * pn[k]
    for (int n = lane; n < n_max; n += 32) {
        float v = 0.f;
        if (__ldg(ir + n) != p_in) {
            const float w = __ldg(pr + n);
            const float sg = w > 0.f ? 1.f : (w < 0.f ? -1.f : 0.f);
            v = sg * (__ldg(gr + n) - dot) / s;
        }
        dp[(size_t)p * n_max + n] = v;
    }
}

// ------------------------------------------------------------ dispatch ------
constexpr int kMaxSmem = 200 * 1024;

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) VCM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return VCM_OK;
}

template <int MT, int VEC>
static int launch_fwd(const Tables *t, const float *x, const float *A, float *z, int B, int I, int M,
                      cudaStream_t st) {
    constexpr int MP = round4(MT);
    ColumnGeom g = column_geom(B, I, VEC, 4, kFwdThreads);
    auto bytes = [&](int tp) { return (size_t)tp * t->n_max * (MP + 1) * sizeof(float); };
    while (g.TP > 1 && bytes(g.TP) > (size_t)kMaxSmem) { g.TP /= 2; g.CB *= 2; }
    if (bytes(g.TP) > (size_t)kMaxSmem)
        return fail(VCM_EUNSUPPORTED, "neighbour count %d too large for the staging buffer", t->n_max);
    auto k = gather_contract_fwd_kernel<MT, VEC>;
    if (int e = set_smem(k, bytes(g.TP))) return e;
    dim3 grid((unsigned)ceil_div(t->p_out, g.TP), (unsigned)ceil_div(g.C, g.CB));
    k<<<grid, kFwdThreads, bytes(g.TP), st>>>(t->idx_sorted, reinterpret_cast<const int2 *>(t->meta_sorted), x, A, z, t->p_in,
                                            t->p_out, t->n_max, I, M, g);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

struct GradPlan {
    ColumnGeom g;
    int n_split, passes;
    size_t smem;
};

template <int MT, int VEC>
static int plan_grad(const Tables *t, int B, int I, GradPlan *pl) {
    constexpr int MP = round4(MT);
    ColumnGeom g = column_geom(B, I, VEC, 8);
    auto bytes = [&](int tp) {
        return ((size_t)tp * t->n_max * MP + (size_t)8 * t->n_max * MT + ((size_t)tp * t->n_max + 3) / 4 * 4 +
                (size_t)8 * 2 * MT * 36) * sizeof(float);
    };
    while (g.TP > 1 && bytes(g.TP) > (size_t)kMaxSmem) { g.TP /= 2; g.CB *= 2; }
    if (bytes(g.TP) > (size_t)kMaxSmem)
        return fail(VCM_EUNSUPPORTED, "neighbour count %d too large for the staging buffer", t->n_max);
    // Column passes of a vertex are split over CTAs (grid.y) when there are few vertices; pick the split that
    // loses the least time to wave quantisation: CTAs run in waves of 2 per SM, each does `passes` passes
    // (+ ~1 pass worth of staging and reduction).
    const int64_t bx = ceil_div(t->p_out, g.TP), total_passes = ceil_div(g.C, g.CB);
    const int64_t slots = 2 * (int64_t)sm_count();
    int64_t split = 1;
    double best = 1e30;
    for (int64_t s2 = 1; s2 <= total_passes; ++s2) {
        const int64_t per = ceil_div(total_passes, s2), eff = ceil_div(total_passes, per);
        const double cost = (double)ceil_div(bx * eff, slots) * (double)(per + 1);
        if (cost < best * 0.97) { best = cost; split = eff; }
    }
    pl->g = g;
    pl->passes = (int)ceil_div(total_passes, split);
    pl->n_split = (int)ceil_div(total_passes, pl->passes);
    pl->smem = bytes(g.TP);
    return VCM_OK;
}

template <int MT, int VEC>
static int launch_grad(const Tables *t, const float *dz, const float *x, const float *A, float *dA, float *dG,
                       float *ws, int B, int I, int M, cudaStream_t st) {
    GradPlan pl;
    if (int e = plan_grad<MT, VEC>(t, B, I, &pl)) return e;
    const size_t n_dA = (size_t)t->p_out * t->n_max * M;
    if (dA != nullptr && pl.n_split > 1 && ws == nullptr)
        return fail(VCM_EINVAL, "workspace required (%d splits)", pl.n_split);
    float *target = dA == nullptr ? nullptr : (pl.n_split > 1 ? ws : dA);
    dim3 grid((unsigned)ceil_div(t->p_out, pl.g.TP), (unsigned)pl.n_split);
    const int2 *meta = reinterpret_cast<const int2 *>(t->meta_sorted);
#define VCM_K5_LAUNCH(DA_, DG_)                                                                                    \
    do {                                                                                                           \
        auto k = coeff_grad_kernel<MT, VEC, DA_, DG_>;                                                             \
        if (int e = set_smem(k, pl.smem)) return e;                                                                \
        k<<<grid, kThreads, pl.smem, st>>>(t->idx_sorted, meta, dz, x, A, target, dG, t->p_in, t->p_out, t->n_max, \
                                           I, M, pl.g, pl.passes, pl.n_split);                                     \
    } while (0)
    if (dA != nullptr && dG != nullptr) VCM_K5_LAUNCH(true, true);
    else if (dA != nullptr) VCM_K5_LAUNCH(true, false);
    else VCM_K5_LAUNCH(false, true);
#undef VCM_K5_LAUNCH
    VCM_LAUNCH_CHECK();
    if (dA != nullptr && pl.n_split > 1) {
        sum_splits_kernel<<<(unsigned)ceil_div((int64_t)n_dA, 256), 256, 0, st>>>(ws, dA, n_dA, pl.n_split);
        VCM_LAUNCH_CHECK();
    }
    return VCM_OK;
}

// smallest compiled accumulator height covering M
#define VCM_DISPATCH_M(M, VEC, CALL)                                   \
    do {                                                               \
        if ((M) == 1) { constexpr int MT = 1; return CALL; }           \
        if ((M) <= 8) { constexpr int MT = 8; return CALL; }           \
        if ((M) <= 17) { constexpr int MT = 17; return CALL; }         \
        if ((M) <= 32) { constexpr int MT = 32; return CALL; }         \
        return fail(VCM_EUNSUPPORTED, "weight_num %d > 32 not built", (M)); \
    } while (0)

static int check_common(const Tables *t, int B, int I, int M) {
    VCM_CHECK(t != nullptr, "tables handle is NULL");
    VCM_CHECK(B > 0 && I > 0 && M > 0, "B, I, M must be positive (got %d, %d, %d)", B, I, M);
    VCM_CHECK((int64_t)B * t->p_out * t->n_max * (int64_t)I < ((int64_t)1 << 40), "problem too large");
    return VCM_OK;
}

}  // namespace vcm

using namespace vcm;

extern "C" {

int vcm_gather_contract_fwd(const vcm_tables *h, const float *x, const float *A, float *z, int B, int I, int M,
                            void *stream) {
    const Tables *t = as_tables(h);
    if (int e = check_common(t, B, I, M)) return e;
    VCM_DEVPTR(x); VCM_DEVPTR(A); VCM_DEVPTR(z);
    if (t->p_out == 0) return VCM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = I % 4 == 0 && aligned16(x) && aligned16(z);
    if (v4) VCM_DISPATCH_M(M, 4, (launch_fwd<MT, 4>(t, x, A, z, B, I, M, st)));
    VCM_DISPATCH_M(M, 1, (launch_fwd<MT, 1>(t, x, A, z, B, I, M, st)));
}

size_t vcm_coeff_grad_workspace(const vcm_tables *h, int B, int I, int M) {
    const Tables *t = as_tables(h);
    if (!t || B <= 0 || I <= 0 || M <= 0 || M > 32 || t->p_out == 0) return 0;
    GradPlan pl;
    int e;
    const int mt = M == 1 ? 1 : M <= 8 ? 8 : M <= 17 ? 17 : 32;
    // the plan depends on VEC only through the column count; be conservative: take the larger need
    size_t need = 0;
    for (int vec = 1; vec <= 4; vec += 3) {
        if (vec == 4 && I % 4 != 0) continue;
        switch (mt) {
            case 1: e = vec == 4 ? plan_grad<1, 4>(t, B, I, &pl) : plan_grad<1, 1>(t, B, I, &pl); break;
            case 8: e = vec == 4 ? plan_grad<8, 4>(t, B, I, &pl) : plan_grad<8, 1>(t, B, I, &pl); break;
            case 17: e = vec == 4 ? plan_grad<17, 4>(t, B, I, &pl) : plan_grad<17, 1>(t, B, I, &pl); break;
            default: e = vec == 4 ? plan_grad<32, 4>(t, B, I, &pl) : plan_grad<32, 1>(t, B, I, &pl); break;
        }
        if (e == VCM_OK && pl.n_split > 1) {
            const size_t n = (size_t)pl.n_split * t->p_out * t->n_max * M;
            need = n > need ? n : need;
        }
    }
    return need;
}

int vcm_coeff_grad(const vcm_tables *h, const float *dz, const float *x, const float *A, float *dA, float *dG,
                   float *ws, int B, int I, int M, void *stream) {
    const Tables *t = as_tables(h);
    if (int e = check_common(t, B, I, M)) return e;
    VCM_DEVPTR(dz); VCM_DEVPTR(x); VCM_DEVPTR(A); VCM_DEVPTR_OPT(dA); VCM_DEVPTR_OPT(dG); VCM_DEVPTR_OPT(ws);
    VCM_CHECK(dA != nullptr || dG != nullptr, "nothing to compute: dA and dG are both NULL");
    if (t->p_out == 0) return VCM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = I % 4 == 0 && aligned16(x) && aligned16(dz) && (dG == nullptr || aligned16(dG));
    if (v4) VCM_DISPATCH_M(M, 4, (launch_grad<MT, 4>(t, dz, x, A, dA, dG, ws, B, I, M, st)));
    VCM_DISPATCH_M(M, 1, (launch_grad<MT, 1>(t, dz, x, A, dA, dG, ws, B, I, M, st)));
}

int vcm_scatter_rev(const vcm_tables *h, const float *dG, float *dx, int B, int I, int accumulate, void *stream) {
    const Tables *t = as_tables(h);
    if (int e = check_common(t, B, I, 1)) return e;
    VCM_DEVPTR(dG); VCM_DEVPTR(dx);
    if (t->p_in == 0) return VCM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = I % 4 == 0 && aligned16(dG) && aligned16(dx);
    ColumnGeom g = column_geom(B, I, v4 ? 4 : 1, 8);
    dim3 grid((unsigned)ceil_div(t->p_in, g.TP), (unsigned)ceil_div(g.C, g.CB));
    if (v4)
        scatter_rev_kernel<4><<<grid, kThreads, 0, st>>>(t->rev_ptr, t->rev_slot, dG, dx, t->p_in, t->p_out, t->n_max,
                                                         I, g, accumulate);
    else
        scatter_rev_kernel<1><<<grid, kThreads, 0, st>>>(t->rev_ptr, t->rev_slot, dG, dx, t->p_in, t->p_out, t->n_max,
                                                         I, g, accumulate);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

int vcm_scatter_rev_scaled(const vcm_tables *h, const float *dy, const float *coef, float *dx, int B, int I,
                           int accumulate, void *stream) {
    const Tables *t = as_tables(h);
    if (int e = check_common(t, B, I, 1)) return e;
    VCM_DEVPTR(dy); VCM_DEVPTR(coef); VCM_DEVPTR(dx);
    if (t->p_in == 0) return VCM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = I % 4 == 0 && aligned16(dy) && aligned16(dx);
    ColumnGeom g = column_geom(B, I, v4 ? 4 : 1, 8);
    dim3 grid((unsigned)ceil_div(t->p_in, g.TP), (unsigned)ceil_div(g.C, g.CB));
    if (v4)
        scatter_rev_scaled_kernel<4><<<grid, kThreads, 0, st>>>(t->rev_ptr, t->rev_slot, coef, dy, dx, t->p_in, t->p_out,
                                                                t->n_max, I, g, accumulate);
    else
        scatter_rev_scaled_kernel<1><<<grid, kThreads, 0, st>>>(t->rev_ptr, t->rev_slot, coef, dy, dx, t->p_in, t->p_out,
                                                                t->n_max, I, g, accumulate);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

int vcm_edge_weights_fwd(const vcm_tables *h, const float *p_nb, float *pn, void *stream) {
    const Tables *t = as_tables(h);
    VCM_CHECK(t != nullptr, "tables handle is NULL");
    VCM_DEVPTR(p_nb); VCM_DEVPTR(pn);
    if (t->p_out == 0 || t->n_max == 0) return VCM_OK;
    edge_weights_fwd_kernel<<<(unsigned)ceil_div((int64_t)t->p_out * 32, 128), 128, 0, (cudaStream_t)stream>>>(
        t->idx, p_nb, pn, t->p_in, t->p_out, t->n_max);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

int vcm_edge_weights_bwd(const vcm_tables *h, const float *p_nb, const float *dpn, float *dp, void *stream) {
    const Tables *t = as_tables(h);
    VCM_CHECK(t != nullptr, "tables handle is NULL");
    VCM_DEVPTR(p_nb); VCM_DEVPTR(dpn); VCM_DEVPTR(dp);
    if (t->p_out == 0 || t->n_max == 0) return VCM_OK;
    edge_weights_bwd_kernel<<<(unsigned)ceil_div((int64_t)t->p_out * 32, 128), 128, 0, (cudaStream_t)stream>>>(
        t->idx, p_nb, dpn, dp, t->p_in, t->p_out, t->n_max);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Project-first order for narrow outputs (last decoder layer, 32 -> 3 channels on the full mesh):
//   u[b,q,m,o] = sum_i W[m,o,i] x[b,q,i]                       (one small GEMM, done by the caller)
//   s[b,p,o]   = sum_{n,m} A[p,n,m] u[b, idx[p,n], m, o]        (vcm_project_gather_fwd)
//   dA[p,n,m]  = sum_{b,o} ds[b,p,o] u[b, idx[p,n], m, o]       (vcm_project_gather_bwd)
//   du[b,q,m,o]= sum_{(p,n) in rev(q)} A[p,n,m] ds[b,p,o]       (vcm_project_gather_bwd)
// The op is bilinear, so this equals the gather-first result; it moves rows of M*O floats (51 for O = 3)
// instead of materialising z (M*I = 544 floats per output vertex). One warp owns one (vertex, batch group);
// lane j owns element j = m*O + o of the projected row (MO <= 64: two elements per lane), so every load of
// a u row is one coalesced 4*MO-byte segment. Fixed-order reductions, no atomics.
namespace vcm {

constexpr int PG_MAXN = 8;        // neighbours handled per register pass of the dA kernel
constexpr int PG_UN = 4;          // neighbours / reverse edges / batch entries whose loads are issued together

// All three kernels are latency-bound (a few MB of data, rows of 4*MO bytes): their loops are written branch-free
// in groups of PG_UN so that the loads of a whole group are in flight before the first FMA needs one.
template <int NB>                 // NB = batch entries per warp pass
__global__ void __launch_bounds__(128)
project_gather_fwd_kernel(const int32_t *__restrict__ idx, const int32_t *__restrict__ last,
                          const float *__restrict__ u, const float *__restrict__ A, float *__restrict__ s, int p_in,
                          int p_out, int n_max, int B, int M, int O) {
    __shared__ float red[4][NB][64];
    const int MO = M * O, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int p = blockIdx.x * 4 + w;
    if (p >= p_out) return;
    const int j0 = lane, j1 = lane + 32;
    const bool v0 = j0 < MO, v1 = j1 < MO;
    const int m0 = v0 ? j0 / O : 0, m1 = v1 ? j1 / O : 0;
    const int e0 = v0 ? j0 : 0, e1 = v1 ? j1 : 0;                     // in-range element to load when the lane is idle
    const int l = __ldg(last + p);
    const int32_t *ip = idx + (size_t)p * n_max;
    const float *Ap = A + (size_t)p * n_max * M;
    for (int b0 = blockIdx.y * NB; b0 < B; b0 += gridDim.y * NB) {
        float a0[NB], a1[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) a0[k] = a1[k] = 0.f;
        for (int n0 = 0; n0 < l; n0 += PG_UN) {
            int q[PG_UN];
            float c0[PG_UN], c1[PG_UN];
#pragma unroll
            for (int t = 0; t < PG_UN; ++t) {
                const int n = min(n0 + t, l - 1);
                const int qq = __ldg(ip + n);
                const bool real = n0 + t < l && qq != p_in;
                q[t] = real ? qq : 0;
                c0[t] = real && v0 ? __ldg(Ap + n * M + m0) : 0.f;     // padding / tail: coefficient 0, row 0
                c1[t] = real && v1 ? __ldg(Ap + n * M + m1) : 0.f;
            }
#pragma unroll
            for (int t = 0; t < PG_UN; ++t) {
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    const int bb = min(b0 + k, B - 1);
                    const float *ur = u + ((size_t)bb * p_in + q[t]) * MO;
                    a0[k] = fmaf(c0[t], __ldg(ur + e0), a0[k]);
                    a1[k] = fmaf(c1[t], __ldg(ur + e1), a1[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            red[w][k][j0] = a0[k];
            red[w][k][j1] = a1[k];
        }
        __syncwarp();
        // lane (k, o) sums its M products in fixed order
        for (int e = lane; e < NB * O; e += 32) {
            const int k = e / O, o = e - k * O;
            if (b0 + k < B) {
                float t = 0.f;
                for (int m = 0; m < M; ++m) t += red[w][k][m * O + o];
                s[((size_t)(b0 + k) * p_out + p) * O + o] = t;
            }
        }
        __syncwarp();
    }
}

// dA: one warp per output vertex, all batch entries, neighbours in register passes of PG_MAXN.
__global__ void __launch_bounds__(128)
project_gather_dA_kernel(const int32_t *__restrict__ idx, const int32_t *__restrict__ last,
                         const float *__restrict__ ds, const float *__restrict__ u, float *__restrict__ dA, int p_in,
                         int p_out, int n_max, int B, int M, int O) {
    __shared__ float red[4][64];
    const int MO = M * O, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int p = blockIdx.x * 4 + w;
    if (p >= p_out) return;
    const int j0 = lane, j1 = lane + 32;
    const bool v0 = j0 < MO, v1 = j1 < MO;
    const int e0 = v0 ? j0 : 0, e1 = v1 ? j1 : 0;
    const int o0 = e0 % O, o1 = e1 % O;
    const int l = __ldg(last + p);
    const int32_t *ip = idx + (size_t)p * n_max;
    float *dAp = dA + (size_t)p * n_max * M;
    for (int nb = 0; nb < n_max; nb += PG_MAXN) {
        float a0[PG_MAXN], a1[PG_MAXN];
        int q[PG_MAXN];
        bool real[PG_MAXN];
#pragma unroll
        for (int k = 0; k < PG_MAXN; ++k) {
            a0[k] = a1[k] = 0.f;
            const int qq = nb + k < l ? __ldg(ip + nb + k) : p_in;
            real[k] = qq != p_in;
            q[k] = real[k] ? qq : 0;
        }
        if (nb < l) {
            for (int b0 = 0; b0 < B; b0 += PG_UN) {
                float g0[PG_UN], g1[PG_UN];
#pragma unroll
                for (int t = 0; t < PG_UN; ++t) {
                    const bool in = b0 + t < B;
                    const float *dr = ds + ((size_t)min(b0 + t, B - 1) * p_out + p) * O;
                    g0[t] = in && v0 ? __ldg(dr + o0) : 0.f;             // idle lanes and the batch tail add exact zeros
                    g1[t] = in && v1 ? __ldg(dr + o1) : 0.f;
                }
#pragma unroll
                for (int t = 0; t < PG_UN; ++t) {
                    const float *ub = u + (size_t)min(b0 + t, B - 1) * p_in * MO;
#pragma unroll
                    for (int k = 0; k < PG_MAXN; ++k) {
                        const float *ur = ub + (size_t)q[k] * MO;
                        a0[k] = fmaf(g0[t], __ldg(ur + e0), a0[k]);
                        a1[k] = fmaf(g1[t], __ldg(ur + e1), a1[k]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PG_MAXN; ++k) {
            if (nb + k >= n_max) break;
            red[w][j0] = a0[k];
            red[w][j1] = a1[k];
            __syncwarp();
            if (lane < M) {
                float t = 0.f;
                if (real[k])
                    for (int o = 0; o < O; ++o) t += red[w][lane * O + o];
                dAp[(size_t)(nb + k) * M + lane] = t;                   // padding slots: exactly 0
            }
            __syncwarp();
        }
    }
}

// du through the reverse CSR: one warp per (input vertex, batch entry group).
template <int NB>
__global__ void __launch_bounds__(128)
project_gather_du_kernel(const int32_t *__restrict__ rev_ptr, const int32_t *__restrict__ rev_slot,
                         const float *__restrict__ ds, const float *__restrict__ A, float *__restrict__ du, int p_in,
                         int p_out, int n_max, int B, int M, int O) {
    const int MO = M * O, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * 4 + w;
    if (q >= p_in) return;
    const int j0 = lane, j1 = lane + 32;
    const bool v0 = j0 < MO, v1 = j1 < MO;
    const int x0 = v0 ? j0 : 0, x1 = v1 ? j1 : 0;
    const int m0 = x0 / O, m1 = x1 / O, o0 = x0 % O, o1 = x1 % O;
    const int e0 = __ldg(rev_ptr + q), e1 = __ldg(rev_ptr + q + 1);
    for (int b0 = blockIdx.y * NB; b0 < B; b0 += gridDim.y * NB) {
        float a0[NB], a1[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) a0[k] = a1[k] = 0.f;
        for (int e = e0; e < e1; e += PG_UN) {
            int pp[PG_UN];
            float c0[PG_UN], c1[PG_UN];
#pragma unroll
            for (int t = 0; t < PG_UN; ++t) {
                const bool in = e + t < e1;
                const int slot = __ldg(rev_slot + min(e + t, e1 - 1));
                pp[t] = slot / n_max;
                c0[t] = in && v0 ? __ldg(A + (size_t)slot * M + m0) : 0.f;
                c1[t] = in && v1 ? __ldg(A + (size_t)slot * M + m1) : 0.f;
            }
#pragma unroll
            for (int t = 0; t < PG_UN; ++t) {
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    const float *dr = ds + ((size_t)min(b0 + k, B - 1) * p_out + pp[t]) * O;
                    a0[k] = fmaf(c0[t], __ldg(dr + o0), a0[k]);
                    a1[k] = fmaf(c1[t], __ldg(dr + o1), a1[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            if (b0 + k < B) {
                float *dst = du + ((size_t)(b0 + k) * p_in + q) * MO;
                if (v0) dst[j0] = a0[k];
                if (v1) dst[j1] = a1[k];
            }
        }
    }
}

}  // namespace vcm

extern "C" {

int vcm_project_gather_fwd(const vcm_tables *h, const float *u, const float *A, float *s, int B, int M, int O,
                           void *stream) {
    const Tables *t = as_tables(h);
    VCM_CHECK(t != nullptr, "tables handle is NULL");
    VCM_CHECK(B > 0 && M > 0 && O > 0 && M * O <= 64, "project-first gather needs M*O <= 64 (got M=%d O=%d)", M, O);
    VCM_DEVPTR(u); VCM_DEVPTR(A); VCM_DEVPTR(s);
    if (t->p_out == 0) return VCM_OK;
    dim3 grid((unsigned)ceil_div(t->p_out, 4), (unsigned)ceil_div(B, 4));
    project_gather_fwd_kernel<4><<<grid, 128, 0, (cudaStream_t)stream>>>(t->idx, t->last, u, A, s, t->p_in, t->p_out,
                                                                         t->n_max, B, M, O);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

int vcm_project_gather_bwd(const vcm_tables *h, const float *ds, const float *u, const float *A, float *dA, float *du,
                           int B, int M, int O, void *stream) {
    const Tables *t = as_tables(h);
    VCM_CHECK(t != nullptr, "tables handle is NULL");
    VCM_CHECK(B > 0 && M > 0 && O > 0 && M * O <= 64 && M <= 32, "project-first gather needs M*O <= 64");
    VCM_DEVPTR(ds); VCM_DEVPTR(A); VCM_DEVPTR_OPT(u); VCM_DEVPTR_OPT(dA); VCM_DEVPTR_OPT(du);
    cudaStream_t st = (cudaStream_t)stream;
    if (dA != nullptr && t->p_out > 0) {
        VCM_CHECK(u != nullptr, "u is required for dA");
        project_gather_dA_kernel<<<(unsigned)ceil_div(t->p_out, 4), 128, 0, st>>>(t->idx, t->last, ds, u, dA, t->p_in,
                                                                                  t->p_out, t->n_max, B, M, O);
        VCM_LAUNCH_CHECK();
    }
    if (du != nullptr && t->p_in > 0) {
        dim3 grid((unsigned)ceil_div(t->p_in, 4), (unsigned)ceil_div(B, 4));
        project_gather_du_kernel<4><<<grid, 128, 0, st>>>(t->rev_ptr, t->rev_slot, ds, A, du, t->p_in, t->p_out, t->n_max,
                                                          B, M, O);
        VCM_LAUNCH_CHECK();
    }
    return VCM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// K1 on the tensor pipe (TF32 mode only). Per output vertex the contraction is a small GEMM
//   D[c, m] = sum_n G[c, n] * A_p[n, m],   c = (b, i) columns of the vertex row, G[c, n] = x[b, idx[p,n], i]
// with M = 17 and at most a few dozen neighbours: far below the 128-row tiles tcgen05 works on, and with a
// different B operand (A_p) for every vertex, so UMMA tiles cannot be shared across vertices. Warp-level
// mma.sync m16n8k8 (tf32 in, fp32 accumulate) fits exactly: 16 columns x 8 bases x 8 neighbours per instruction,
// fragments taken straight from L2 (x, 32-byte segments per neighbour row) and shared memory (A_p). It replaces
// ~70 FMA-pipe instructions per 16 columns and neighbour octet by 3 HMMA + 4 LDG + 6 LDS, needs ~40 registers
// (32 warps per SM instead of 16) and leaves the kernel bound by the z store stream. Operands are fp32 bit
// patterns (the tensor core reads the top 19 bits): same arithmetic class as the TF32 basis GEMMs.
namespace vcm {

constexpr int kMmaMS = 40;         // smem row stride of the coefficient tile: t*40 mod 32 = 0,8,16,24 -> conflict-free

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int MTILES>              // 8-wide basis tiles: 3 for M <= 24, 1 for M <= 8
__global__ void __launch_bounds__(128, 8)
gather_contract_fwd_mma_kernel(const int32_t *__restrict__ idx_sorted, const int2 *__restrict__ meta_sorted,
                               const float *__restrict__ x, const float *__restrict__ A, float *__restrict__ z,
                               int p_in, int p_out, int n_max, int I, int M, int B, int tiles_per_cta) {
    extern __shared__ __align__(16) float smem[];
    const int slot = blockIdx.x;
    const int2 meta = __ldg(meta_sorted + slot);
    const int p = meta.x, l = meta.y;
    const int n_pad = (l + 7) & ~7;
    float *A_s = smem;                                              // [n_pad][kMmaMS], zero padded
    int32_t *idx_s = reinterpret_cast<int32_t *>(smem + (size_t)((n_max + 7) & ~7) * kMmaMS);   // [n_pad], -1 = no row
    const int tid = threadIdx.x;
    for (int e = tid; e < n_pad * kMmaMS; e += 128) {
        const int n = e / kMmaMS, m = e - n * kMmaMS;
        A_s[e] = (n < l && m < M) ? __ldg(A + ((size_t)p * n_max + n) * M + m) : 0.f;
    }
    for (int n = tid; n < n_pad; n += 128) {
        const int q = n < l ? __ldg(idx_sorted + (size_t)slot * n_max + n) : p_in;
        idx_s[n] = q == p_in ? -1 : q;
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int tiles_per_b = I >> 4, n_tiles = B * tiles_per_b;
    const int tile_end = min(n_tiles, (int)(blockIdx.y + 1) * tiles_per_cta);
    float *zp = z + (size_t)p * M * I;
    for (int ct = blockIdx.y * tiles_per_cta + warp; ct < tile_end; ct += 4) {
        const int b = ct / tiles_per_b, i0 = (ct - b * tiles_per_b) << 4;
        const float *xb = x + (size_t)b * p_in * I + i0 + g;
        float acc[MTILES][4];
#pragma unroll
        for (int mt = 0; mt < MTILES; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
#pragma unroll 2
        for (int ks = 0; ks < n_pad; ks += 8) {
            const int q_lo = idx_s[ks + t], q_hi = idx_s[ks + t + 4];
            uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            if (q_lo >= 0) {
                const float *r = xb + (size_t)q_lo * I;
                a0 = __float_as_uint(__ldg(r));
                a1 = __float_as_uint(__ldg(r + 8));
            }
            if (q_hi >= 0) {
                const float *r = xb + (size_t)q_hi * I;
                a2 = __float_as_uint(__ldg(r));
                a3 = __float_as_uint(__ldg(r + 8));
            }
            const float *Alo = A_s + (ks + t) * kMmaMS + g, *Ahi = Alo + 4 * kMmaMS;
#pragma unroll
            for (int mt = 0; mt < MTILES; ++mt)
                mma_tf32_16x8x8(acc[mt], a0, a1, a2, a3, __float_as_uint(Alo[mt * 8]), __float_as_uint(Ahi[mt * 8]));
        }
        float *zr = zp + (size_t)b * p_out * M * I + i0 + g;
#pragma unroll
        for (int mt = 0; mt < MTILES; ++mt) {
            const int m0 = mt * 8 + 2 * t;
            if (m0 < M) {
                zr[(size_t)m0 * I] = acc[mt][0];
                zr[(size_t)m0 * I + 8] = acc[mt][2];
            }
            if (m0 + 1 < M) {
                zr[(size_t)(m0 + 1) * I] = acc[mt][1];
                zr[(size_t)(m0 + 1) * I + 8] = acc[mt][3];
            }
        }
    }
}

}  // namespace vcm

extern "C" int vcm_gather_contract_fwd_ex(const vcm_tables *h, const float *x, const float *A, float *z, int B, int I,
                                          int M, int precision, void *stream) {
    using namespace vcm;
    const Tables *t = as_tables(h);
    {
        // TF32 mode default: shared-memory staged neighbour rows + mma.sync (csrc/gather_fwd_tile.cu)
        K1TilePlan tp;
        if (k1_tile_plan(t, B, I, M, precision, &tp)) {
            VCM_DEVPTR(x); VCM_DEVPTR(A); VCM_DEVPTR(z);
            if (int e0 = check_common(t, B, I, M)) return e0;
            return k1_tile_launch(t, tp, x, A, z, B, I, M, (cudaStream_t)stream);
        }
    }
    // opt-in: correct (tests/test_gpu_gemm_tc.py) but measured 1.5 % slower than the fp32 kernel in the full
    // training step (3.675 vs 3.618 ms, same box): K1 is bound by its memory system, not by the FMA pipe
    static const int use_mma = [] { const char *e = getenv("VCM_K1_MMA"); return e ? atoi(e) : 0; }();
    if (precision != VCM_PREC_TF32 || !use_mma || t == nullptr || M > 24 || M < 2 || (I & 15) != 0 || B <= 0 ||
        t->p_out == 0)
        return vcm_gather_contract_fwd(h, x, A, z, B, I, M, stream);
    VCM_DEVPTR(x); VCM_DEVPTR(A); VCM_DEVPTR(z);
    const int n_tiles = B * (I >> 4);
    // ~2 CTAs of 4 warps per SM slot set, each warp several column tiles
    int64_t want = ceil_div(8 * (int64_t)sm_count(), t->p_out);
    if (want < 1) want = 1;
    int64_t gy = want < ceil_div(n_tiles, 4) ? want : ceil_div(n_tiles, 4);
    const int tiles_per_cta = (int)(ceil_div(ceil_div(n_tiles, gy), 4) * 4);
    gy = ceil_div(n_tiles, tiles_per_cta);
    const size_t smem = ((size_t)((t->n_max + 7) & ~7) * (kMmaMS + 1)) * sizeof(float);
    dim3 grid((unsigned)t->p_out, (unsigned)gy);
    cudaStream_t st = (cudaStream_t)stream;
    if (M <= 8) {
        auto k = gather_contract_fwd_mma_kernel<1>;
        if (int e = set_smem(k, smem)) return e;
        k<<<grid, 128, smem, st>>>(t->idx_sorted, reinterpret_cast<const int2 *>(t->meta_sorted), x, A, z, t->p_in, t->p_out,
                                   t->n_max, I, M, B, tiles_per_cta);
    } else {
        auto k = gather_contract_fwd_mma_kernel<3>;
        if (int e = set_smem(k, smem)) return e;
        k<<<grid, 128, smem, st>>>(t->idx_sorted, reinterpret_cast<const int2 *>(t->meta_sorted), x, A, z, t->p_in, t->p_out,
                                   t->n_max, I, M, B, tiles_per_cta);
    }
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

// ---------------------------------------------------------------------------
// K5 on the tensor pipe. Per output vertex both gradients are small GEMMs over the vertex's columns c = (b, i):
//   dG[c, n] = sum_m dz[c, m] * A_p[n, m]          (M = columns, N = neighbours, K = bases)
//   dA[n, m] = sum_c G[n, c]  * dz[c, m]           (M = neighbours, N = bases, K = columns)
// With 17 bases and a few dozen neighbours per vertex (and a different A_p per vertex) these are far below
// tcgen05's 128-row tiles; warp-level mma.sync m16n8k8 (tf32 operands, fp32 accumulate) fits them exactly. The
// fp32 kernel above is issue-bound (~345 instructions per neighbour and 128 columns, 2/3 of them not FMAs: the
// 17-row lane reduction of dA and integer addressing); here dA accumulates in MMA fragments across ALL column
// tiles of a warp - no per-neighbour reduction at all - and dG is 3 HMMA per 8 neighbours.
// X3 = fp32-grade: every operand is split into hi (what the tensor core reads) and lo = x - trunc_tf32(x), and
// each product is hi*hi + lo*hi + hi*lo (coefficients are pre-split when staged in shared memory).
namespace vcm {

constexpr int kK5S = 36;           // smem row stride of the coefficient tiles: (g*36 + t) mod 32 is a permutation

__device__ __forceinline__ uint32_t tf32_lo(uint32_t a) {
    return __float_as_uint(__uint_as_float(a) - __uint_as_float(a & 0xFFFFE000u));
}

template <bool X3>
__device__ __forceinline__ void mma_x(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&al)[4], uint32_t b0,
                                      uint32_t b1, uint32_t bl0, uint32_t bl1) {
    mma_tf32_16x8x8(d, a[0], a[1], a[2], a[3], b0, b1);
    if (X3) {
        mma_tf32_16x8x8(d, al[0], al[1], al[2], al[3], b0, b1);
        mma_tf32_16x8x8(d, a[0], a[1], a[2], a[3], bl0, bl1);
    }
}

constexpr int kTC = 32;            // columns (b, i) per staged tile
constexpr int kTS = kTC;           // smem row = 32 words, no padding: the 16-byte chunks of a row are XOR-swizzled
// chunk swizzle of row r: eight consecutive rows get eight different values (ldmatrix of one chunk column is
// conflict-free), and rows r, r+1, r+2, r+3 land in four different chunk PAIRS (the scalar transposed dz loads of
// dG, lanes = 4 rows x 8 columns, are conflict-free as well)
__host__ __device__ constexpr int k5_swz(int r) { return ((r & 3) << 1) | ((r >> 2) & 1); }
constexpr int kDzRows = 24;        // bases padded to three 8-wide MMA tiles

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t &r0, uint32_t &r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int kNBuf = 3;           // staged tiles per warp: two in flight while one is multiplied

// One CTA = one output vertex (blockIdx.x, processing order) x one share of its columns (blockIdx.y); each of the
// four warps walks its own 32-column tiles. Per tile the warp stages dz[M][32] and the gathered x rows [n_max][32]
// in shared memory with cp.async (16-byte pieces, three buffers: two tiles are in flight while one is multiplied),
// then
//   dA[n][m] += X[n][c] * dz[m][c]^T   M = neighbours, N = bases, K = columns: both operands are k-contiguous in
//                                      the staged tiles, so every fragment is one ldmatrix (a tf32 = two b16);
//   dG[n][c]  = A_p[n][m] * dz[m][c]   M = neighbours, N = columns, K = bases: coefficient fragments by ldmatrix
//                                      (kept in registers for <= 32 neighbours), dz fragments by scalar LDS.
// dA accumulates in MMA fragments over all tiles of the warp: no per-neighbour lane reduction at all. Rows that do
// not exist (bases >= M, padding neighbours) are not staged: their ldmatrix / LDS addresses point at one zero row.
template <int NT16, bool X3, bool HAS_DA, bool HAS_DG>
__global__ void __launch_bounds__(128, NT16 <= 2 ? 4 : 3)
coeff_grad_mma_kernel(const int32_t *__restrict__ idx_sorted, const int2 *__restrict__ meta_sorted,
                      const float *__restrict__ dz, const float *__restrict__ x, const float *__restrict__ A,
                      float *__restrict__ dA_out, float *__restrict__ dG, int p_in, int p_out, int n_max, int I, int M,
                      int B, int tiles_per_cta, int slot0, int n_stage) {
    // slot0 / n_stage: this launch covers the slots from slot0 on (one degree bucket of the processing order) and
    // stages at most n_stage neighbour rows per tile (<= n_max, the row stride of the tables and of dA / dG)
    constexpr int NP = NT16 * 16;                                    // padded neighbour count
    constexpr bool KEEP = NT16 <= 2;                                 // coefficient fragments live in registers
    extern __shared__ __align__(16) float smem[];
    float *A_s = smem;                                               // [NP][kK5S]  coefficients (hi = as is)
    float *Al_s = A_s + NP * kK5S;                                   // [NP][kK5S]  lo parts (X3 only)
    int32_t *idx_s = reinterpret_cast<int32_t *>(Al_s + (X3 ? NP * kK5S : 0));   // [NP], -1 = no neighbour
    float *zero_s = reinterpret_cast<float *>(idx_s + NP);           // [kTS] zeros
    float *stage_s = zero_s + kTS;                                   // [4 warps][kNBuf][M + n_stage][kTS]; part_s at the end
    float *part_s = stage_s;                                         // [4 warps][NP][24]
    const int buf_words = (M + n_stage) * kTS, warp_words = kNBuf * buf_words;

    const int slot = blockIdx.x + slot0;
    const int2 meta = __ldg(meta_sorted + slot);
    const int p = meta.x, l = meta.y;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int tiles_per_b = I / kTC, n_tiles = B * tiles_per_b;
    const int tile_end = min(n_tiles, (int)(blockIdx.y + 1) * tiles_per_cta);
    float *my = stage_s + (size_t)warp * warp_words;
    const uint32_t my_u = (uint32_t)__cvta_generic_to_shared(my);
    const uint32_t zero_u = (uint32_t)__cvta_generic_to_shared(zero_s);
    const uint32_t buf_bytes = (uint32_t)buf_words * 4u, x_off = (uint32_t)(M * kTS) * 4u;

    // This lane's sixteen-byte pieces of a tile are the same for every tile: dz piece e = lane + 32k is row e/8,
    // columns 4*(e%8). The dz half of the first tiles is requested before anything else: it needs no table.
    constexpr int DZ_P = (24 * (kTC / 4) + 31) / 32;                // 6 rounds cover up to 24 bases
    int dz_src[DZ_P];
    uint32_t dz_dst[DZ_P];
#pragma unroll
    for (int k = 0; k < DZ_P; ++k) {
        const int e = lane + 32 * k, m = e >> 3, j = e & 7;
        dz_src[k] = m < M ? m * I + 4 * j : -1;
        dz_dst[k] = (uint32_t)(m * kTS + 4 * (j ^ k5_swz(m))) * 4u;
    }
    auto stage_dz = [&](int b, int i0, int buf) {
        const float *dzp = dz + (((size_t)b * p_out + p) * M) * I + i0;
        const uint32_t base = my_u + (uint32_t)buf * buf_bytes;
#pragma unroll
        for (int k = 0; k < DZ_P; ++k)
            if (dz_src[k] >= 0) cp_async_16(base + dz_dst[k], dzp + dz_src[k]);
    };
    int ct = blockIdx.y * tiles_per_cta + warp;
    int b = ct / tiles_per_b, it = ct - b * tiles_per_b;             // tile -> (batch entry, 32-column block)
    int nb = b, nit = it;                                            // cursor of the tile being prefetched
    auto advance = [&]() {
        nit += 4;
        while (nit >= tiles_per_b) { nit -= tiles_per_b; ++nb; }
    };
    if (ct < tile_end) stage_dz(b, it * kTC, 0);
    int b1 = 0, it1 = 0;
    if (ct + 4 < tile_end) { advance(); b1 = nb; it1 = nit; stage_dz(nb, nit * kTC, 1); }

    for (int e = tid; e < NP * kK5S; e += 128) {
        const int n = e / kK5S, m = e - n * kK5S;
        const float a = (n < l && m < M) ? __ldg(A + ((size_t)p * n_max + n) * M + m) : 0.f;
        A_s[e] = a;
        if (X3) Al_s[e] = __uint_as_float(tf32_lo(__float_as_uint(a)));
    }
    for (int n = tid; n < NP; n += 128) {
        const int q = n < l ? __ldg(idx_sorted + (size_t)slot * n_max + n) : p_in;
        idx_s[n] = q == p_in ? -1 : q;
    }
    if (tid < kTS) zero_s[tid] = 0.f;
    __syncthreads();

    // x pieces of this lane (rounds beyond the first two are looped), then the x half of the first tiles
    constexpr int X_P = 2;
    int x_src[X_P];
    uint32_t x_dst[X_P];
#pragma unroll
    for (int k = 0; k < X_P; ++k) {
        const int e = lane + 32 * k, n = e >> 3, j = e & 7;
        const int q = (HAS_DA && n < l) ? idx_s[n] : -1;
        x_src[k] = q >= 0 ? q * I + 4 * j : -1;
        x_dst[k] = x_off + (uint32_t)(n * kTS + 4 * (j ^ k5_swz(n))) * 4u;
    }
    auto stage_x = [&](int b, int i0, int buf) {
        if (!HAS_DA) return;
        const float *xp = x + (size_t)b * p_in * I + i0;
        const uint32_t base = my_u + (uint32_t)buf * buf_bytes;
#pragma unroll
        for (int k = 0; k < X_P; ++k)
            if (x_src[k] >= 0) cp_async_16(base + x_dst[k], xp + x_src[k]);
        for (int e = lane + 32 * X_P; e < l * (kTC / 4); e += 32) {
            const int n = e >> 3, j = e & 7;
            const int q = idx_s[n];
            if (q >= 0) cp_async_16(base + x_off + (uint32_t)(n * kTS + 4 * (j ^ k5_swz(n))) * 4u, xp + (size_t)q * I + 4 * j);
        }
    };
    if (ct < tile_end) stage_x(b, it * kTC, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (ct + 4 < tile_end) stage_x(b1, it1 * kTC, 1);
    asm volatile("cp.async.commit_group;" ::: "memory");

    // dG rows this lane stores (bit nt*2: row nt*16+g, bit nt*2+1: row nt*16+g+8): real neighbours only
    uint32_t st_mask = 0;
#pragma unroll
    for (int nt = 0; nt < NT16; ++nt) {
        const int n0 = nt * 16 + g, n1 = n0 + 8;
        if (n0 < l && idx_s[n0] >= 0) st_mask |= 1u << (2 * nt);
        if (n1 < l && idx_s[n1] >= 0) st_mask |= 2u << (2 * nt);
    }

    float acc[NT16][3][4];
#pragma unroll
    for (int a = 0; a < NT16; ++a)
#pragma unroll
        for (int b2 = 0; b2 < 3; ++b2) acc[a][b2][0] = acc[a][b2][1] = acc[a][b2][2] = acc[a][b2][3] = 0.f;

    // ldmatrix rows of this lane (byte offset of the row inside a buffer, or -1 = the zero row) and the row's chunk
    // swizzle. Operand A (16 x 8): matrices (rows 0-7 | 8-15) x (k 0-3 | 4-7) in the order a0..a3; operand B (8 x 8
    // as [n][k]): k 0-3, k 4-7. The chunk of k-block kc is 2*kc + (0|1), XORed with the row's swizzle.
    const int lr = lane & 7, lj = lane >> 3;
    int a_row[NT16], a_sw[NT16], b4_row, b2_row, b4_sw, b2_sw;
#pragma unroll
    for (int nt = 0; nt < NT16; ++nt) {
        const int n = nt * 16 + (lj & 1) * 8 + lr;
        a_row[nt] = (n < l && idx_s[n] >= 0) ? (int)x_off + n * kTS * 4 : -1;
        a_sw[nt] = k5_swz(n) ^ (lj >> 1);
    }
    {
        const int m4 = (lj >> 1) * 8 + lr, m2 = 16 + lr;
        b4_row = m4 < M ? m4 * kTS * 4 : -1;
        b2_row = m2 < M ? m2 * kTS * 4 : -1;
        b4_sw = k5_swz(m4) ^ (lj & 1);
        b2_sw = k5_swz(m2) ^ (lj & 1);
    }
    const uint32_t c_row_off = (uint32_t)(((lj & 1) * 8 + lr) * kK5S + (lj >> 1) * 4) * 4u;      // coefficient tile
    const uint32_t A_u = (uint32_t)__cvta_generic_to_shared(A_s), Al_u = (uint32_t)__cvta_generic_to_shared(Al_s);
    // dz words read by the scalar B-fragment loads of dG: basis ks*8 + t (+4), column c8*8 + g. In the swizzled
    // row that is word ((c8 ^ t) * 8) + (((g >> 2) ^ (r & 1)) * 4 + (g & 3)); -1 = basis >= M (zero)
    int zrow[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const int m = (r >> 1) * 8 + t + (r & 1) * 4;
        zrow[r] = m < M ? m * kTS + (((g >> 2) ^ (r & 1)) << 2) + (g & 3) : -1;
    }

    uint32_t cA[KEEP ? NT16 : 1][3][4], cAl[KEEP ? NT16 : 1][3][4];
    if (HAS_DG && KEEP) {
#pragma unroll
        for (int nt = 0; nt < NT16; ++nt)
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                ldsm_x4(cA[nt][ks], A_u + c_row_off + (uint32_t)(nt * 16 * kK5S + ks * 8) * 4u);
                if (X3) ldsm_x4(cAl[nt][ks], Al_u + c_row_off + (uint32_t)(nt * 16 * kK5S + ks * 8) * 4u);
            }
    }

    int buf = 0, pbuf = 2;                                           // buffer being multiplied / being filled
    for (; ct < tile_end; ct += 4) {
        if (ct + 8 < tile_end) {
            advance();
            stage_dz(nb, nit * kTC, pbuf);
            stage_x(nb, nit * kTC, pbuf);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();
        const uint32_t dz_u = my_u + (uint32_t)buf * buf_bytes;
        const float *dz_t = my + (size_t)buf * buf_words;
        const int i0 = it * kTC;

        if (HAS_DA) {
#pragma unroll
            for (int kc = 0; kc < kTC / 8; ++kc) {
                uint32_t zB[3][2], zBl[3][2];
                {
                    uint32_t r[4];
                    ldsm_x4(r, b4_row >= 0 ? dz_u + (uint32_t)b4_row + (uint32_t)(((2 * kc) ^ b4_sw) << 4) : zero_u);   // bases 0-7, 8-15
                    zB[0][0] = r[0]; zB[0][1] = r[1]; zB[1][0] = r[2]; zB[1][1] = r[3];
                    ldsm_x2(zB[2][0], zB[2][1], b2_row >= 0 ? dz_u + (uint32_t)b2_row + (uint32_t)(((2 * kc) ^ b2_sw) << 4) : zero_u);
                }
                if (X3) {
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt) { zBl[mt][0] = tf32_lo(zB[mt][0]); zBl[mt][1] = tf32_lo(zB[mt][1]); }
                }
#pragma unroll
                for (int nt = 0; nt < NT16; ++nt) {
                    if (NT16 > 1 && nt * 16 >= l) break;                       // warp-uniform
                    uint32_t xa[4], xal[4] = {0u, 0u, 0u, 0u};
                    ldsm_x4(xa, a_row[nt] >= 0 ? dz_u + (uint32_t)a_row[nt] + (uint32_t)(((2 * kc) ^ a_sw[nt]) << 4) : zero_u);
                    if (X3) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) xal[r] = tf32_lo(xa[r]);
                    }
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt)
                        mma_x<X3>(acc[nt][mt], xa, xal, zB[mt][0], zB[mt][1], zBl[mt][0], zBl[mt][1]);
                }
            }
        }
        if (HAS_DG && KEEP) {
            float *dGp = dG + (((size_t)b * p_out + p) * n_max) * I + i0 + 2 * t;
#pragma unroll
            for (int c8 = 0; c8 < kTC / 8; ++c8) {
                // dz as the B operand: k = basis, n = column
                uint32_t zb[3][2], zbl[3][2];
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const uint32_t v = zrow[r] >= 0 ? __float_as_uint(dz_t[zrow[r] + ((c8 ^ t) << 3)]) : 0u;
                    zb[r >> 1][r & 1] = v;
                    if (X3) zbl[r >> 1][r & 1] = tf32_lo(v);
                }
#pragma unroll
                for (int nt = 0; nt < NT16; ++nt) {
                    if (NT16 > 1 && nt * 16 >= l) break;                       // warp-uniform
                    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks)
                        mma_x<X3>(d, cA[nt][ks], cAl[nt][ks], zb[ks][0], zb[ks][1], zbl[ks][0], zbl[ks][1]);
                    float *row0 = dGp + (size_t)(nt * 16 + g) * I + c8 * 8;
                    if (st_mask & (1u << (2 * nt))) *reinterpret_cast<float2 *>(row0) = make_float2(d[0], d[1]);
                    if (st_mask & (2u << (2 * nt))) *reinterpret_cast<float2 *>(row0 + (size_t)8 * I) = make_float2(d[2], d[3]);
                }
            }
        }
        if (HAS_DG && !KEEP) {
            // many neighbours: the dz fragments of the whole tile stay in registers, the coefficient fragments of one
            // 16-neighbour block are loaded once per tile
            float *dGp = dG + (((size_t)b * p_out + p) * n_max) * I + i0 + 2 * t;
            uint32_t zb[kTC / 8][3][2], zbl[X3 ? kTC / 8 : 1][3][2];
#pragma unroll
            for (int c8 = 0; c8 < kTC / 8; ++c8)
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const uint32_t v = zrow[r] >= 0 ? __float_as_uint(dz_t[zrow[r] + ((c8 ^ t) << 3)]) : 0u;
                    zb[c8][r >> 1][r & 1] = v;
                    if (X3) zbl[c8][r >> 1][r & 1] = tf32_lo(v);
                }
#pragma unroll
            for (int nt = 0; nt < NT16; ++nt) {
                if (nt * 16 >= l) break;                                       // warp-uniform
                uint32_t ca[3][4], cal[3][4];
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) {
                    ldsm_x4(ca[ks], A_u + c_row_off + (uint32_t)(nt * 16 * kK5S + ks * 8) * 4u);
                    if (X3) ldsm_x4(cal[ks], Al_u + c_row_off + (uint32_t)(nt * 16 * kK5S + ks * 8) * 4u);
                }
#pragma unroll
                for (int c8 = 0; c8 < kTC / 8; ++c8) {
                    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks)
                        mma_x<X3>(d, ca[ks], cal[ks], zb[c8][ks][0], zb[c8][ks][1], zbl[X3 ? c8 : 0][ks][0], zbl[X3 ? c8 : 0][ks][1]);
                    float *row0 = dGp + (size_t)(nt * 16 + g) * I + c8 * 8;
                    if (st_mask & (1u << (2 * nt))) *reinterpret_cast<float2 *>(row0) = make_float2(d[0], d[1]);
                    if (st_mask & (2u << (2 * nt))) *reinterpret_cast<float2 *>(row0 + (size_t)8 * I) = make_float2(d[2], d[3]);
                }
            }
        }
        __syncwarp();                                  // every lane is done with this buffer before it is refilled
        // the tile just multiplied came from the cursor two steps back: recompute (b, it) of the next one
        it += 4;
        while (it >= tiles_per_b) { it -= tiles_per_b; ++b; }
        pbuf = buf;
        buf = buf == kNBuf - 1 ? 0 : buf + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!HAS_DA) return;

    // fragments -> smem (the staging area is free now), then a fixed-order sum over the four warps
    __syncthreads();
    float *pw = part_s + (size_t)warp * NP * 24;
#pragma unroll
    for (int nt = 0; nt < NT16; ++nt)
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
            const int n = nt * 16 + g, m = mt * 8 + 2 * t;
            pw[n * 24 + m] = acc[nt][mt][0];
            pw[n * 24 + m + 1] = acc[nt][mt][1];
            pw[(n + 8) * 24 + m] = acc[nt][mt][2];
            pw[(n + 8) * 24 + m + 1] = acc[nt][mt][3];
        }
    __syncthreads();
    float *dst = dA_out + ((size_t)blockIdx.y * p_out + p) * n_max * M;
    for (int e = tid; e < n_max * M; e += 128) {
        const int n = e / M, m = e - n * M;
        float s = 0.f;
        if (n < l && idx_s[n] >= 0)
            s = (part_s[n * 24 + m] + part_s[(NP + n) * 24 + m]) + (part_s[(2 * NP + n) * 24 + m] + part_s[(3 * NP + n) * 24 + m]);
        dst[e] = s;
    }
}

struct MmaGradPlan { int nt16, gy, tiles_per_cta; size_t smem; };

static size_t grad_mma_smem(int nt16, int M, int n_stage) {
    const int NP = nt16 * 16;
    const size_t stage = (size_t)4 * kNBuf * (M + n_stage) * kTS, part = (size_t)4 * NP * 24;
    return ((size_t)2 * NP * kK5S + NP + kTS + (stage > part ? stage : part)) * sizeof(float);
}

static bool plan_grad_mma(const Tables *t, int B, int I, int M, MmaGradPlan *pl) {
    if (M > 24 || M < 2 || I % kTC != 0 || t->n_max > 64 || t->p_out == 0) return false;
    pl->nt16 = (t->n_max + 15) / 16;
    const int n_tiles = B * (I / kTC);
    int64_t want = ceil_div(8 * (int64_t)sm_count(), t->p_out);        // ~8 CTAs of 4 warps per SM overall
    if (want < 1) want = 1;
    const int64_t max_gy = ceil_div(n_tiles, 8);                       // at least two tiles per warp
    int64_t gy = want < max_gy ? want : max_gy;
    if (gy < 1) gy = 1;
    pl->tiles_per_cta = (int)(ceil_div(ceil_div(n_tiles, gy), 4) * 4);
    pl->gy = (int)ceil_div(n_tiles, pl->tiles_per_cta);
    pl->smem = grad_mma_smem(pl->nt16, M, t->n_max);
    return true;
}

// One launch = the slots [slot0, slot0 + n_slots) of the processing order (vertices sorted by neighbour count,
// descending), all of which have at most min(16 * NT16, n_stage) neighbours.
template <int NT16, bool X3>
static int launch_grad_mma(const Tables *t, const MmaGradPlan &pl, const float *dz, const float *x, const float *A,
                           float *target, float *dG, int B, int I, int M, cudaStream_t st, int slot0, int n_slots,
                           int n_stage) {
    if (n_slots <= 0) return VCM_OK;
    dim3 grid((unsigned)n_slots, (unsigned)pl.gy);
    const int2 *meta = reinterpret_cast<const int2 *>(t->meta_sorted);
    const size_t smem = grad_mma_smem(NT16, M, n_stage);
#define VCM_K5M(DA_, DG_)                                                                                           \
    do {                                                                                                            \
        auto k = coeff_grad_mma_kernel<NT16, X3, DA_, DG_>;                                                         \
        if (int e = set_smem(k, smem)) return e;                                                                    \
        k<<<grid, 128, smem, st>>>(t->idx_sorted, meta, dz, x, A, target, dG, t->p_in, t->p_out, t->n_max, I, M,   \
                                   B, pl.tiles_per_cta, slot0, n_stage);                                            \
    } while (0)
    if (target != nullptr && dG != nullptr) VCM_K5M(true, true);
    else if (target != nullptr) VCM_K5M(true, false);
    else VCM_K5M(false, true);
#undef VCM_K5M
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

}  // namespace vcm

extern "C" {

size_t vcm_coeff_grad_workspace_ex(const vcm_tables *h, int B, int I, int M, int precision) {
    using namespace vcm;
    size_t need = vcm_coeff_grad_workspace(h, B, I, M);
    const Tables *t = as_tables(h);
    MmaGradPlan pl;
    if (t && precision != VCM_PREC_FP32 && B > 0 && plan_grad_mma(t, B, I, M, &pl) && pl.gy > 1) {
        const size_t n = (size_t)pl.gy * t->p_out * t->n_max * M;
        need = n > need ? n : need;
    }
    K5TmaPlan tp;
    if (k5_tma_plan(t, B, I, M, precision, &tp)) {
        const size_t n = k5_tma_workspace(t, tp, M);
        need = n > need ? n : need;
    }
    return need;
}

// Tensor-pipe K5 (shared-memory staged tiles, ldmatrix fragments): in TF32 mode it is the default for layers with
// <= 32 neighbours per vertex, where it is 17-27 % faster than the fp32 kernel (C2 step: 3.49 -> 3.38 ms); with
// more neighbours (the pooling layers) the coefficient fragments no longer fit in registers and the fp32 kernel
// wins. VCM_K5_MMA: 0 = never, 1 = wherever it applies (also in tf32x3 mode), 2 = the default rule.
static bool use_grad_mma(const vcm::Tables *t, int B, int I, int M, int precision, vcm::MmaGradPlan *pl) {
    using namespace vcm;
    static const int use_mma = [] { const char *e = getenv("VCM_K5_MMA"); return e ? atoi(e) : 2; }();
    if (!use_mma || precision == VCM_PREC_FP32 || t == nullptr || B <= 0 || !plan_grad_mma(t, B, I, M, pl)) return false;
    return !(use_mma >= 2 && (pl->nt16 > 2 || precision != VCM_PREC_TF32));
}

int vcm_coeff_grad_variant_ex(const vcm_tables *h, int B, int I, int M, int precision, int want_dA, int want_dG) {
    if (vcm::k5_small_ok(vcm::as_tables(h), B, I, M, precision, want_dA != 0, want_dG != 0)) return 3;
    return vcm_coeff_grad_variant(h, B, I, M, precision);
}

int vcm_coeff_grad_variant(const vcm_tables *h, int B, int I, int M, int precision) {
    vcm::MmaGradPlan pl;
    vcm::K5TmaPlan tp;
    if (vcm::k5_tma_plan(vcm::as_tables(h), B, I, M, precision, &tp)) return 2;
    return use_grad_mma(vcm::as_tables(h), B, I, M, precision, &pl) ? 1 : 0;
}

int vcm_coeff_grad_slabs(const vcm_tables *h, int B, int I, int M, int precision) {
    using namespace vcm;
    const Tables *t = as_tables(h);
    K5TmaPlan tp;
    if (t == nullptr || k5_small_ok(t, B, I, M, precision, true, true) || !k5_tma_plan(t, B, I, M, precision, &tp)) return 0;
    return tp.gy > 1 ? tp.gy : 0;
}

int vcm_coeff_grad_slabs_launch(const vcm_tables *h, const float *dz, const float *x, const float *A, float *slabs,
                                float *dG, int B, int I, int M, int precision, void *stream) {
    using namespace vcm;
    const Tables *t = as_tables(h);
    K5TmaPlan tp;
    VCM_CHECK(vcm_coeff_grad_slabs(h, B, I, M, precision) > 1 && k5_tma_plan(t, B, I, M, precision, &tp),
              "this shape leaves no partial slabs (query vcm_coeff_grad_slabs)");
    VCM_DEVPTR(dz); VCM_DEVPTR(x); VCM_DEVPTR(A); VCM_DEVPTR(slabs); VCM_DEVPTR_OPT(dG);
    if (int e0 = check_common(t, B, I, M)) return e0;
    // the first slab pointer doubles as the (unused) final destination
    return k5_tma_launch(t, tp, dz, x, A, slabs, dG, slabs, B, I, M, (cudaStream_t)stream, true);
}

int vcm_sum_slabs(const float *slabs, float *out, size_t n, int n_slabs, void *stream) {
    VCM_CHECK(n_slabs >= 1, "bad slab count %d", n_slabs);
    VCM_DEVPTR(slabs); VCM_DEVPTR(out);
    if (n == 0) return VCM_OK;
    vcm::launch_sum_splits(slabs, out, n, n_slabs, (cudaStream_t)stream);        // counts its launch
    const cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return vcm::fail((int)e, "kernel launch failed: %s", cudaGetErrorString(e));
    return VCM_OK;
}

int vcm_coeff_grad_ex(const vcm_tables *h, const float *dz, const float *x, const float *A, float *dA, float *dG,
                      float *ws, int B, int I, int M, int precision, void *stream) {
    using namespace vcm;
    const Tables *t = as_tables(h);
    MmaGradPlan pl;
    K5TmaPlan tp;
    if (k5_small_ok(t, B, I, M, precision, dA != nullptr, dG != nullptr)) {
        // a handful of channels, dA only: one warp per vertex, fragments straight from global memory
        VCM_DEVPTR(dz); VCM_DEVPTR(x); VCM_DEVPTR(dA);
        if (int e0 = check_common(t, B, I, M)) return e0;
        return k5_small_launch(t, dz, x, dA, B, I, M, precision, (cudaStream_t)stream);
    }
    if (k5_tma_plan(t, B, I, M, precision, &tp)) {
        VCM_DEVPTR(dz); VCM_DEVPTR(x); VCM_DEVPTR(A); VCM_DEVPTR_OPT(dA); VCM_DEVPTR_OPT(dG); VCM_DEVPTR_OPT(ws);
        VCM_CHECK(dA != nullptr || dG != nullptr, "nothing to compute: dA and dG are both NULL");
        if (int e0 = check_common(t, B, I, M)) return e0;
        return k5_tma_launch(t, tp, dz, x, A, dA, dG, ws, B, I, M, (cudaStream_t)stream);
    }
    if (!use_grad_mma(t, B, I, M, precision, &pl)) return vcm_coeff_grad(h, dz, x, A, dA, dG, ws, B, I, M, stream);
    VCM_DEVPTR(dz); VCM_DEVPTR(x); VCM_DEVPTR(A); VCM_DEVPTR_OPT(dA); VCM_DEVPTR_OPT(dG); VCM_DEVPTR_OPT(ws);
    VCM_CHECK(dA != nullptr || dG != nullptr, "nothing to compute: dA and dG are both NULL");
    if (dA != nullptr && pl.gy > 1 && ws == nullptr) return fail(VCM_EINVAL, "workspace required (%d splits)", pl.gy);
    float *target = dA == nullptr ? nullptr : (pl.gy > 1 ? ws : dA);
    cudaStream_t st = (cudaStream_t)stream;
    const bool x3 = precision == VCM_PREC_TF32X3;
    int e = VCM_OK;
    // Degree buckets: the processing order lists the vertices by neighbour count, descending, so the vertices with
    // more than 16 neighbours are the first n_gt16 slots. With 17..32 table columns they get the 32-neighbour kernel;
    // everything behind them runs the 16-neighbour kernel with a 16-row staging area (4 instead of 3 CTAs per SM,
    // half the fragment registers). Both launches write disjoint rows of dA / dG.
#define VCM_K5L(NT_, S0_, NS_, ST_)                                                                          \
    (x3 ? launch_grad_mma<NT_, true>(t, pl, dz, x, A, target, dG, B, I, M, st, S0_, NS_, ST_)                \
        : launch_grad_mma<NT_, false>(t, pl, dz, x, A, target, dG, B, I, M, st, S0_, NS_, ST_))
    static const int bucketed = [] { const char *e2 = getenv("VCM_K5_BUCKETS"); return e2 ? atoi(e2) : 1; }();
    switch (pl.nt16) {
        case 1: e = VCM_K5L(1, 0, t->p_out, t->n_max); break;
        case 2:
            if (bucketed && t->n_gt16 < t->p_out) {
                e = VCM_K5L(2, 0, t->n_gt16, t->n_max);
                if (!e) e = VCM_K5L(1, t->n_gt16, t->p_out - t->n_gt16, 16);
            } else {
                e = VCM_K5L(2, 0, t->p_out, t->n_max);
            }
            break;
        case 3: e = VCM_K5L(3, 0, t->p_out, t->n_max); break;
        default: e = VCM_K5L(4, 0, t->p_out, t->n_max); break;
    }
#undef VCM_K5L
    if (e) return e;
    if (dA != nullptr && pl.gy > 1) {
        const size_t n_dA = (size_t)t->p_out * t->n_max * M;
        sum_splits_kernel<<<(unsigned)ceil_div((int64_t)n_dA, 256), 256, 0, st>>>(ws, dA, n_dA, pl.gy);
        VCM_LAUNCH_CHECK();
    }
    return VCM_OK;
}

}  // extern "C"
