// K1 (stage 1 of the vc-conv forward, graphVAESSW.py:111-123) on the tensor pipe with shared-memory staged
// neighbour rows (TF32 mode):
//
//   z[b,p,m,i] = sum_n A[p,n,m] x[b,idx[p,n],i]
//
// One CTA = one output vertex (processing order, degree buckets) x a share of its UNITS; a unit is one batch entry x
// CW consecutive channels. Per unit the warp stages the vertex's neighbour rows x[idx[p,n]][CW] in shared memory with
// 16-byte cp.async pieces (a per-row gather of 128/256-byte segments: 128-bit coalesced loads, the index pattern has no
// box for TMA), S units deep (cp.async groups; the warp that waits on a stage is the warp that refills it), and
// contracts on mma.sync m16n8k8 (tf32 operands, fp32 accumulate):
//
//   z_unit[m][c] = AT[m][n] x[n][c]      M = bases (two 16-row tiles for 17 <= M <= 24), N = 8 columns, K = neighbours
//
// The coefficient fragments AT (per vertex, constant over its units) live in registers (<= 16 neighbours) or are
// re-read from the shared tile once per unit. The x fragments need x[n][c] with k = n, i.e. the transposed access:
// the MMA's n slots are PERMUTED (slot g of the j-th 8-column block is column c0(g) + j, c0 = 0,16,4,20,8,24,12,28), so a
// lane reads rows t, t+4 at four consecutive columns with one bank-conflict-free 128-bit LDS per row for FOUR MMA
// blocks; every address is one per-unit base plus immediates. Results leave as 128-bit stores that write whole sectors.
// Padding neighbours meet a zero coefficient and a zeroed row. Per 32-column unit of an 8-neighbour vertex: ~65 warp
// instructions for 2176 bytes of z (the fp32 kernel: ~320 per 1088 bytes).
#include "common.cuh"
#include "k1_tile.h"
#include "tc_common.cuh"

namespace vcm {
namespace {

constexpr int kW = 4;              // warps per CTA (independent: one vertex each)

struct K1Params {
    const int32_t *idx_sorted;
    const int2 *meta_sorted;
    const float *x, *A;
    float *z;
    int p_in, p_out, n_max, I, M, B;
    int units_per_cta, slot0, stages, cpi_shift;
};

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <int OFF>
__device__ __forceinline__ uint4 lds128_o(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ void stg128(float *p, float a, float b, float c, float d) {       // explicit global space
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint32_t comp(const uint4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
// 16-byte cp.async, skipped when off == 0xFFFFFFFF (predicated inside the asm: no branch around it)
template <int DOFF>
__device__ __forceinline__ void cp16_if(uint32_t dst, const char *base, uint32_t off) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a;\n\t"
        "setp.ne.u32 p, %2, 0xFFFFFFFF;\n\t"
        "cvt.u64.u32 a, %2;\n\t"
        "add.u64 a, a, %1;\n\t"
        "@p cp.async.cg.shared.global [%0+%3], [a], 16;\n\t}"
        ::"r"(dst), "l"(base), "r"(off), "n"(DOFF) : "memory");
}
__device__ __forceinline__ void cp_wait_all_but(int n) {            // n is warp-uniform
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    }
}
template <int K0, int K1, int STEP, int N>
__device__ __forceinline__ void x_rounds(uint32_t dst, const char *xb, const uint32_t (&xq)[N]) {
    if constexpr (K0 < K1) {
        cp16_if<K0 * STEP>(dst, xb, xq[K0]);
        x_rounds<K0 + 1, K1, STEP, N>(dst, xb, xq);
    }
}

// One WARP = one output vertex x a share of its units: the warp loads the vertex's coefficient fragments straight
// into registers, stages and multiplies its units, and never meets the other warps of the CTA (no __syncthreads: the
// prologue of one warp - two dependent global loads - hides behind the main loops of the others).
// NK = 8-neighbour k-steps the table width needs (<= 3: coefficient fragments live in registers), CW = columns per
// unit, MT = 16-row basis tiles (1: M <= 16, 2: M <= 24 -- rows 24..31 of the second tile do not exist).
template <int NK, int CW, int MT>
__global__ void __launch_bounds__(32 * kW, 4)
gather_fwd_tile_kernel(const K1Params P) {
    constexpr int NP = NK * 8;
    constexpr int RSB = (CW + 4) * 4;             // x row stride in bytes
    constexpr int PIECES = CW / 4, RPR = 32 / PIECES, XK = NP / RPR;
    constexpr int STAGE = NP * RSB;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int S = P.stages, M = P.M, I = P.I;
    const uint32_t my_ring = smem_u32(smem_raw) + (uint32_t)(warp * S) * STAGE;
    uint8_t *my_ring_p = smem_raw + (size_t)(warp * S) * STAGE;

    const int slot = blockIdx.x * kW + warp;
    if (slot >= P.p_out) return;
    const int2 meta = __ldg(P.meta_sorted + slot);
    const int p = meta.x, l = meta.y;
    const int total_units = P.B << P.cpi_shift, cpi_mask = (1 << P.cpi_shift) - 1;
    const int u_begin = blockIdx.y * P.units_per_cta;
    const int u_end = min(total_units, u_begin + P.units_per_cta);
    const int32_t *idx_row = P.idx_sorted + (size_t)slot * P.n_max;

    // x rows of a unit: round k of the warp covers rows k*RPR .. k*RPR+RPR-1, lane -> (row, 16-byte piece); byte
    // offsets q * row_bytes of this lane's rows in registers (0xFFFFFFFF = no such neighbour: the row is zeroed once)
    const int xr0 = lane / PIECES, xj = lane % PIECES;
    const uint32_t x_dst0 = (uint32_t)(xr0 * RSB + xj * 16);
    const uint32_t row_bytes = (uint32_t)I * 4u;                     // q * row_bytes < 2^32 (checked by the plan)
    uint32_t xq[XK];
#pragma unroll
    for (int k = 0; k < XK; ++k) {
        const int n = k * RPR + xr0;
        const int q = n < l ? __ldg(idx_row + n) : P.p_in;
        xq[k] = q != P.p_in ? (uint32_t)q * row_bytes : 0xFFFFFFFFu;
    }
    // coefficient fragments AT[m][n] of operand A: a0 = (m, t), a1 = (m+8, t), a2 = (m, t+4), a3 = (m+8, t+4) of
    // k-step ks (neighbours ks*8 ..), m = mt*16 + g; zero where basis or neighbour does not exist
    uint32_t cT[MT][NK][4];
    {
        const float *Ap = P.A + (size_t)p * P.n_max * M;
#pragma unroll
        for (int ks = 0; ks < NK; ++ks) {
            const int n0 = ks * 8 + t, n1 = n0 + 4;
            const bool v0 = n0 < l && __ldg(idx_row + min(n0, P.n_max - 1)) != P.p_in;
            const bool v1 = n1 < l && __ldg(idx_row + min(n1, P.n_max - 1)) != P.p_in;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int m0 = mt * 16 + g, m1 = m0 + 8;
                cT[mt][ks][0] = (v0 && m0 < M) ? __float_as_uint(__ldg(Ap + (size_t)n0 * M + m0)) : 0u;
                cT[mt][ks][1] = (v0 && m1 < M) ? __float_as_uint(__ldg(Ap + (size_t)n0 * M + m1)) : 0u;
                cT[mt][ks][2] = (v1 && m0 < M) ? __float_as_uint(__ldg(Ap + (size_t)n1 * M + m0)) : 0u;
                cT[mt][ks][3] = (v1 && m1 < M) ? __float_as_uint(__ldg(Ap + (size_t)n1 * M + m1)) : 0u;
            }
        }
    }
    // rows of the ring that no copy will ever write (padding neighbours) must be finite: zero them once
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int k = 0; k < XK; ++k)
            if (xq[k] == 0xFFFFFFFFu)
                *reinterpret_cast<float4 *>(my_ring_p + (size_t)s * STAGE + x_dst0 + k * RPR * RSB) = make_float4(0.f, 0.f, 0.f, 0.f);

    const int nrounds = (l + RPR - 1) / RPR;                          // warp-uniform
    auto issue = [&](int u, int s) {
        const int b = u >> P.cpi_shift, ch = u & cpi_mask;
        const char *xb = reinterpret_cast<const char *>(P.x + (size_t)b * P.p_in * I + ch * CW + xj * 4);
        asm volatile("" : "+l"(xb));                                 // one 64-bit base per unit, not one per piece
        const uint32_t dst = my_ring + (uint32_t)s * STAGE + x_dst0;
        if (XK <= 2 || nrounds > 2) {
            x_rounds<0, XK, RPR * RSB, XK>(dst, xb, xq);
        } else {
            x_rounds<0, (XK < 2 ? XK : 2), RPR * RSB, XK>(dst, xb, xq);
        }
    };
    int u = u_begin;                                                 // unit being multiplied
    int u_issue = u;                                                 // next unit to request
    __syncwarp();
    for (int s = 0; s < S - 1; ++s, ++u_issue) {
        if (u_issue < u_end) issue(u_issue, s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    const int nks = (l + 7) >> 3;                                    // k-steps this vertex needs (warp-uniform)
    // x fragments (operand B: k = neighbour, n = column). The order of the MMA's n slots is free: slot g of the j-th
    // 8-column block stands for column c0(g) + j of a 32-column block, c0 = 0,16,4,20,8,24,12,28 for g = 0..7. ONE
    // 128-bit load (row t + 4h, columns c0(g)..c0(g)+3) then feeds four MMA blocks, and a lane's results are columns
    // 4t..4t+3 (slot 2t) and 16+4t.. (slot 2t+1) of its basis rows: two 128-bit stores per row, each instruction writing
    // whole 32-byte sectors. (A quarter warp - g in {0,1}, t = 0..3 - reads 16-byte bank groups 9t and 9t + 4 mod 8.)
    const uint32_t b_lane = (uint32_t)(t * RSB + ((g & 1) ? 64 + (g >> 1) * 16 : (g >> 1) * 16));   // row t, columns c0(g)..
    const bool st1 = MT > 1 && 16 + g < M;                           // this lane stores a row of the second tile
    const bool st0b = g + 8 < M, st0a = g < M;

    int stage = 0, fill = S - 1;                                     // stage being multiplied / refilled next
    for (; u < u_end; ++u) {
        cp_wait_all_but(S - 2);                        // the oldest of the S - 1 groups in flight = this unit
        __syncwarp();
        if (u_issue < u_end) issue(u_issue, fill);
        asm volatile("cp.async.commit_group;" ::: "memory");
        ++u_issue;
        if (++fill == S) fill = 0;

        const uint32_t xb = my_ring + (uint32_t)stage * STAGE + b_lane;
        const int b = u >> P.cpi_shift, ch = u & cpi_mask;
        float *zp = P.z + (((size_t)b * P.p_out + p) * M + g) * I + ch * CW + 4 * t;
        float *zp8 = zp + (size_t)8 * I, *zp16 = zp + (size_t)16 * I;
        asm volatile("" : "+l"(zp), "+l"(zp8), "+l"(zp16));
#pragma unroll
        for (int blk = 0; blk < CW / 32; ++blk) {
            float d[4][MT][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) d[j][mt][0] = d[j][mt][1] = d[j][mt][2] = d[j][mt][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < NK; ++ks) {
                if (NK > 1 && ks >= nks) break;                      // warp-uniform
                uint4 b0, b1;
                switch (ks) {                                        // immediates: ks * 8 rows and the odd row
                    case 0: b0 = lds128_o<0 * 8 * RSB + 0>(xb + blk * 128); b1 = lds128_o<0 * 8 * RSB + 4 * RSB>(xb + blk * 128); break;
                    case 1: b0 = lds128_o<1 * 8 * RSB + 0>(xb + blk * 128); b1 = lds128_o<1 * 8 * RSB + 4 * RSB>(xb + blk * 128); break;
                    default: b0 = lds128_o<2 * 8 * RSB + 0>(xb + blk * 128); b1 = lds128_o<2 * 8 * RSB + 4 * RSB>(xb + blk * 128); break;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) mma_tf32(d[j][mt], cT[mt][ks], comp(b0, j), comp(b1, j));
            }
            // block j, fragment (c0, c1) = columns (4t + j, 16 + 4t + j) of basis row g; (c2, c3) the same of row g + 8
            if (st0a) {
                stg128(zp + blk * 32, d[0][0][0], d[1][0][0], d[2][0][0], d[3][0][0]);
                stg128(zp + blk * 32 + 16, d[0][0][1], d[1][0][1], d[2][0][1], d[3][0][1]);
            }
            if (st0b) {
                stg128(zp8 + blk * 32, d[0][0][2], d[1][0][2], d[2][0][2], d[3][0][2]);
                stg128(zp8 + blk * 32 + 16, d[0][0][3], d[1][0][3], d[2][0][3], d[3][0][3]);
            }
            if (MT > 1 && st1) {
                stg128(zp16 + blk * 32, d[0][MT - 1][0], d[1][MT - 1][0], d[2][MT - 1][0], d[3][MT - 1][0]);
                stg128(zp16 + blk * 32 + 16, d[0][MT - 1][1], d[1][MT - 1][1], d[2][MT - 1][1], d[3][MT - 1][1]);
            }
        }
        if (++stage == S) stage = 0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

size_t cta_smem(int nk, int cw, int stages) { return (size_t)kW * stages * nk * 8 * (cw + 4) * 4; }

template <int NK, int CW, int MT>
int launch_tile(const Tables *t, const K1TilePlan &pl, K1Params P, cudaStream_t st) {
    const size_t smem = cta_smem(NK, CW, pl.stages);
    auto k = gather_fwd_tile_kernel<NK, CW, MT>;
    static size_t configured = 0;
    if (smem > configured) {
        VCM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((unsigned)ceil_div(t->p_out, kW), (unsigned)pl.gy);
    k<<<grid, 32 * kW, smem, st>>>(P);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

template <int CW, int MT>
int launch_all(const Tables *t, const K1TilePlan &pl, K1Params P, cudaStream_t st) {
    switch (pl.nk) {
        case 1: return launch_tile<1, CW, MT>(t, pl, P, st);
        case 2: return launch_tile<2, CW, MT>(t, pl, P, st);
        default: return launch_tile<3, CW, MT>(t, pl, P, st);
    }
}

}  // namespace

bool k1_tile_plan(const Tables *t, int B, int I, int M, int precision, K1TilePlan *pl) {
    static const int enabled = tune_flag("VCM_K1_TILE", 1);
    if (!enabled || precision != VCM_PREC_TF32 || t == nullptr || B <= 0 || t->p_out == 0) return false;
    if (M < 2 || M > 24 || I % 32 != 0 || t->n_max > 64) return false;
    const int cpi32 = I / 32;
    if ((cpi32 & (cpi32 - 1)) != 0) return false;
    if ((int64_t)t->p_in * I * 4 >= (int64_t)1 << 32) return false;
    pl->nk = (t->n_max + 7) / 8;
    if (pl->nk > 3) return false;                                    // the pooling layers keep the fp32 kernel
    pl->cw = (I % 64 == 0) ? tune_flag("VCM_K1_CW", 32) : 32;
    if (pl->cw != 32 && pl->cw != 64) pl->cw = 32;
    pl->stages = tune_flag("VCM_K1_STAGES", 3);
    if (pl->stages < 2) pl->stages = 2;
    if (pl->stages > 8) pl->stages = 8;
    const int units = B * (I / pl->cw);
    // one warp per (vertex, share of its units): enough warps to fill the machine, at least 4 units each
    int64_t want = ceil_div((int64_t)tune_flag("VCM_K1_WARPS_PER_SM", 32) * sm_count(), t->p_out);
    const int64_t max_gy = ceil_div(units, 4);
    if (want > max_gy) want = max_gy;
    if (want < 1) want = 1;
    pl->units_per_cta = (int)ceil_div(units, want);
    pl->gy = (int)ceil_div(units, pl->units_per_cta);
    return true;
}

int k1_tile_launch(const Tables *t, const K1TilePlan &pl, const float *x, const float *A, float *z, int B, int I, int M,
                   cudaStream_t st) {
    if (!aligned16(x) || (reinterpret_cast<uintptr_t>(z) & 7) != 0) return fail(VCM_EINVAL, "x must be 16-byte, z 8-byte aligned");
    K1Params P;
    P.idx_sorted = t->idx_sorted;
    P.meta_sorted = reinterpret_cast<const int2 *>(t->meta_sorted);
    P.x = x; P.A = A; P.z = z;
    P.p_in = t->p_in; P.p_out = t->p_out; P.n_max = t->n_max; P.I = I; P.M = M; P.B = B;
    P.units_per_cta = pl.units_per_cta; P.slot0 = 0; P.stages = pl.stages;
    int sh = 0;
    while ((1 << sh) < I / pl.cw) ++sh;
    P.cpi_shift = sh;
    if (pl.cw == 64) return M <= 16 ? launch_all<64, 1>(t, pl, P, st) : launch_all<64, 2>(t, pl, P, st);
    return M <= 16 ? launch_all<32, 1>(t, pl, P, st) : launch_all<32, 2>(t, pl, P, st);
}

}  // namespace vcm
