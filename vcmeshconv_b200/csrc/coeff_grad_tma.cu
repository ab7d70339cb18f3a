// K5 with TMA-staged dz tiles and an mbarrier pipeline per warp (TF32 mode).
//
//   dA[p,n,m]   = sum_{b,i} dz[b,p,m,i] x[b,idx[p,n],i]        autograd of graphVAESSW.py:123 wrt the coefficients
//   dG[b,p,n,i] = sum_m A[p,n,m] dz[b,p,m,i]                   ... wrt the gathered rows (K6 scatters it into dx)
//
// One CTA = one output vertex (processing order, degree buckets) x a share of its UNITS; a unit is one batch entry x
// CW consecutive channels: the dz block [M][CW] of that vertex and its <= n_max gathered rows of x [n][CW]. Each warp
// owns a ring of `stages` unit buffers:
//   * the dz block arrives by ONE cp.async.bulk.tensor (TMA) per unit: a 3-D box (32 columns, M bases, CW/32 column
//     blocks) of the tensor (32, B*P_out*M rows, I/32) laid over dz, 128-byte swizzled in shared memory, completing
//     on the stage's mbarrier (expect_tx);
//   * the gathered x rows arrive as 16-byte cp.async pieces (a per-row gather; every lane's pieces complete on the
//     same mbarrier through cp.async.mbarrier.arrive.noinc), rows padded by 16 bytes so ldmatrix is conflict-free.
// The warp that waits on a stage is the warp that refills it (no cross-warp traffic in the main loop).
//   dA[n][m] += X[n][c] dz[m][c]^T  mma.sync m16n8k8 tf32: M = neighbours, N = bases, K = columns; both operands are
//                                   k-contiguous, every fragment is one ldmatrix; accumulates in fragments over all
//                                   units of the warp, fixed-order sum over warps (and CTAs) at the end: no atomics;
//   dG[n][c]  = A_p[n][m] dz[m][c]  M = neighbours, N = columns, K = bases. dz is needed with k = basis, the transposed
//                                   access; the MMA's n slots are PERMUTED (slot g of the j-th 8-column block is
//                                   column c0(g) + j), so one 128-bit LDS per basis row feeds four MMA blocks, is
//                                   bank-conflict-free under TMA's 128-byte swizzle, and the results leave as 128-bit
//                                   stores that write whole sectors.
// Rows that do not exist (bases >= M, padding neighbours) alias rows that do (see below): no predicates in the loop.
// Against the cp.async-only kernel of gather.cu (per 32-column tile ~375 warp instructions, 170 of them staging) this
// one spends ~25 on staging per 64-column unit.
#include "k5_tma.h"
#include "tc_common.cuh"

namespace vcm {
namespace {

constexpr int kAS = 36;            // coefficient tile row stride (words): (g*36 + t) mod 32 is a permutation
constexpr int kMaxWarps = 4;

struct K5Params {
    const int32_t *idx_sorted;
    const int2 *meta_sorted;
    const float *x, *A;
    float *dA_out, *dG;
    int p_in, p_out, n_max, I, M, B;
    int units_per_cta, slot0, stages, cpi_shift;
    uint32_t stage_bytes, x_off, zero_off;
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
__device__ __forceinline__ void ldsm2(uint32_t &r0, uint32_t &r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
template <int OFF>
__device__ __forceinline__ void ldsm4_o(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4+%5];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr), "n"(OFF));
}
template <int OFF>
__device__ __forceinline__ void ldsm2_o(uint32_t &r0, uint32_t &r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2+%3];" : "=r"(r0), "=r"(r1) : "r"(addr), "n"(OFF));
}
template <int OFF>
__device__ __forceinline__ uint4 lds128_o(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ uint32_t comp(const uint4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
template <int OFF>
__device__ __forceinline__ uint32_t lds32_o(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
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
__device__ __forceinline__ void cp_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    const long long t0 = clock64();
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000LL) __trap();       // a protocol bug aborts the launch, never hangs
    } while (!ok);
}

template <int K0, int K1, int STEP, int N>
__device__ __forceinline__ void x_rounds_impl(uint32_t dst, const char *xb, const uint32_t (&xq)[N]) {
    if constexpr (K0 < K1) {
        cp16_if<K0 * STEP>(dst, xb, xq[K0]);
        x_rounds_impl<K0 + 1, K1, STEP, N>(dst, xb, xq);
    }
}

template <int NT16, int CW, bool HAS_DA, bool HAS_DG>
__global__ void __launch_bounds__(32 * kMaxWarps, NT16 <= 1 ? 4 : (NT16 == 2 ? 3 : 2))
coeff_grad_tma_kernel(const __grid_constant__ CUtensorMap map_dz, const K5Params P) {
    constexpr int NP = NT16 * 16;                 // padded neighbour count
    constexpr int NBLK = CW / 32;                 // 32-column TMA blocks per unit
    constexpr int RSB = (CW + 4) * 4;             // x row stride in bytes (16-byte pad: conflict-free ldmatrix)
    constexpr int PIECES = CW / 4;                // 16-byte pieces per x row
    constexpr int RPR = 32 / PIECES;              // x rows per cp.async round of the warp
    constexpr bool KEEP = NT16 <= 2;              // coefficient fragments of dG live in registers
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u = smem_u32(smem_raw);
    const uint32_t base_u = (raw_u + 1023u) & ~1023u;            // the 128-byte swizzle is a function of address bits 7-9
    uint8_t *base = smem_raw + (base_u - raw_u);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3, lr = lane & 7, lj = lane >> 3;
    const int S = P.stages, M = P.M, I = P.I;
    const uint32_t stage_bytes = P.stage_bytes;
    const uint32_t ring_bytes = (uint32_t)(W * S) * stage_bytes;
    float *A_s = reinterpret_cast<float *>(base + ring_bytes);                   // [NP][kAS], k slots permuted
    int32_t *idx_s = reinterpret_cast<int32_t *>(A_s + NP * kAS);               // [NP], -1 = no neighbour
    uint64_t *bars = reinterpret_cast<uint64_t *>(idx_s + NP);                   // [W][S]
    float *raw_s = reinterpret_cast<float *>(bars + W * S);                      // [n_max][M] coefficients as stored
    const uint32_t my_ring = base_u + (uint32_t)(warp * S) * stage_bytes;
    const uint32_t my_bar = smem_u32(bars + warp * S);

    const int slot = blockIdx.x + P.slot0;
    const int2 meta = __ldg(P.meta_sorted + slot);
    const int p = meta.x, l = meta.y;
    const int total_units = P.B << P.cpi_shift, cpi_mask = (1 << P.cpi_shift) - 1;
    const int u_begin = blockIdx.y * P.units_per_cta;
    const int u_end = min(total_units, u_begin + P.units_per_cta);

    // this warp's barriers; the dz tiles of its first units are requested before anything else
    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(bars + warp * S + s, HAS_DA ? 33 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dz) : "memory");
    }
    // the (24 - M) rows behind a dz tile are read as bases >= M of its last column block: keep them finite
    for (int s = 0; s < S; ++s)
        for (int e = lane; e < (24 - M) * 8; e += 32)
            *reinterpret_cast<float4 *>(base + (size_t)(warp * S + s) * stage_bytes + P.x_off + e * 16) =
                make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    auto issue_dz = [&](int u, int s) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        if (leader) {
            const uint32_t bar = my_bar + 8u * s;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(M * CW * 4))
                         : "memory");
            const int b = u >> P.cpi_shift, ch = u & cpi_mask;
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(my_ring + (uint32_t)s * stage_bytes), "l"(&map_dz), "r"(bar), "r"(0), "r"((b * P.p_out + p) * M),
                "r"(ch * NBLK)
                : "memory");
        }
    };
    int u = u_begin + warp;                                          // unit being multiplied
    int u_issue = u;                                                 // next unit to request
    int n_pro = 0;
    for (; n_pro < S && u_issue < u_end; ++n_pro, u_issue += W) issue_dz(u_issue, n_pro);

    // neighbour ids and the vertex's coefficient rows as stored (one coalesced round of independent loads)
    for (int n = tid; n < NP; n += blockDim.x) {
        const int q = n < l ? __ldg(P.idx_sorted + (size_t)slot * P.n_max + n) : P.p_in;
        idx_s[n] = q == P.p_in ? -1 : q;
    }
    {
        const float *Ap = P.A + (size_t)p * P.n_max * M;
        for (int e = tid; e < l * M; e += blockDim.x) raw_s[e] = __ldg(Ap + e);
    }
    __syncthreads();

    // x rows of a unit: round k of the warp covers rows k*RPR .. k*RPR+RPR-1, lane -> (row, 16-byte piece); the byte
    // offsets q * row_bytes of this lane's rows live in registers (-1 = no such neighbour)
    constexpr int XR = NP / RPR;                                     // rounds that cover all staged rows
    constexpr int XK = XR < 8 ? XR : 8;                              // rounds kept in registers
    const int xr0 = lane / PIECES, xj = lane % PIECES;
    const uint32_t x_dst0 = P.x_off + (uint32_t)(xr0 * RSB + xj * 16);
    const uint32_t row_bytes = (uint32_t)I * 4u;                     // q * row_bytes < 2^32 (checked by the plan)
    uint32_t xq[XK];
#pragma unroll
    for (int k = 0; k < XK; ++k) {
        const int q = idx_s[k * RPR + xr0];
        xq[k] = q >= 0 ? (uint32_t)q * row_bytes : 0xFFFFFFFFu;
    }
    auto issue_x = [&](int u, int s) {
        if (!HAS_DA) return;
        const int b = u >> P.cpi_shift, ch = u & cpi_mask;
        const char *xb = reinterpret_cast<const char *>(P.x + (size_t)b * P.p_in * I + ch * CW + xj * 4);
        asm volatile("" : "+l"(xb));                                 // one 64-bit base per unit, not one per piece
        const uint32_t dst = my_ring + (uint32_t)s * stage_bytes + x_dst0;
        if (l > 4 * RPR) {                                            // warp-uniform: rounds 0-3 cover 4*RPR rows
            x_rounds_impl<0, XK, RPR * RSB, XK>(dst, xb, xq);
        } else {
            x_rounds_impl<0, (XK < 4 ? XK : 4), RPR * RSB, XK>(dst, xb, xq);
        }
        for (int k = XK; k * RPR < l; ++k) {                         // more than 8 rounds: the pooling layers
            const int q = idx_s[k * RPR + xr0];
            if (q >= 0) cp16(dst + (uint32_t)(k * RPR * RSB), xb + (uint32_t)q * row_bytes);
        }
        cp_arrive_noinc(my_bar + 8u * s);
    };
    {
        int uu = u;
        for (int s = 0; s < n_pro; ++s, uu += W) issue_x(uu, s);
    }

    // coefficient tile [neighbour][basis]; rows and bases that do not exist are zero
    for (int e = tid; e < NP * 24; e += blockDim.x) {
        const int n = e / 24, m = e - n * 24;
        A_s[n * kAS + m] = (n < l && m < M && idx_s[n] >= 0) ? raw_s[n * M + m] : 0.f;
    }
    __syncthreads();

    // ldmatrix rows of this lane (stage-relative byte offsets). Rows that do not exist alias rows that do: a padding
    // neighbour row only feeds the dA rows of that neighbour and a basis >= M only the dA columns of that basis, which
    // are never written out; in dG such a basis meets a zero coefficient (its dz alias is finite data of this tile).
    // x rows as operand A of dA (16 x 8): matrices (rows 0-7 | 8-15) x (k 0-3 | 4-7)
    uint32_t offA[NT16];
#pragma unroll
    for (int nt = 0; nt < NT16; ++nt) {
        const int n = nt * 16 + (lj & 1) * 8 + lr;
        offA[nt] = (n < l && idx_s[n] >= 0) ? P.x_off + (uint32_t)(n * RSB + (lj >> 1) * 16) : (uint32_t)((lj >> 1) * 16);
    }
    // dz rows as operand B of dA (n = basis, k = column): row r = blk*M + m of the swizzled tile, chunk ^= r & 7
    // (the x2 load of bases 16-23 is the x4 row + 16 rows: same swizzle, immediate offset 2048)
    uint32_t offB[NBLK];
    {
        const int m4 = (lj >> 1) * 8 + lr;
#pragma unroll
        for (int blk = 0; blk < NBLK; ++blk) {
            const int r4 = blk * M + m4;
            offB[blk] = (uint32_t)(r4 * 128 + ((((lj & 1) ^ r4) & 7) << 4));
        }
    }
    // dz fragments of dG (operand B: k = basis, n = column). The order of the MMA's n slots is free: slot g of the j-th
    // 8-column block stands for column c0(g) + j of a 32-column block, c0 = 0,16,4,20,8,24,12,28 for g = 0..7. ONE
    // 128-bit load (row = basis ks*8 + t + 4h, 16-byte chunk c0(g)/4) then feeds four MMA blocks, and a lane's results
    // are columns 4t..4t+3 (slot 2t) and 16+4t..16+4t+3 (slot 2t+1) of its two neighbour rows: two 128-bit stores per
    // row, each instruction writing whole 32-byte sectors. ks adds 8 rows = 1024 bytes and leaves the swizzle alone
    // (immediate offsets); a quarter warp (g in {0,1}, t = 0..3) touches the 8 different chunks {s_t, s_t ^ 4} of its
    // four rows (s_t = (r0 + t) & 7). Rows past the tile (bases >= M of the last block) fall into the head of the x
    // area, which is kept finite (zeroed above).
    uint32_t zoff[2][NBLK];
    {
        const int chunk = (g & 1) ? 4 + (g >> 1) : (g >> 1);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk) {
                const int r = blk * M + t + 4 * h;
                zoff[h][blk] = (uint32_t)(r * 128 + (((chunk ^ r) & 7) << 4));
            }
    }
    // coefficient fragments (operand A of dG) and the dG rows this lane stores
    const uint32_t A_u = smem_u32(A_s);
    const uint32_t c_row_off = (uint32_t)(((lj & 1) * 8 + lr) * kAS + (lj >> 1) * 4) * 4u;
    uint32_t cA[KEEP ? NT16 : 1][3][4];
    if (HAS_DG && KEEP) {
#pragma unroll
        for (int nt = 0; nt < NT16; ++nt)
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) ldsm4(cA[nt][ks], A_u + c_row_off + (uint32_t)(nt * 16 * kAS + ks * 8) * 4u);
    }
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

    int stage = 0;
    uint32_t phase = 0;
    for (; u < u_end; u += W) {
        mbar_wait_u32(my_bar + 8u * stage, phase);
        const uint32_t sb = my_ring + (uint32_t)stage * stage_bytes;
        // absolute addresses of this stage (sb is 1024-aligned: the XORs below only touch bits 5-6)
        uint32_t aB[NBLK], aZ[2][NBLK], aA[NT16];
#pragma unroll
        for (int blk = 0; blk < NBLK; ++blk) { aB[blk] = sb + offB[blk]; aZ[0][blk] = sb + zoff[0][blk]; aZ[1][blk] = sb + zoff[1][blk]; }
#pragma unroll
        for (int nt = 0; nt < NT16; ++nt) aA[nt] = sb + offA[nt];

        if (HAS_DA) {
#pragma unroll
            for (int kc = 0; kc < CW / 8; ++kc) {
                const int blk = kc >> 2;
                const uint32_t kx = (uint32_t)(kc & 3) << 5;
                uint32_t zB[3][2], r[4];
                const uint32_t ab = aB[blk] ^ kx;
                ldsm4_o<0>(r, ab);                                   // bases 0-7, 8-15
                zB[0][0] = r[0]; zB[0][1] = r[1]; zB[1][0] = r[2]; zB[1][1] = r[3];
                ldsm2_o<2048>(zB[2][0], zB[2][1], ab);               // bases 16-23
#pragma unroll
                for (int nt = 0; nt < NT16; ++nt) {
                    if (NT16 > 1 && nt * 16 >= l) break;             // warp-uniform
                    uint32_t xa[4];
                    ldsm4(xa, aA[nt] + (uint32_t)(kc * 32));
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt) mma_tf32(acc[nt][mt], xa, zB[mt][0], zB[mt][1]);
                }
            }
        }
        if (HAS_DG) {
            const int b = u >> P.cpi_shift, ch = u & cpi_mask;
            float *dGp = P.dG + (((size_t)b * P.p_out + p) * P.n_max + g) * I + ch * CW + 4 * t;
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk) {
                uint4 zb[3][2];                                      // [k-step][h]: basis ks*8 + t + 4h, columns c0(g)..c0(g)+3
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    zb[0][h] = lds128_o<0>(aZ[h][blk]);
                    zb[1][h] = lds128_o<1024>(aZ[h][blk]);
                    zb[2][h] = lds128_o<2048>(aZ[h][blk]);
                }
#pragma unroll
                for (int nt = 0; nt < NT16; ++nt) {
                    if (NT16 > 1 && nt * 16 >= l) break;             // warp-uniform
                    uint32_t ca[3][4];
                    if (!KEEP) {
#pragma unroll
                        for (int ks = 0; ks < 3; ++ks) ldsm4(ca[ks], A_u + c_row_off + (uint32_t)(nt * 16 * kAS + ks * 8) * 4u);
                    }
                    float d[4][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
#pragma unroll
                        for (int ks = 0; ks < 3; ++ks)
                            mma_tf32(d[j], KEEP ? cA[KEEP ? nt : 0][ks] : ca[ks], comp(zb[ks][0], j), comp(zb[ks][1], j));
                    }
                    // block j, fragment (c0, c1) = columns (4t + j, 16 + 4t + j) of row g; (c2, c3) the same of row g + 8
                    float *row0 = dGp + (size_t)(nt * 16) * I + blk * 32;
                    if (st_mask & (1u << (2 * nt))) {
                        *reinterpret_cast<float4 *>(row0) = make_float4(d[0][0], d[1][0], d[2][0], d[3][0]);
                        *reinterpret_cast<float4 *>(row0 + 16) = make_float4(d[0][1], d[1][1], d[2][1], d[3][1]);
                    }
                    if (st_mask & (2u << (2 * nt))) {
                        float *row8 = row0 + (size_t)8 * I;
                        *reinterpret_cast<float4 *>(row8) = make_float4(d[0][2], d[1][2], d[2][2], d[3][2]);
                        *reinterpret_cast<float4 *>(row8 + 16) = make_float4(d[0][3], d[1][3], d[2][3], d[3][3]);
                    }
                }
            }
        }
        __syncwarp();                                  // every lane is done with this stage before it is refilled
        if (u_issue < u_end) {
            issue_dz(u_issue, stage);
            issue_x(u_issue, stage);
        }
        u_issue += W;
        if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    if (!HAS_DA) return;

    // fragments -> smem (the rings are drained: every requested unit was waited for), fixed-order sum over the warps
    __syncthreads();
    float *part_s = reinterpret_cast<float *>(base);                 // [W][NP][24]
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
    float *dst = P.dA_out + ((size_t)blockIdx.y * P.p_out + p) * P.n_max * M;
    for (int e = tid; e < P.n_max * M; e += blockDim.x) {
        const int n = e / M, m = e - n * M;
        float s = 0.f;
        if (n < l && idx_s[n] >= 0)
            for (int w = 0; w < W; ++w) s += part_s[(w * NP + n) * 24 + m];
        dst[e] = s;
    }
}

struct StageGeom { uint32_t x_off, zero_off, stage_bytes; };

StageGeom stage_geom(int M, int cw, int np, bool has_da) {
    StageGeom g;
    const uint32_t dz_bytes = (uint32_t)(cw / 32) * M * 128;
    g.x_off = dz_bytes;
    const uint32_t x_bytes = has_da ? (uint32_t)np * (cw + 4) * 4 : 0;
    g.zero_off = 0;
    g.stage_bytes = (dz_bytes + x_bytes + 1023u) & ~1023u;
    return g;
}

size_t cta_smem(int M, int cw, int nt16, int warps, int stages) {
    const int np = nt16 * 16;
    const StageGeom g = stage_geom(M, cw, np, true);
    size_t ring = (size_t)warps * stages * g.stage_bytes;
    const size_t part = (size_t)warps * np * 24 * 4;
    if (part > ring) ring = part;
    return 1024 + ring + (size_t)np * kAS * 4 + (size_t)np * 4 + (size_t)warps * stages * 8 + (size_t)np * M * 4 + 16;
}

template <int NT16, int CW>
int launch_bucket(const CUtensorMap &map, const Tables *t, const K5TmaPlan &pl, K5Params P, int M, int n_slots,
                  cudaStream_t st) {
    if (n_slots <= 0) return VCM_OK;
    const bool has_da = P.dA_out != nullptr, has_dg = P.dG != nullptr;
    // the ring geometry of this bucket (NT16 may be below the plan's for the low-degree bucket)
    int stages = pl.stages;
    while (stages > 2 && cta_smem(M, CW, NT16, pl.warps, stages) > 112 * 1024) --stages;
    const StageGeom g = stage_geom(M, CW, NT16 * 16, true);
    P.stages = stages;
    P.stage_bytes = g.stage_bytes; P.x_off = g.x_off; P.zero_off = g.zero_off;
    const size_t smem = cta_smem(M, CW, NT16, pl.warps, stages);
    dim3 grid((unsigned)n_slots, (unsigned)pl.gy);
#define VCM_K5T(DA_, DG_)                                                                                    \
    do {                                                                                                     \
        auto k = coeff_grad_tma_kernel<NT16, CW, DA_, DG_>;                                                  \
        static size_t configured = 0;                                                                        \
        if (smem > configured) {                                                                             \
            VCM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
            configured = smem;                                                                               \
        }                                                                                                    \
        k<<<grid, 32 * pl.warps, smem, st>>>(map, P);                                                        \
    } while (0)
    if (has_da && has_dg) VCM_K5T(true, true);
    else if (has_da) VCM_K5T(true, false);
    else VCM_K5T(false, true);
#undef VCM_K5T
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

template <int CW>
int launch_cw(const CUtensorMap &map, const Tables *t, const K5TmaPlan &pl, const K5Params &P0, int M, cudaStream_t st) {
    K5Params P = P0;
    // Degree buckets: the processing order lists the vertices by neighbour count, descending; the first n_gt16 slots
    // have more than 16 neighbours and run the kernel sized for the table width, the rest the 16-neighbour kernel.
    const int n_hi = pl.nt16 > 1 ? t->n_gt16 : 0;
    int e = VCM_OK;
    P.slot0 = 0;
    switch (pl.nt16) {
        case 1: break;
        case 2: e = launch_bucket<2, CW>(map, t, pl, P, M, n_hi, st); break;
        case 3: e = launch_bucket<3, CW>(map, t, pl, P, M, n_hi, st); break;
        default: e = launch_bucket<4, CW>(map, t, pl, P, M, n_hi, st); break;
    }
    if (e) return e;
    P.slot0 = n_hi;
    return launch_bucket<1, CW>(map, t, pl, P, M, t->p_out - n_hi, st);
}

}  // namespace

bool k5_tma_plan(const Tables *t, int B, int I, int M, int precision, K5TmaPlan *pl) {
    static const int enabled = tune_flag("VCM_K5_TMA", 1);
    if (!enabled || precision != VCM_PREC_TF32 || t == nullptr || B <= 0 || t->p_out == 0) return false;
    if (M < 2 || M > 24 || I % 32 != 0 || t->n_max > 64) return false;
    const int cpi32 = I / 32;
    if ((cpi32 & (cpi32 - 1)) != 0) return false;                    // channel counts are powers of two times 32
    if ((int64_t)B * t->p_out * M >= (int64_t)1 << 31 || (int64_t)t->p_in * I * 4 >= (int64_t)1 << 32) return false;
    pl->nt16 = (t->n_max + 15) / 16;
    // measured on the C2 layer shapes: 32-column units with four 2-stage warps (five CTAs = 20 warps per SM) beat
    // 64-column units (fewer instructions per byte, but 8-12 warps per SM) on every shape but two, where they tie
    pl->cw = (I % 64 == 0) ? tune_flag("VCM_K5_CW", 32) : 32;
    if (pl->cw != 32 && pl->cw != 64) pl->cw = 32;
    if (I % pl->cw != 0) pl->cw = 32;
    pl->warps = tune_flag("VCM_K5_WARPS", 4);
    if (pl->warps < 1 || pl->warps > kMaxWarps) pl->warps = 4;
    pl->stages = tune_flag("VCM_K5_STAGES", 2);
    if (pl->stages < 2) pl->stages = 2;
    if (pl->stages > 8) pl->stages = 8;
    while (pl->stages > 2 && cta_smem(M, pl->cw, pl->nt16, pl->warps, pl->stages) > 112 * 1024) --pl->stages;
    if (cta_smem(M, pl->cw, pl->nt16, pl->warps, pl->stages) > 220 * 1024) return false;
    const int units = B * (I / pl->cw);
    // enough CTAs to fill the machine several times over, at least 4 units per warp
    int64_t want = ceil_div((int64_t)tune_flag("VCM_K5_CTAS_PER_SM", 8) * sm_count(), t->p_out);
    const int64_t max_gy = ceil_div(units, (int64_t)pl->warps * 4);
    if (want > max_gy) want = max_gy;
    if (want < 1) want = 1;
    pl->units_per_cta = (int)(ceil_div(ceil_div(units, want), pl->warps) * pl->warps);
    pl->gy = (int)ceil_div(units, pl->units_per_cta);
    return true;
}

size_t k5_tma_workspace(const Tables *t, const K5TmaPlan &pl, int M) {
    return pl.gy > 1 ? (size_t)pl.gy * t->p_out * t->n_max * M : 0;
}

int k5_tma_launch(const Tables *t, const K5TmaPlan &pl, const float *dz, const float *x, const float *A, float *dA,
                  float *dG, float *ws, int B, int I, int M, cudaStream_t st, bool defer_sum) {
    if (dA != nullptr && pl.gy > 1 && ws == nullptr) return fail(VCM_EINVAL, "workspace required (%d splits)", pl.gy);
    if (!aligned16(dz) || !aligned16(x) || (dG && !aligned16(dG))) return fail(VCM_EINVAL, "dz, x and dG must be 16-byte aligned");
    CUtensorMap map;
    const uint64_t dims[3] = {32, (uint64_t)B * t->p_out * M, (uint64_t)(I / 32)};
    const uint64_t strides[3] = {4, (uint64_t)I * 4, 128};
    const uint32_t box[3] = {32, (uint32_t)M, (uint32_t)(pl.cw / 32)};
    if (int e = make_map(&map, dz, 3, dims, strides, box)) return e;
    K5Params P;
    P.idx_sorted = t->idx_sorted;
    P.meta_sorted = reinterpret_cast<const int2 *>(t->meta_sorted);
    P.x = x; P.A = A;
    P.dA_out = dA == nullptr ? nullptr : (pl.gy > 1 ? ws : dA);
    P.dG = dG;
    P.p_in = t->p_in; P.p_out = t->p_out; P.n_max = t->n_max; P.I = I; P.M = M; P.B = B;
    P.units_per_cta = pl.units_per_cta; P.slot0 = 0; P.stages = pl.stages;
    int sh = 0;
    while ((1 << sh) < I / pl.cw) ++sh;
    P.cpi_shift = sh;
    P.stage_bytes = P.x_off = P.zero_off = 0;
    int e = pl.cw == 64 ? launch_cw<64>(map, t, pl, P, M, st) : launch_cw<32>(map, t, pl, P, M, st);
    if (e) return e;
    if (dA != nullptr && pl.gy > 1 && !defer_sum) launch_sum_splits(ws, dA, (size_t)t->p_out * t->n_max * M, pl.gy, st);
    return VCM_OK;
}

}  // namespace vcm
