// Basis GEMMs on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with
// fp32 accumulators in TMEM, operands staged by TMA (cp.async.bulk.tensor) into
// 128-byte-swizzled shared memory through an mbarrier ring.
//
//   FWD  out[r,o]   = epi( sum_{m,i} z[r,m,i] W[m,o,i] )        K2  graphVAESSW.py:125-138,161
//   DZ   dz[r,m,i]  = sum_o ds[r,o] Wt[(m,i),o]                  K3
//   DW   dW[(m,i),o]= sum_r z[r,(m,i)] ds[r,o]                   K4  (split over r, fixed-order reduce)
//
// One CTA computes one 128 x BN accumulator tile (UMMA M = 128, N = BN <= 256,
// K = 8 per instruction, 4 instructions per 32-float k-block):
//   warp 0    TMA producer (one lane)      warp 1    tcgen05.mma issuer (one lane)
//   warp 2    TMEM allocator               warps 4-7 epilogue: tcgen05.ld -> registers
//                                                    -> bias / ELU / blend -> global
// Operand layouts
//   FWD: A = z tile [128 r][32 k] K-major; B = W tile [BN o][32 i] K-major, cut out
//        of the reference's native [M][O][I] parameter by a 3-D tensor map
//        (i, o, m) - the stacked-basis matrix is never materialised for K2;
//   DZ : A = ds tile [128 r][32 o] K-major; B = Wt tile [BN k][32 o] K-major from the
//        transposed bases Wt[K][O] (a few MB, rebuilt per call by transpose_kernel);
//   DW : the reduction index r is the slow index of both z and ds, so both
//        operands are MN-major: 3-D tensor maps (32 contiguous, r, block) drop
//        [blk][32 r][32] boxes in TMA's SWIZZLE_128B_ATOM_32B pattern, which is the
//        SWIZZLE_128B_BASE32B descriptor layout - the one MN-major layout tf32 has.
// Shapes need I % 32 == 0 and O % 32 == 0; everything else (I = 3, O = 3) stays on
// the CUDA-core kernels of gemm_simt.cu.
#include "gemm.h"
#include "tc_common.cuh"

namespace vcm {
namespace {

constexpr int NTHREADS = 256;

enum { MODE_FWD = 0, MODE_DZ = 1, MODE_DW = 2 };

struct TcParams {
    // problem
    int64_t R;          // rows of the row-major activations
    int K, O, I;        // K = M*I
    int P_out;
    int num_kb;         // k-blocks of the whole reduction
    int kb_per_m;       // FWD: I/32 (k-blocks per basis), otherwise num_kb
    int kb_per_split;   // DW: k-blocks per CTA along r; otherwise num_kb
    // epilogue
    const float *bias;
    const float *res;
    float *y_act;
    float *out;         // FWD: out; DZ: dz; DW: partial [split][K][O]
    int per_point, act;
    float sy, sr;
    int direct;         // DW with a single split: write dW in the native [M][O][I] layout, no reduce pass
    float *partial;     // FWD with a split reduction: raw partial sums [split][R][O], epilogue runs afterwards
};


// --------------------------------------------------------------------- kernel --
// STAGES: depth of the TMA -> MMA ring. Long reductions (K2, K4) use 4 stages and own the SM; the dz GEMM
// has a short reduction (O/32 <= 16 k-blocks) and a large output tile, so it runs with 2 stages and two CTAs
// per SM: one CTA's epilogue stores overlap the other's loads and MMAs.
// RT: 128-row tiles per CTA sharing one B tile (RT = 2: a 256 x BN output tile in two TMEM accumulators).
// The dz GEMM is bound by operand re-reads from L2 (its reduction is short, its output huge); two row
// tiles per B tile cut that traffic by a third.
// X3: fp32-grade arithmetic on the tensor cores. The MMA reads only the top 19 bits of an fp32 operand (TF32), so
// every operand tile is split in shared memory into hi = the tile as loaded (the hardware truncates it) and
// lo = x - trunc_tf32(x) (exact in fp32), and each k-step issues hi*hi + lo*hi + hi*lo: the dropped lo*lo term and
// the truncation of lo are ~2^-22 relative. The split is elementwise, hence layout-agnostic (works on the swizzled
// tiles as they are); the four epilogue warps, idle during the main loop, do it between TMA arrival and MMA issue.
template <int BN, int MODE, int STAGES, int RT, bool X3>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const TcParams p) {
    constexpr int B_BYTES = BN * TK * 4;
    constexpr int HI_BYTES = RT * A_BYTES + B_BYTES;            // what TMA fills per stage
    constexpr int STAGE_BYTES = (X3 ? 2 : 1) * HI_BYTES;        // + the lo copies
    static_assert(RT == 1 || MODE != MODE_DW, "row-tile pairing is for the row-major GEMMs");
    static_assert(RT * BN <= 512, "TMEM has 512 columns");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms (8 rows x 128 B)
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *acc_full = empty + STAGES;
    uint64_t *split_done = acc_full + 1;                         // [STAGES], X3 only
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(split_done + STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, n_tile = blockIdx.y, split = blockIdx.z;
    const int kb_begin = split * p.kb_per_split;            // kb_per_split == num_kb when the reduction is not split
    const int kb_end = min(kb_begin + p.kb_per_split, p.num_kb);
    const int n_iter = kb_end - kb_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&split_done[s], 128);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, RT * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem_base = smem_u32(smem);

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int it = 0; it < n_iter; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_expect_tx(&full[s], HI_BYTES);
            const uint32_t a_dst = smem_base + s * STAGE_BYTES, b_dst = a_dst + RT * A_BYTES;
            const int kb = kb_begin + it;
            if (MODE == MODE_DW) {
                // A: z  as (32 kz, r, kz-block): box (32, 32 r, 4 blocks)   -> [4][32 r][32]
                // B: ds as (32 o,  r, o-block) : box (32, 32 r, BN/32)      -> [BN/32][32 r][32]
                tma_load_3d(a_dst, &map_a, &full[s], 0, kb * TK, m_tile * (TM / 32));
                tma_load_3d(b_dst, &map_b, &full[s], 0, kb * TK, n_tile * (BN / 32));
            } else {
                // A: row-major activations [R][Kred]: box (32 k, 128 r)
#pragma unroll
                for (int t = 0; t < RT; ++t)
                    tma_load_2d(a_dst + t * A_BYTES, &map_a, &full[s], kb * TK, (m_tile * RT + t) * TM);
                // B: (inner 32, rows, basis): FWD cuts W[m][o][i]; DZ reads Wt[k][o] with one "basis"
                const int m = kb / p.kb_per_m;
                tma_load_3d(b_dst, &map_b, &full[s], (kb - m * p.kb_per_m) * TK, n_tile * BN, m);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc(BN, MODE == MODE_DW, MODE == MODE_DW);
        for (int it = 0; it < n_iter; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(X3 ? &split_done[s] : &full[s], ph);
            tc_fence_after();
            const uint32_t a_base = smem_base + s * STAGE_BYTES, b_base = a_base + RT * A_BYTES;
#pragma unroll
            for (int tj = 0; tj < RT * (TK / 8); ++tj) {
                const int t = tj / (TK / 8), j = tj % (TK / 8);
                uint64_t da, db;
                if (MODE == MODE_DW) {
                    // MN-major tf32 only exists as SWIZZLE_128B_BASE32B (32-byte chunks XOR row % 4, written by
                    // TMA's SWIZZLE_128B_ATOM_32B): 32-element MN blocks 4096 B apart (LBO), 4-row k groups
                    // 512 B apart (SBO); one K = 8 instruction covers 8 r-rows = 1024 B.
                    da = make_desc(a_base + j * 1024, 4096, 512, LT_SW128_32B);
                    db = make_desc(b_base + j * 1024, 4096, 512, LT_SW128_32B);
                } else {
                    // K-major SW128: 8-row groups 1024 B apart (SBO); k advances 32 B inside the swizzle row
                    da = make_desc(a_base + t * A_BYTES + j * 32, 16, 1024);
                    db = make_desc(b_base + j * 32, 16, 1024);
                }
                umma_tf32(tmem_base + (uint32_t)(t * BN), da, db, idesc, (it > 0 || j > 0) ? 1u : 0u);
                if (X3) {
                    // the lo tiles sit HI_BYTES further (descriptor address field is in 16-byte units)
                    const uint64_t lo_off = (uint64_t)(HI_BYTES >> 4);
                    umma_tf32(tmem_base + (uint32_t)(t * BN), da + lo_off, db, idesc, 1u);
                    umma_tf32(tmem_base + (uint32_t)(t * BN), da, db + lo_off, idesc, 1u);
                }
            }
            umma_commit(&empty[s]);            // frees the smem stage when these MMAs retire
        }
        umma_commit(acc_full);                 // accumulator complete
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int q = warp - 4;                // TMEM lane quarter of this warp
        const int row = q * 32 + lane;         // accumulator row owned by this thread
        if (X3) {
            // ===== operand splitter (main loop) =====
            const int t128 = threadIdx.x - 128;
            for (int it = 0; it < n_iter; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                mbar_wait(&full[s], ph);                           // TMA bytes have landed
                float4 *hi = reinterpret_cast<float4 *>(smem + (size_t)s * STAGE_BYTES);
                float4 *lo = reinterpret_cast<float4 *>(smem + (size_t)s * STAGE_BYTES + HI_BYTES);
#pragma unroll 4
                for (int e = t128; e < HI_BYTES / 16; e += 128) {
                    const float4 v = hi[e];
                    float4 r;
                    r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                    r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                    r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                    r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                    lo[e] = r;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> MMA (async proxy)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&split_done[s])) : "memory");
            }
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int tc = 0; tc < RT * (BN / 32); ++tc) {
            const int t = tc / (BN / 32), c = tc - t * (BN / 32);
            const int64_t gr = ((int64_t)m_tile * RT + t) * TM + row;   // FWD/DZ: activation row; DW: kz
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * BN + c * 32), v);
            const int col0 = n_tile * BN + c * 32;
            if (MODE == MODE_FWD && p.partial != nullptr) {
                if (gr < p.R && col0 < p.O) {
                    float *dst = p.partial + ((size_t)split * p.R + gr) * p.O + col0;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        st4(dst + 4 * j4, make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]));
                }
            } else if (MODE == MODE_FWD) {
                // lane = accumulator row: each thread streams its own row segment (bias, residual, y, out). The
                // rows of a warp are O*4 bytes apart, i.e. one dense region for these layer widths; a transposed,
                // lane = column variant was measured slower in the full step (profiles/README.md).
                if (gr < p.R && col0 < p.O) {
                    const size_t at = (size_t)gr * p.O + col0;
                    const float *bp = p.bias ? p.bias + (p.per_point ? (size_t)(gr % p.P_out) * p.O : 0) + col0 : nullptr;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float t[4] = {v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]};
                        if (bp) {
                            const float4 b = ldg4(bp + 4 * j4);
                            t[0] += b.x; t[1] += b.y; t[2] += b.z; t[3] += b.w;
                        }
                        if (p.act) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) t[e] = elu1(t[e]);
                        }
                        if (p.y_act) st4(p.y_act + at + 4 * j4, make_float4(t[0], t[1], t[2], t[3]));
                        float4 o = make_float4(t[0] * p.sy, t[1] * p.sy, t[2] * p.sy, t[3] * p.sy);
                        if (p.res) {
                            const float4 rr = ldg4(p.res + at + 4 * j4);
                            o.x = fmaf(rr.x, p.sr, o.x); o.y = fmaf(rr.y, p.sr, o.y);
                            o.z = fmaf(rr.z, p.sr, o.z); o.w = fmaf(rr.w, p.sr, o.w);
                        }
                        st4(p.out + at + 4 * j4, o);
                    }
                }
            } else if (MODE == MODE_DZ) {
                // dz is the big output of the backward pass: stage the 128 x 32 chunk in (free) pipeline smem
                // in the 128B-swizzled layout and let the TMA engine write it as full lines, asynchronously.
                // A lane = row register layout stored directly would be 32-way address-divergent.
                const uint32_t stage_buf = smem_base + (uint32_t)(tc & 1) * A_BYTES;
                if (tc >= 2) {                                    // the buffer's previous TMA store has read it
                    if (threadIdx.x == 128) tma_store_wait_read<1>();
                    epilogue_bar();
                }
                const uint32_t rowaddr = stage_buf + (uint32_t)row * 128u;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)((j4 ^ (row & 7)) << 4)),
                                 "f"(v[4 * j4]), "f"(v[4 * j4 + 1]), "f"(v[4 * j4 + 2]), "f"(v[4 * j4 + 3])
                                 : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                epilogue_bar();
                if (threadIdx.x == 128) tma_store_2d(&map_c, stage_buf, col0, (int)(((int64_t)m_tile * RT + t) * TM));
            } else {
                // Written in dW's native [M][O][I] layout (lanes = consecutive kz = consecutive i of one basis,
                // so every store below is one 128-byte line); with a split reduction each split owns one such
                // slab and sum_partials_kernel adds the slabs in fixed order.
                if (gr < p.K && col0 < p.O) {
                    const int m = (int)gr / p.I, i = (int)gr - m * p.I;
                    float *dst = p.out + (size_t)split * p.K * p.O + ((size_t)m * p.O + col0) * p.I + i;
#pragma unroll
                    for (int j = 0; j < 32; ++j) dst[(size_t)j * p.I] = n_iter > 0 ? v[j] : 0.f;
                }
            }
        }
        if (MODE == MODE_DZ && threadIdx.x == 128) tma_store_wait_all();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, RT * BN);
    }
}

// Second pass of a split-reduction forward GEMM: fixed-order sum of the partials, then the K2 epilogue.
__global__ void splitk_fwd_epilogue_kernel(const float *__restrict__ part, int n_split, int64_t R, int O, int P_out,
                                           const float *__restrict__ bias, int per_point,
                                           const float *__restrict__ res, float *__restrict__ y_act,
                                           float *__restrict__ out, int act, float sy, float sr) {
    const size_t e4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const size_t n = (size_t)R * O;
    if (e4 >= n) return;
    float4 a = ldg4(part + e4);
    for (int k = 1; k < n_split; ++k) {
        const float4 b = ldg4(part + (size_t)k * n + e4);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float t[4] = {a.x, a.y, a.z, a.w};
    if (bias) {
        const size_t r = e4 / O;
        const int o = (int)(e4 - r * O);
        const float4 b = ldg4(bias + (per_point ? (size_t)(r % P_out) * O : 0) + o);
        t[0] += b.x; t[1] += b.y; t[2] += b.z; t[3] += b.w;
    }
    if (act) {
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = elu1(t[i]);
    }
    if (y_act) st4(y_act + e4, make_float4(t[0], t[1], t[2], t[3]));
    float4 o4 = make_float4(t[0] * sy, t[1] * sy, t[2] * sy, t[3] * sy);
    if (res) {
        const float4 rr = ldg4(res + e4);
        o4.x = fmaf(rr.x, sr, o4.x); o4.y = fmaf(rr.y, sr, o4.y); o4.z = fmaf(rr.z, sr, o4.z); o4.w = fmaf(rr.w, sr, o4.w);
    }
    st4(out + e4, o4);
}

__global__ void sum_partials_kernel(const float *__restrict__ part, float *__restrict__ out, size_t n, int n_split) {
    const size_t e4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (e4 >= n) return;
    float4 a = ldg4(part + e4);
#pragma unroll 8
    for (int k = 1; k < n_split; ++k) {
        const float4 b = ldg4(part + (size_t)k * n + e4);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    st4(out + e4, a);
}

// Many slabs of a small output: one warp per float4, lane l adds slabs l, l + 32, ... and the 32 partial sums meet in a
// fixed butterfly order (bitwise reproducible): 148 slabs cost 5 dependent loads instead of 19 rounds of 8.
__global__ void sum_partials_wide_kernel(const float *__restrict__ part, float *__restrict__ out, size_t n, int n_split) {
    const size_t e4 = (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4;
    const int lane = threadIdx.x & 31;
    if (e4 >= n) return;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = lane; k < n_split; k += 32) {
        const float4 b = ldg4(part + (size_t)k * n + e4);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
#pragma unroll
    for (int s2 = 16; s2 >= 1; s2 >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, s2);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, s2);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, s2);
        a.w += __shfl_xor_sync(0xffffffffu, a.w, s2);
    }
    if (lane == 0) st4(out + e4, a);
}

// Wt[k = m*I + i][o] = W[m][o][i]  (the stacked bases [M*I, O] of the north_star)
__global__ void transpose_bases_kernel(const float *__restrict__ W, float *__restrict__ Wt, int M, int I, int O) {
    __shared__ float tile[32][33];
    const int m = blockIdx.z, i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int o = o0 + r, i = i0 + tx;
        tile[r][tx] = (o < O && i < I) ? __ldg(W + ((size_t)m * O + o) * I + i) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, o = o0 + tx;
        if (i < I && o < O) Wt[((size_t)m * I + i) * O + o] = tile[tx][r];
    }
}

constexpr int stages_for(int mode, int bn, bool x3) {
    // dz: short reduction, big output -> 2 stages (two CTAs per SM when they fit); K2 / K4: as deep as fits in 227 KB
    if (mode == MODE_DZ) return 2;
    if (!x3) return 4;
    return bn >= 256 ? 2 : (bn >= 128 ? 3 : 4);
}

template <int MODE, int BN_, bool X3>
int launch_one(dim3 grid, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mc, const TcParams &p,
               cudaStream_t st) {
    constexpr int ST = stages_for(MODE, BN_, X3);
    auto k = tc_gemm_kernel<BN_, MODE, ST, 1, X3>;
    const int smem = ST * (X3 ? 2 : 1) * (A_BYTES + BN_ * TK * 4) + 1024 + 512;
    VCM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<<<grid, NTHREADS, smem, st>>>(ma, mb, mc, p);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

template <int MODE>
int launch(int bn, bool x3, dim3 grid, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mc,
           const TcParams &p, cudaStream_t st) {
    switch (bn) {
        case 32: return x3 ? launch_one<MODE, 32, true>(grid, ma, mb, mc, p, st) : launch_one<MODE, 32, false>(grid, ma, mb, mc, p, st);
        case 64: return x3 ? launch_one<MODE, 64, true>(grid, ma, mb, mc, p, st) : launch_one<MODE, 64, false>(grid, ma, mb, mc, p, st);
        case 128: return x3 ? launch_one<MODE, 128, true>(grid, ma, mb, mc, p, st) : launch_one<MODE, 128, false>(grid, ma, mb, mc, p, st);
        case 256: return x3 ? launch_one<MODE, 256, true>(grid, ma, mb, mc, p, st) : launch_one<MODE, 256, false>(grid, ma, mb, mc, p, st);
    }
    return fail(VCM_EUNSUPPORTED, "no tensor-core tile of width %d", bn);
}

struct DwSplit { int n_split, kb_per_split, num_kb; };
DwSplit plan_dw_tc(int64_t R, int K, int O) {
    DwSplit s;
    s.num_kb = (int)ceil_div(R, TK);
    const int64_t tiles = ceil_div(K, TM) * ceil_div(O, pick_bn(O));
    // Enough output tiles to fill the machine: one pass, written straight into dW. Otherwise split the row
    // reduction (wave-quantisation aware; partial slabs are summed in fixed order afterwards).
    int64_t split = 1;
    static const int policy = tune_flag("VCM_SPLIT_POLICY", 1);
    if (policy == 0) {
        split = tiles * 10 >= (int64_t)sm_count() * 7 ? 1 : ceil_div(3 * (int64_t)sm_count(), 2 * tiles);
        if (split > s.num_kb / 2) split = s.num_kb / 2;
        if (split > 24) split = 24;
        if (split < 1) split = 1;
    } else if (tiles < 2 * (int64_t)sm_count()) {
        // at most 24 slabs, except for small outputs (the residual projections: K*O = 32x64 floats, one output tile,
        // 28976 rows to reduce), where 4 MB of slabs allow a split that fills the machine
        int64_t cap = ((int64_t)1 << 20) / ((int64_t)K * O);
        if (cap < 24) cap = 24;
        if (cap > 2 * (int64_t)sm_count()) cap = 2 * (int64_t)sm_count();
        if (cap > s.num_kb / 2) cap = s.num_kb / 2;
        split = best_split(tiles, s.num_kb, (int)cap, sm_count());
    }
    s.kb_per_split = (int)ceil_div(s.num_kb, split);
    s.n_split = (int)ceil_div(s.num_kb, s.kb_per_split);
    return s;
}

}  // namespace

bool tc_supported(int which, int64_t R, int M, int I, int O, int precision) {
    if (precision != VCM_PREC_TF32 && precision != VCM_PREC_TF32X3) return false;
    // bring-up switch: VCM_TC_DISABLE is a bit mask (1 fwd, 2 dz, 4 dW) of GEMMs kept on the CUDA-core kernels
    static const int disabled = [] { const char *e = getenv("VCM_TC_DISABLE"); return e ? atoi(e) : 0; }();
    if (disabled & (1 << which)) return false;
    if (I % 32 != 0 || O % 32 != 0) return false;
    if (R < 1 || R >= ((int64_t)1 << 31)) return false;
    (void)which; (void)M;
    return encode_fn() != nullptr;
}

// Few row tiles but a long reduction (the 51-vertex bottleneck layers: 816 rows, K = 8704): split K so
// that the machine is filled; partial sums go through a second pass that also runs the epilogue.
static int fwd_splits(int64_t R, int K, int O) {
    const int64_t tiles = ceil_div(R, TM) * ceil_div(O, pick_bn(O));
    const int num_kb = K / TK;
    static const int policy = tune_flag("VCM_SPLIT_POLICY", 1);
    if (policy == 0) {
        if (tiles * 2 > sm_count() || num_kb < 16) return 1;
        int64_t split = ceil_div((int64_t)sm_count(), tiles);
        if (split > num_kb / 4) split = num_kb / 4;
        return split < 1 ? 1 : (int)split;
    }
    if (tiles >= 2 * (int64_t)sm_count() || num_kb < 16) return 1;
    return best_split(tiles, num_kb, num_kb / 4 < 32 ? num_kb / 4 : 32, sm_count());
}

int splitk_fwd_epilogue(const float *part, int n_split, int64_t R, int O, int P_out, const float *bias, int per_point,
                        const float *res, float *y_act, float *out, int act, float sy, float sr, cudaStream_t st) {
    splitk_fwd_epilogue_kernel<<<(unsigned)ceil_div(R * O / 4, 256), 256, 0, st>>>(part, n_split, R, O, P_out, bias,
                                                                                  per_point, res, y_act, out, act, sy,
                                                                                  sr);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

size_t tc_basis_fwd_workspace(int64_t R, int M, int I, int O) {
    const int s = fwd_splits(R, M * I, O);
    return s > 1 ? (size_t)s * R * O : 0;
}

int tc_basis_fwd(const float *z, const float *W, const float *bias, int per_point, const float *res, float *y_act,
                 float *out, float *ws, int64_t R, int P_out, int M, int I, int O, int act, float sy, float sr,
                 int precision, cudaStream_t st) {
    const int K = M * I, bn = pick_bn(O);
    CUtensorMap ma, mb;
    {
        const uint64_t d[2] = {(uint64_t)K, (uint64_t)R}, s[2] = {4, (uint64_t)K * 4};
        const uint32_t b[2] = {TK, TM};
        if (int e = make_map(&ma, z, 2, d, s, b)) return e;
    }
    {
        const uint64_t d[3] = {(uint64_t)I, (uint64_t)O, (uint64_t)M}, s[3] = {4, (uint64_t)I * 4, (uint64_t)O * I * 4};
        const uint32_t b[3] = {TK, (uint32_t)bn, 1};
        if (int e = make_map(&mb, W, 3, d, s, b)) return e;
    }
    TcParams p{};
    p.R = R; p.K = K; p.O = O; p.I = I; p.P_out = P_out;
    int n_split = fwd_splits(R, K, O);
    if (n_split > 1 && ws == nullptr) n_split = 1;
    p.num_kb = K / TK; p.kb_per_m = I / TK;
    p.kb_per_split = (int)ceil_div(p.num_kb, n_split);
    n_split = (int)ceil_div(p.num_kb, p.kb_per_split);
    p.bias = bias; p.res = res; p.y_act = y_act; p.out = out; p.per_point = per_point; p.act = act; p.sy = sy; p.sr = sr;
    p.partial = n_split > 1 ? ws : nullptr;
    dim3 grid((unsigned)ceil_div(R, TM), (unsigned)ceil_div(O, bn), (unsigned)n_split);
    if (int e = launch<MODE_FWD>(bn, precision == VCM_PREC_TF32X3, grid, ma, mb, ma, p, st)) return e;
    if (n_split > 1) return splitk_fwd_epilogue(ws, n_split, R, O, P_out, bias, per_point, res, y_act, out, act, sy, sr, st);
    return VCM_OK;
}

size_t tc_basis_bwd_dz_workspace(int M, int I, int O) { return (size_t)M * I * O; }

int tc_stack_bases(const float *W, float *Wt, int M, int I, int O, cudaStream_t st) {
    dim3 tg((unsigned)ceil_div(I, 32), (unsigned)ceil_div(O, 32), (unsigned)M);
    transpose_bases_kernel<<<tg, 256, 0, st>>>(W, Wt, M, I, O);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

int tc_basis_bwd_dz(const float *ds, const float *W, float *dz, float *ws, int64_t R, int M, int I, int O,
                    int accumulate, int precision, cudaStream_t st) {
    if (accumulate) return fail(VCM_EUNSUPPORTED, "tensor-core dz does not accumulate");
    const int K = M * I, bn = pick_bn(K);
    // stacked bases Wt[K][O]: built here, or already in `ws` when W is NULL (vcm_basis_bwd_dz_stacked)
    if (W != nullptr)
        if (int e = tc_stack_bases(W, ws, M, I, O, st)) return e;
    CUtensorMap ma, mb;
    {
        const uint64_t d[2] = {(uint64_t)O, (uint64_t)R}, s[2] = {4, (uint64_t)O * 4};
        const uint32_t b[2] = {TK, TM};
        if (int e = make_map(&ma, ds, 2, d, s, b)) return e;
    }
    {
        const uint64_t d[3] = {(uint64_t)O, (uint64_t)K, 1}, s[3] = {4, (uint64_t)O * 4, (uint64_t)K * O * 4};
        const uint32_t b[3] = {TK, (uint32_t)bn, 1};
        if (int e = make_map(&mb, ws, 3, d, s, b)) return e;
    }
    TcParams p{};
    p.R = R; p.K = K; p.O = O; p.I = I;
    p.num_kb = O / TK; p.kb_per_m = p.num_kb; p.kb_per_split = p.num_kb;
    p.out = dz;
    CUtensorMap mc;                                   // store map over dz [R][K]: box (32 columns, 128 rows)
    {
        const uint64_t d[2] = {(uint64_t)K, (uint64_t)R}, s2[2] = {4, (uint64_t)K * 4};
        const uint32_t b[2] = {TK, TM};
        if (int e = make_map(&mc, dz, 2, d, s2, b)) return e;
    }
    dim3 grid((unsigned)ceil_div(R, TM), (unsigned)ceil_div(K, bn), 1);
    return launch<MODE_DZ>(bn, precision == VCM_PREC_TF32X3, grid, ma, mb, mc, p, st);
}

size_t tc_basis_bwd_dw_workspace(int64_t R, int M, int I, int O, int precision) {
    const DwSplit s = plan_dw_tc(R, M * I, O);
    return (size_t)s.n_split * M * I * O;
}

int tc_basis_bwd_dw(const float *ds, const float *z, float *dW, float *ws, int64_t R, int M, int I, int O,
                    int precision, cudaStream_t st) {
    const int K = M * I, bn = pick_bn(O);
    const DwSplit sp = plan_dw_tc(R, K, O);
    CUtensorMap ma, mb;
    {
        const uint64_t d[3] = {32, (uint64_t)R, (uint64_t)(K / 32)}, s[3] = {4, (uint64_t)K * 4, 128};
        const uint32_t b[3] = {32, TK, TM / 32};
        if (int e = make_map(&ma, z, 3, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
    }
    {
        const uint64_t d[3] = {32, (uint64_t)R, (uint64_t)(O / 32)}, s[3] = {4, (uint64_t)O * 4, 128};
        const uint32_t b[3] = {32, TK, (uint32_t)(bn / 32)};
        if (int e = make_map(&mb, ds, 3, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
    }
    TcParams p{};
    p.R = R; p.K = K; p.O = O; p.I = I;
    p.num_kb = sp.num_kb; p.kb_per_m = sp.num_kb; p.kb_per_split = sp.kb_per_split;
    p.direct = sp.n_split == 1;
    p.out = p.direct ? dW : ws;
    dim3 grid((unsigned)ceil_div(K, TM), (unsigned)ceil_div(O, bn), (unsigned)sp.n_split);
    if (int e = launch<MODE_DW>(bn, precision == VCM_PREC_TF32X3, grid, ma, mb, ma, p, st)) return e;
    if (p.direct) return VCM_OK;
    const size_t n = (size_t)K * O;
    if (sp.n_split > 32)
        sum_partials_wide_kernel<<<(unsigned)ceil_div((int64_t)(n / 4) * 32, 256), 256, 0, st>>>(ws, dW, n, sp.n_split);
    else
        sum_partials_kernel<<<(unsigned)ceil_div((int64_t)(n / 4), 256), 256, 0, st>>>(ws, dW, n, sp.n_split);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

}  // namespace vcm

// ---------------------------------------------------------------------------
// Bring-up harness (not part of the product path): loads ONE k-block of the dW
// operands through the production tensor maps, dumps the two shared-memory tiles,
// issues the four MMAs with caller-chosen descriptor parameters and dumps the
// accumulator. Lets the MN-major descriptor encoding be probed in one GPU run.
namespace vcm {
namespace {
__global__ void __launch_bounds__(128, 1)
debug_dw_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float *dump_a,
                float *dump_b, float *dump_d, int bn, uint32_t idesc, uint32_t lbo, uint32_t sbo, uint32_t jstride,
                int kb, uint32_t ltype) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + A_BYTES + 256 * TK * 4);
    uint64_t *done = full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_dst = smem_u32(smem), b_dst = a_dst + A_BYTES;
    if (threadIdx.x == 0) {
        mbar_expect_tx(full, A_BYTES + bn * TK * 4);
        tma_load_3d(a_dst, &map_a, full, 0, kb * TK, 0);
        tma_load_3d(b_dst, &map_b, full, 0, kb * TK, 0);
    }
    mbar_wait(full, 0);
    __syncthreads();
    for (int e = threadIdx.x; e < A_BYTES / 4; e += blockDim.x) dump_a[e] = reinterpret_cast<float *>(smem)[e];
    for (int e = threadIdx.x; e < bn * TK; e += blockDim.x) dump_b[e] = reinterpret_cast<float *>(smem + A_BYTES)[e];
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        for (int j = 0; j < 4; ++j)
            umma_tf32(tmem_base, make_desc(a_dst + j * jstride, lbo, sbo, ltype),
                      make_desc(b_dst + j * jstride, lbo, sbo, ltype), idesc, j > 0 ? 1u : 0u);
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    for (int c = 0; c < bn / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        for (int i = 0; i < 32; ++i) dump_d[(size_t)(warp * 32 + lane) * bn + c * 32 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}
}  // namespace
}  // namespace vcm

extern "C" VCM_API int vcm_debug_dw_stage(const float *ds, const float *z, float *dump_a, float *dump_b, float *dump_d,
                                          int64_t R, int K, int O, int a_mn, int b_mn, int lbo, int sbo, int jstride,
                                          int kb, int tma_swz32, int ltype, void *stream) {
    using namespace vcm;
    const int bn = pick_bn(O);
    CUtensorMap ma, mb;
    {
        const uint64_t d[3] = {32, (uint64_t)R, (uint64_t)(K / 32)}, s[3] = {4, (uint64_t)K * 4, 128};
        const uint32_t b[3] = {32, TK, TM / 32};
        if (int e = make_map(&ma, z, 3, d, s, b, tma_swz32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    }
    {
        const uint64_t d[3] = {32, (uint64_t)R, (uint64_t)(O / 32)}, s[3] = {4, (uint64_t)O * 4, 128};
        const uint32_t b[3] = {32, TK, (uint32_t)(bn / 32)};
        if (int e = make_map(&mb, ds, 3, d, s, b, tma_swz32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    }
    const int smem = A_BYTES + 256 * TK * 4 + 1024 + 256;
    VCM_CUDA(cudaFuncSetAttribute(debug_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    debug_dw_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(ma, mb, dump_a, dump_b, dump_d, bn,
                                                            make_idesc(bn, a_mn, b_mn), lbo, sbo, jstride, kb, (uint32_t)ltype);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}
