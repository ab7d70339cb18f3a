// K1 + K2 in one kernel: the vertex-wise contraction z = A (x) gather(x) is produced tile by tile in shared
// memory, directly in the tcgen05 operand layout, and consumed by the tensor cores; z only goes to HBM when the
// caller wants it for the backward pass (one write, no read-back), and never for inference.
//
//   out[b,p,o] = epi( sum_{m,i} ( sum_n A[p,n,m] x[b, idx[p,n], i] ) W[m,o,i] )      graphVAESSW.py:111-138,161
//
// Tile: 128 accumulator rows = PV vertices x B batch entries (row = v*B + b), BN output channels.
// Reduction order: k = (i-chunk of 8, basis m, i within chunk). One "chunk" is 8 input channels of all M bases;
// it is cut into G = ceil(M/4) k-steps of [4 bases x 8 channels] = 32 tf32 = one 128-byte swizzle row.
//   warps 0-7   producers: thread = (row, half of the 8-channel chunk). Neighbour rows of x arrive through a
//               per-thread cp.async double buffer (bursts of <= 8, no registers held across the L2 latency); coefficients come
//               from shared memory pre-duplicated as (a, a) pairs so that the contraction is pure FFMA2; at the
//               end of a chunk the 17 x 4 accumulators are written into the A stages (SWIZZLE_128B, K-major).
//               With STORE_Z the same registers also go to z[b][p][m][i] (16-byte streaming stores).
//   warp 8      TMA loads of the W tiles: 3-D map (i, o, m) over the native [M][O][I] parameter, box (8, BN, 4),
//               SWIZZLE_32B -> four [BN o][8 i] sub-tiles (one per basis) = K-major B operands of the four K = 8
//               MMAs of a k-step. (With a swizzled map TMA pads every inner box row to the swizzle span, so a
//               32-byte inner box only packs densely under the 32-byte swizzle.)
//   warp 9      tcgen05.mma issuer (kind::tf32, fp32 accumulators in TMEM).
//   warp 10     TMEM allocator.
//   epilogue    the 8 producer warps: tcgen05.ld -> bias / ELU / residual blend -> out (or split-K partials).
#include "gemm.h"
#include "tc_common.cuh"

namespace vcm {
namespace {

constexpr int FF_PRODUCERS = 256;
constexpr int FF_THREADS = FF_PRODUCERS + 96;
constexpr int FF_SEG = 8;                      // neighbours per gather burst (two bursts resident per thread)
constexpr int FF_MT = 17;                      // bases accumulated per thread
constexpr int FF_MP = 18;                      // padded to pairs
constexpr int FF_G = (FF_MT + 3) / 4;          // k-steps per chunk at M = 17
constexpr int FF_FIFO_BYTES = 2 * FF_SEG * FF_PRODUCERS * 16;

struct FusedParams {
    const float *x, *A;
    const int32_t *idx, *last;
    int p_in, p_out, n_max, B, I, M, O, PV;
    int n_chunks, chunks_per_split, G, tail_k8;
    const float *bias, *res;
    float *y_act, *out, *partial, *z;
    int per_point, act;
    float sy, sr;
    int64_t R;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int lds32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN, int STAGES, bool STORE_Z>
__global__ void __launch_bounds__(FF_THREADS, 1)
conv_fwd_fused_kernel(const __grid_constant__ CUtensorMap map_w, const FusedParams p) {
    constexpr int B_BYTES = BN * 128;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *fifo = smem + STAGES * STAGE_BYTES;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(fifo + FF_FIFO_BYTES);
    uint64_t *b_full = a_full + STAGES;
    uint64_t *empty = b_full + STAGES;
    uint64_t *acc_full = empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);
    float2 *coef_s = reinterpret_cast<float2 *>(reinterpret_cast<uint8_t *>(a_full) + 256);   // [PV][n_max][FF_MP]
    int32_t *idx_s = reinterpret_cast<int32_t *>(coef_s + (size_t)p.PV * p.n_max * FF_MP);    // [PV][n_max]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int p0 = blockIdx.x * p.PV, n_tile = blockIdx.y, split = blockIdx.z;
    const int c_begin = split * p.chunks_per_split;
    const int c_end = min(c_begin + p.chunks_per_split, p.n_chunks);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&a_full[s], FF_PRODUCERS / 32);
            mbar_init(&b_full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 10) tmem_alloc(tmem_slot, BN);
    if (tid < FF_PRODUCERS) {
        // coefficients of the tile's vertices, zero-padded to FF_MP bases and duplicated for FFMA2
        const int per_v = p.n_max * FF_MP;
        for (int e = tid; e < p.PV * per_v; e += FF_PRODUCERS) {
            const int v = e / per_v, rem = e - v * per_v, n = rem / FF_MP, m = rem - n * FF_MP;
            const int pv = p0 + v;
            const float a = (pv < p.p_out && m < p.M) ? __ldg(p.A + ((size_t)pv * p.n_max + n) * p.M + m) : 0.f;
            coef_s[e] = make_float2(a, a);
        }
        // neighbour -> element offset of its x row (q * I), -1 for the padding sentinel
        for (int e = tid; e < p.PV * p.n_max; e += FF_PRODUCERS) {
            const int pv = p0 + e / p.n_max;
            const int q = pv < p.p_out ? __ldg(p.idx + (size_t)p0 * p.n_max + e) : p.p_in;
            idx_s[e] = q != p.p_in ? q * p.I : -1;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem_base = smem_u32(smem);

    if (warp < FF_PRODUCERS / 32) {
        // ===== producers =====
        const int r = warp * 16 + (lane & 15), h = lane >> 4;
        const int v = r / p.B, b = r - v * p.B, pv = p0 + v;
        const bool live = v < p.PV && pv < p.p_out;
        const int l = live ? __ldg(p.last + pv) : 0;                       // warp-uniform (B % 16 == 0)
        const uint32_t idx_v = smem_u32(idx_s) + (uint32_t)((live ? v : 0) * p.n_max) * 4u;
        const uint32_t coef_v = smem_u32(coef_s) + (uint32_t)((live ? v : 0) * p.n_max) * (FF_MP * 8u);
        const float *xcol = p.x + (size_t)b * p.p_in * p.I + h * 4;       // this thread's 4 channels of chunk 0
        const uint32_t my_fifo = smem_u32(fifo) + (uint32_t)tid * 16u;

        // The neighbour list is cut into segments of <= FF_SEG; the loads of segment k+1 are issued in one burst
        // (one cp.async group) before segment k is consumed. Two reasons for bursts instead of a rolling
        // prefetch: the wait is a compile-time wait_group<1>, and the proxy fence of the flush below drains
        // every outstanding cp.async of the thread - with bursts the youngest one is a whole segment old.
        const int nseg = (l + FF_SEG - 1) / FF_SEG;
        const int seg_len = nseg ? (l + nseg - 1) / nseg : 0;
        auto issue_seg = [&](int c, int sg, int buf) {
            const int n0 = sg * seg_len, n1 = min(l, n0 + seg_len);
            const float *xc = xcol + c * 8;
            uint32_t dst = my_fifo + (uint32_t)(buf * FF_SEG) * (FF_PRODUCERS * 16);
            for (int n = n0; n < n1; ++n, dst += FF_PRODUCERS * 16) {
                const int off = lds32(idx_v + (uint32_t)n * 4u);
                if (off >= 0) cp_async16(dst, xc + off, 16u);
                else          // padding sentinel: contributes an exact zero
                    asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "f"(0.f) : "memory");
            }
        };
        if (nseg) issue_seg(c_begin, 0, 0);
        cp_async_commit();

        // z[b][pv][m][i]: this thread's 4 channels of every chunk (only the first N tile writes it)
        float *zrow = (STORE_Z && live && n_tile == 0) ? p.z + ((size_t)b * p.p_out + pv) * p.M * p.I + h * 4 : nullptr;
        const uint32_t row_off = (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        int step = 0, buf = 0;
        for (int c = c_begin; c < c_end; ++c) {
            float2 acc[FF_MT][2];
#pragma unroll
            for (int m = 0; m < FF_MT; ++m) acc[m][0] = acc[m][1] = make_float2(0.f, 0.f);
            for (int sg = 0; sg < nseg; ++sg) {
                const bool last_seg = sg + 1 == nseg;
                if (!last_seg || c + 1 < c_end) issue_seg(last_seg ? c + 1 : c, last_seg ? 0 : sg + 1, buf ^ 1);
                cp_async_commit();
                cp_async_wait<1>();
                const int n0 = sg * seg_len, n1 = min(l, n0 + seg_len);
                uint32_t src = my_fifo + (uint32_t)(buf * FF_SEG) * (FF_PRODUCERS * 16);
                uint32_t cf = coef_v + (uint32_t)n0 * (FF_MP * 8u);
#pragma unroll 2
                for (int n = n0; n < n1; ++n, src += FF_PRODUCERS * 16, cf += FF_MP * 8u) {
                    const float4 xv = lds128(src);
                    const float2 x01 = make_float2(xv.x, xv.y), x23 = make_float2(xv.z, xv.w);
#pragma unroll
                    for (int k = 0; k < FF_MP / 2; ++k) {
                        const float4 a = lds128(cf + k * 16);              // (a_2k, a_2k, a_2k+1, a_2k+1)
                        acc[2 * k][0] = fma2(make_float2(a.x, a.y), x01, acc[2 * k][0]);
                        acc[2 * k][1] = fma2(make_float2(a.x, a.y), x23, acc[2 * k][1]);
                        if (2 * k + 1 < FF_MT) {
                            acc[2 * k + 1][0] = fma2(make_float2(a.z, a.w), x01, acc[2 * k + 1][0]);
                            acc[2 * k + 1][1] = fma2(make_float2(a.z, a.w), x23, acc[2 * k + 1][1]);
                        }
                    }
                }
                buf ^= 1;
            }
            // the chunk's accumulators -> A stages: row r, 16-byte piece (m_local*2 + h) XOR (r % 8)
#pragma unroll
            for (int g = 0; g < FF_G; ++g) {
                if (g < p.G) {
                    const int s = step % STAGES;
                    const uint32_t ph = (uint32_t)(step / STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    const uint32_t rowaddr = smem_base + (uint32_t)s * STAGE_BYTES + row_off;
#pragma unroll
                    for (int ml = 0; ml < 4; ++ml) {
                        const int m = g * 4 + ml;
                        if (m < FF_MT)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                                         ::"r"(rowaddr + ((((uint32_t)(ml * 2 + h)) ^ sw) << 4)), "f"(acc[m][0].x),
                                         "f"(acc[m][0].y), "f"(acc[m][1].x), "f"(acc[m][1].y)
                                         : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[s]);
                    ++step;
                }
            }
            // z goes out after the hand-over: the fence above waits for the thread's outstanding memory
            // operations, so stores issued before it would put a DRAM round trip into every k-step
            if (STORE_Z && zrow != nullptr) {
#pragma unroll
                for (int m = 0; m < FF_MT; ++m)
                    if (m < p.M)
                        st4_cs(zrow + (size_t)m * p.I + c * 8,
                               make_float4(acc[m][0].x, acc[m][0].y, acc[m][1].x, acc[m][1].y));
            }
        }
        cp_async_wait<0>();

        // ===== epilogue =====
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int ev = row / p.B, eb = row - ev * p.B, epv = p0 + ev;
        const bool valid = ev < p.PV && epv < p.p_out;
        const int64_t gr = (int64_t)eb * p.p_out + epv;
#pragma unroll 1
        for (int cc = warp >> 2; cc < BN / 32; cc += 2) {
            float vv[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), vv);
            const int col0 = n_tile * BN + cc * 32;
            if (!valid || col0 >= p.O) continue;
            if (p.partial != nullptr) {
                float *dst = p.partial + ((size_t)split * p.R + gr) * p.O + col0;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    st4(dst + 4 * j4, make_float4(vv[4 * j4], vv[4 * j4 + 1], vv[4 * j4 + 2], vv[4 * j4 + 3]));
                continue;
            }
            const size_t at = (size_t)gr * p.O + col0;
            const float *bp = p.bias ? p.bias + (p.per_point ? (size_t)epv * p.O : 0) + col0 : nullptr;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                float t[4] = {vv[4 * j4], vv[4 * j4 + 1], vv[4 * j4 + 2], vv[4 * j4 + 3]};
                if (bp) {
                    const float4 bb = ldg4(bp + 4 * j4);
                    t[0] += bb.x; t[1] += bb.y; t[2] += bb.z; t[3] += bb.w;
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
        tc_fence_before();
    } else if (warp == 8) {
        if (lane == 0) {
            // ===== W tiles =====
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
            int step = 0;
            for (int c = c_begin; c < c_end; ++c)
                for (int g = 0; g < p.G; ++g, ++step) {
                    const int s = step % STAGES;
                    const uint32_t ph = (uint32_t)(step / STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&b_full[s], B_BYTES);
                    tma_load_3d(smem_base + (uint32_t)s * STAGE_BYTES + A_BYTES, &map_w, &b_full[s], c * 8, n_tile * BN,
                                g * 4);
                }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = make_idesc(BN, 0, 0);
            int step = 0;
            for (int c = c_begin; c < c_end; ++c)
                for (int g = 0; g < p.G; ++g, ++step) {
                    const int s = step % STAGES;
                    const uint32_t ph = (uint32_t)(step / STAGES) & 1u;
                    mbar_wait(&a_full[s], ph);
                    mbar_wait(&b_full[s], ph);
                    tc_fence_after();
                    const uint32_t a_base = smem_base + (uint32_t)s * STAGE_BYTES, b_base = a_base + A_BYTES;
                    const int nk = g == p.G - 1 ? p.tail_k8 : 4;
                    for (int j = 0; j < nk; ++j)
                        umma_tf32(tmem_base, make_desc(a_base + j * 32, 16, 1024),
                                  make_desc(b_base + j * (BN * 32), 16, 256, LT_SW32), idesc, (step > 0 || j > 0) ? 1u : 0u);
                    umma_commit(&empty[s]);
                }
            umma_commit(acc_full);
        }
    }
    __syncthreads();
    if (warp == 10) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

constexpr int ff_stages(int bn) { return bn >= 128 ? 3 : (bn >= 64 ? 4 : 5); }
int ff_bn(int O) { const int bn = pick_bn(O); return bn > 128 ? 128 : bn; }

size_t ff_smem(int bn, int PV, int n_max) {
    return (size_t)ff_stages(bn) * (A_BYTES + bn * 128) + FF_FIFO_BYTES + 256 + (size_t)PV * n_max * (FF_MP * 8 + 4) +
           1024 + 64;
}

struct FusedPlan { int PV, bn, n_tiles, n_chunks, G, tail_k8, n_split, chunks_per_split; int64_t tiles; size_t smem; };

bool plan_fused(const Tables *t, int B, int M, int I, int O, FusedPlan *pl) {
    if (B % 16 != 0 || B > TM || M < 1 || M > FF_MT || I % 8 != 0 || O % 32 != 0 || I < 8) return false;
    if (encode_fn() == nullptr) return false;
    pl->PV = TM / B;
    pl->bn = ff_bn(O);
    pl->n_tiles = (int)ceil_div(O, pl->bn);
    pl->n_chunks = I / 8;
    pl->G = (M + 3) / 4;
    pl->tail_k8 = M - 4 * (pl->G - 1);
    pl->tiles = ceil_div(t->p_out, pl->PV);
    pl->smem = ff_smem(pl->bn, pl->PV, t->n_max);
    if (pl->smem > 227 * 1024) return false;
    const int64_t ctas = pl->tiles * pl->n_tiles;
    int split = 1;
    if (ctas < 2 * (int64_t)sm_count() && pl->n_chunks >= 4)
        split = best_split(ctas, pl->n_chunks * pl->G, pl->n_chunks / 2 < 32 ? pl->n_chunks / 2 : 32, sm_count());
    // best_split works in k-steps; a split owns whole chunks
    pl->chunks_per_split = (int)ceil_div(pl->n_chunks, split);
    pl->n_split = (int)ceil_div(pl->n_chunks, pl->chunks_per_split);
    return true;
}

template <int BN_>
int launch_fused(bool store_z, dim3 grid, size_t smem, const CUtensorMap &mw, const FusedParams &p, cudaStream_t st) {
    constexpr int ST = ff_stages(BN_);
    if (store_z) {
        auto k = conv_fwd_fused_kernel<BN_, ST, true>;
        VCM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, FF_THREADS, smem, st>>>(mw, p);
    } else {
        auto k = conv_fwd_fused_kernel<BN_, ST, false>;
        VCM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, FF_THREADS, smem, st>>>(mw, p);
    }
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

}  // namespace

bool fused_fwd_supported(const Tables *t, int B, int M, int I, int O, int precision) {
    FusedPlan pl;
    return precision == VCM_PREC_TF32 && plan_fused(t, B, M, I, O, &pl);
}

size_t fused_fwd_workspace(const Tables *t, int B, int M, int I, int O) {
    FusedPlan pl;
    if (!plan_fused(t, B, M, I, O, &pl)) return 0;
    return pl.n_split > 1 ? (size_t)pl.n_split * B * t->p_out * O : 0;
}

int fused_fwd(const Tables *t, const float *x, const float *A, const float *W, const float *bias, int per_point,
              const float *res, float *y_act, float *out, float *z, float *ws, int B, int M, int I, int O, int act,
              float sy, float sr, cudaStream_t st) {
    FusedPlan pl;
    if (!plan_fused(t, B, M, I, O, &pl)) return fail(VCM_EUNSUPPORTED, "shape not supported by the fused forward");
    if (pl.n_split > 1 && ws == nullptr) return fail(VCM_EINVAL, "workspace required (%d splits)", pl.n_split);
    CUtensorMap mw;
    {
        // W[m][o][i] seen as (i, o, m): box (8 i, BN o, 4 m) lands as [m][o][8 i]: four K-major [BN][8] sub-tiles
        const uint64_t d[3] = {(uint64_t)I, (uint64_t)O, (uint64_t)M};
        const uint64_t s[3] = {4, (uint64_t)I * 4, (uint64_t)O * I * 4};
        const uint32_t b[3] = {8, (uint32_t)pl.bn, 4};
        if (int e = make_map(&mw, W, 3, d, s, b, CU_TENSOR_MAP_SWIZZLE_32B)) return e;
    }
    FusedParams p{};
    p.x = x; p.A = A; p.idx = t->idx; p.last = t->last;
    p.p_in = t->p_in; p.p_out = t->p_out; p.n_max = t->n_max; p.B = B; p.I = I; p.M = M; p.O = O; p.PV = pl.PV;
    p.n_chunks = pl.n_chunks; p.chunks_per_split = pl.chunks_per_split; p.G = pl.G; p.tail_k8 = pl.tail_k8;
    p.bias = bias; p.res = res; p.y_act = y_act; p.out = out; p.partial = pl.n_split > 1 ? ws : nullptr; p.z = z;
    p.per_point = per_point; p.act = act; p.sy = sy; p.sr = sr; p.R = (int64_t)B * t->p_out;
    dim3 grid((unsigned)pl.tiles, (unsigned)pl.n_tiles, (unsigned)pl.n_split);
    int e;
    switch (pl.bn) {
        case 32: e = launch_fused<32>(z != nullptr, grid, pl.smem, mw, p, st); break;
        case 64: e = launch_fused<64>(z != nullptr, grid, pl.smem, mw, p, st); break;
        default: e = launch_fused<128>(z != nullptr, grid, pl.smem, mw, p, st); break;
    }
    if (e) return e;
    if (pl.n_split > 1) return splitk_fwd_epilogue(ws, pl.n_split, p.R, O, t->p_out, bias, per_point, res, y_act, out,
                                                   act, sy, sr, st);
    return VCM_OK;
}

}  // namespace vcm
