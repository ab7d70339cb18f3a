// Data-parallel gradient reduction fused with the optimizer over NVLink peer memory (graphVAE_train.py:249, 260,
// 315-316: the reference sums the replicas' gradients on GPU 0 and runs Adam there).
//
// Every rank's flat gradient buffer, flat parameter buffer and a small flag buffer live in symmetric memory: each GPU
// of the node maps every peer's buffers (NVLink 5 through the NVSwitch). For one contiguous range [offset, offset + n)
// of the flat buffers ONE kernel per rank does
//
//   barrier (all ranks have finished writing their gradients of the range)
//   for the 1/world slice this rank owns:  g = g_0 + g_1 + ... + g_{world-1}      128-bit loads straight out of the
//                                                                                   peers' gradient buffers, fixed order
//                                          Adam on the slice (moments of the slice are resident on this rank only)
//                                          the updated parameters -> every rank's parameter buffer (peer stores)
//   barrier (all ranks have received all slices)
//
// i.e. reduce-scatter + optimizer + all-gather with one barrier pair, 1/world of the Adam work per rank, the same
// bytes over NVLink as an all-reduce, and no gradient sum ever written back. Each element is summed by exactly one
// rank in rank order, so all replicas receive bit-identical parameters and the result equals a single process that
// accumulates the shards' gradients in rank order. The barriers are per-CTA flag exchanges on the peers' flag buffers
// (system-scope acquire/release CAS, 0 -> 1 by the sender, 1 -> 0 by the receiver), bounded: a rank that never shows
// up aborts the launch after ~4 s instead of hanging the GPU. Plain kernel launches: capturable in the step graph.
#include <stdlib.h>

#include "adam.cuh"
#include "common.cuh"

namespace vcm {
namespace {

constexpr int kMaxWorld = 8;
constexpr int kThreadsC = 256;
constexpr size_t kChunk4 = 1024;     // ownership chunk: 1024 float4 = 16 KB
constexpr int kUnroll = 4;           // kChunk4 = kThreadsC * kUnroll: one pass of a CTA per chunk

struct Peers {
    const float *grad[kMaxWorld];
    float *param[kMaxWorld];
    uint32_t *flag[kMaxWorld];
};

__device__ __forceinline__ void flag_put(uint32_t *addr) {          // 0 -> 1, release
    const long long t0 = clock64();
    uint32_t old;
    do {
        asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
        if (old != 0 && clock64() - t0 > 8000000000LL) __trap();
    } while (old != 0);
}
__device__ __forceinline__ void flag_wait(uint32_t *addr) {         // 1 -> 0, acquire
    const long long t0 = clock64();
    uint32_t old;
    do {
        asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
        if (old != 1 && clock64() - t0 > 8000000000LL) __trap();
    } while (old != 1);
}
// CTA b of every rank meets CTA b of every other rank
__device__ __forceinline__ void peer_barrier(const Peers &P, int rank, int world, int slot) {
    __syncthreads();
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
        const int peer = threadIdx.x;
        flag_put(P.flag[peer] + (size_t)slot * kMaxWorld + rank);
        flag_wait(P.flag[rank] + (size_t)slot * kMaxWorld + peer);
    }
    __syncthreads();
}

__device__ __forceinline__ float4 ld_peer(const float *p) {         // never served from a stale L1 line
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// Ownership is block-cyclic over the ABSOLUTE element index (chunks of kChunk4 float4 = 16 KB): chunk c of the flat
// buffers belongs to rank c % world, whatever range a launch covers. A rank's Adam moments are therefore live exactly
// on the chunks it owns, independent of how the step is cut into ranges.
__global__ void __launch_bounds__(kThreadsC)
reduce_adam_bcast_kernel(const Peers P, float *__restrict__ m, float *__restrict__ v, size_t lo4, size_t hi4,
                         const float *__restrict__ state, float b1, float b2, float eps, int rank, int world, int slot0) {
    const int slot = slot0 + blockIdx.x;
    peer_barrier(P, rank, world, slot);
    const float lr = state[1], bc1 = state[2], bc2_sqrt = state[3];
    const float step_size = lr / bc1;
    const size_t c_first = lo4 / kChunk4, c_last = (hi4 - 1) / kChunk4;
    size_t c = c_first + (size_t)((rank + world - (int)(c_first % world)) % world) + (size_t)blockIdx.x * world;
    for (; c <= c_last; c += (size_t)gridDim.x * world) {
        const size_t a = c * kChunk4 > lo4 ? c * kChunk4 : lo4;
        const size_t b = (c + 1) * kChunk4 < hi4 ? (c + 1) * kChunk4 : hi4;
        // one pass per chunk: kUnroll float4 per thread, every load of the pass in flight before the first use (a peer
        // load is a full NVLink round trip)
        float4 g[kUnroll], pp[kUnroll], mm[kUnroll], vv[kUnroll];
        size_t e[kUnroll];
        bool on[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const size_t i = a + threadIdx.x + (size_t)j * kThreadsC;
            on[j] = i < b;
            e[j] = 4 * (on[j] ? i : a);
            g[j] = ld_peer(P.grad[0] + e[j]);
        }
        for (int k = 1; k < world; ++k) {
            float4 q[kUnroll];
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) q[j] = ld_peer(P.grad[k] + e[j]);
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) { g[j].x += q[j].x; g[j].y += q[j].y; g[j].z += q[j].z; g[j].w += q[j].w; }
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            pp[j] = *reinterpret_cast<const float4 *>(P.param[rank] + e[j]);
            mm[j] = *reinterpret_cast<const float4 *>(m + e[j]);
            vv[j] = *reinterpret_cast<const float4 *>(v + e[j]);
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            adam_one(pp[j].x, g[j].x, mm[j].x, vv[j].x, b1, b2, eps, step_size, bc2_sqrt, 1.f);
            adam_one(pp[j].y, g[j].y, mm[j].y, vv[j].y, b1, b2, eps, step_size, bc2_sqrt, 1.f);
            adam_one(pp[j].z, g[j].z, mm[j].z, vv[j].z, b1, b2, eps, step_size, bc2_sqrt, 1.f);
            adam_one(pp[j].w, g[j].w, mm[j].w, vv[j].w, b1, b2, eps, step_size, bc2_sqrt, 1.f);
            if (on[j]) {
                *reinterpret_cast<float4 *>(m + e[j]) = mm[j];
                *reinterpret_cast<float4 *>(v + e[j]) = vv[j];
                for (int k = 0; k < world; ++k) *reinterpret_cast<float4 *>(P.param[k] + e[j]) = pp[j];
            }
        }
    }
    __threadfence_system();
    peer_barrier(P, rank, world, slot);
}

}  // namespace
}  // namespace vcm

extern "C" {

int vcm_reduce_adam_chunk(void) { return (int)(4 * vcm::kChunk4); }

int vcm_reduce_adam_blocks(size_t n, int world) {
    if (world < 1) world = 1;
    int64_t blocks = vcm::ceil_div((int64_t)vcm::ceil_div((int64_t)(n / 4), (int64_t)vcm::kChunk4) + 1, world);
    // few CTAs: each thread keeps kUnroll x world 16-byte peer loads in flight (measured on 2 GPUs inside the step: 16 CTAs 2.68 ms, 32 2.54, 64 2.52, 120 2.54; stand-alone more is faster
    // - 597, 315, 115 us per 55.6 MB), and the CTAs wait at the first barrier next to the backward kernels they overlap with
    static const int cap = [] { const char *e = getenv("VCM_COLL_BLOCKS"); const int v = e ? atoi(e) : 64; return v < 1 ? 1 : (v > 120 ? 120 : v); }();
    if (blocks < 1) blocks = 1;
    if (blocks > cap) blocks = cap;
    return (int)blocks;
}

int vcm_reduce_adam_bcast(const float *const *grad_peers, float *const *param_peers, uint32_t *const *flag_peers,
                          float *exp_avg, float *exp_avg_sq, size_t offset, size_t n, const float *opt_state, float beta1,
                          float beta2, float eps, int rank, int world, int flag_slot, void *stream) {
    using namespace vcm;
    VCM_CHECK(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "rank %d / world %d (at most %d ranks)", rank,
              world, kMaxWorld);
    VCM_CHECK(grad_peers && param_peers && flag_peers, "peer pointer arrays are NULL");
    VCM_CHECK(n % 4 == 0 && offset % 4 == 0, "offset and length must be multiples of 4 elements");
    VCM_DEVPTR(exp_avg); VCM_DEVPTR(exp_avg_sq); VCM_DEVPTR(opt_state);
    if (n == 0) return VCM_OK;
    Peers P;
    for (int k = 0; k < kMaxWorld; ++k) {
        P.grad[k] = k < world ? grad_peers[k] : nullptr;
        P.param[k] = k < world ? param_peers[k] : nullptr;
        P.flag[k] = k < world ? flag_peers[k] : nullptr;
        if (k < world) VCM_CHECK(P.grad[k] && P.param[k] && P.flag[k], "peer %d: NULL buffer", k);
    }
    VCM_CHECK(((reinterpret_cast<uintptr_t>(P.grad[rank]) | reinterpret_cast<uintptr_t>(P.param[rank]) |
                reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0,
              "buffers must be 16-byte aligned");
    const int blocks = vcm_reduce_adam_blocks(n, world);
    reduce_adam_bcast_kernel<<<blocks, kThreadsC, 0, (cudaStream_t)stream>>>(P, exp_avg, exp_avg_sq, offset / 4,
                                                                            (offset + n) / 4, opt_state, beta1, beta2, eps,
                                                                            rank, world, flag_slot);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

}  // extern "C"
