// Loss head of the autoencoder step as four launches instead of ~110 tiny ATen kernels:
//   vcm_reparam_kl_fwd / _bwd   reparameterisation + KL term   graphVAE_train.py:199-206
//   vcm_loss_head_fwd / _bwd    L1 pose + graph-Laplacian L2 + face-normal losses and their
//                               weighted sum                    graphVAE_train.py:212-225,
//                               graphVAESSW.py:416-421 (l1), :507-542 (laplace l2)
// At 6 890 vertices the step is latency-bound here: the decoder's last kernel and the first backward kernel
// are separated by a chain of dependent few-microsecond launches (200 us of a 2.8 ms step). Everything is
// deterministic: per-CTA partial sums are combined in a fixed order, the backward gathers through the reverse
// CSR of the two tables (ascending slots), no atomics.
#include <algorithm>

#include "common.cuh"

namespace vcm {

constexpr int kLossThreads = 256;
constexpr int kLossMaxParts = 2048;          // partial-sum slots a forward launch may use

// fixed-order block reduction of NV values per thread; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float *sm /* [NV][32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], s);
        if (lane == 0) sm[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float t = lane < nw ? sm[k * 32 + lane] : 0.f;
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) t += __shfl_xor_sync(0xffffffffu, t, s);
            v[k] = t;
        }
    }
}

// ------------------------------------------------------------------ reparameterisation + KL ------
// mu = raw_mu * s_mu, ls = raw_ls * s_ls (MCEnc scales its two heads, graphVAESSW.py:286-288)
// z = mu + exp(ls) * eps;  kl = mean(-0.5 - ls + 0.5 mu^2 + 0.5 exp(2 ls))
__global__ void __launch_bounds__(1024)
reparam_kl_fwd_kernel(const float *__restrict__ raw_mu, const float *__restrict__ raw_ls, const float *__restrict__ eps,
                      float *__restrict__ z, float *__restrict__ part, float *__restrict__ kl, size_t n, float s_mu,
                      float s_ls) {
    __shared__ float sm[32];
    float acc[1] = {0.f};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float mu = __ldg(raw_mu + i) * s_mu, ls = __ldg(raw_ls + i) * s_ls;
        z[i] = mu + expf(ls) * __ldg(eps + i);
        acc[0] += -0.5f - ls + 0.5f * (mu * mu) + 0.5f * expf(2.f * ls);
    }
    block_sum<1>(acc, sm);
    if (threadIdx.x == 0) {
        if (gridDim.x == 1) kl[0] = acc[0] / (float)n;
        else part[blockIdx.x] = acc[0];
    }
}

__global__ void reparam_kl_final_kernel(const float *__restrict__ part, int n_part, float *__restrict__ kl, size_t n) {
    __shared__ float sm[32];
    float acc[1] = {0.f};
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) acc[0] += part[i];
    block_sum<1>(acc, sm);
    if (threadIdx.x == 0) kl[0] = acc[0] / (float)n;
}

// d_raw_mu = s_mu (dz + dkl mu / n);  d_raw_ls = s_ls (dz eps exp(ls) + dkl (exp(2 ls) - 1) / n)
__global__ void reparam_kl_bwd_kernel(const float *__restrict__ raw_mu, const float *__restrict__ raw_ls,
                                      const float *__restrict__ eps, const float *__restrict__ dz,
                                      const float *__restrict__ dkl, float *__restrict__ d_raw_mu,
                                      float *__restrict__ d_raw_ls, size_t n, float s_mu, float s_ls) {
    const float gk = dkl != nullptr ? __ldg(dkl) / (float)n : 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float mu = __ldg(raw_mu + i) * s_mu, ls = __ldg(raw_ls + i) * s_ls;
        const float g = dz != nullptr ? __ldg(dz + i) : 0.f;
        d_raw_mu[i] = s_mu * (g + gk * mu);
        d_raw_ls[i] = s_ls * (g * __ldg(eps + i) * expf(ls) + gk * (expf(2.f * ls) - 1.f));
    }
}

// ------------------------------------------------------------------ loss head, forward ------
// One thread per (b, vertex) and per (b, face). e = gt - out.
//   vertex:  l1 += |e|;  lv = sum_n coef[p,n] e[b, idx[p,n]] (stored: the backward needs it);  lap += |lv|^2
//   face:    s = ((sum_k -e[b, face[f,k]]) / 3 . nor[b,f]) (stored);  nor += s^2
__global__ void __launch_bounds__(kLossThreads)
loss_head_fwd_kernel(const int32_t *__restrict__ lap_idx, const int32_t *__restrict__ lap_last,
                     const float *__restrict__ lap_coef, const int32_t *__restrict__ face_idx,
                     const float *__restrict__ out, const float *__restrict__ gt, const float *__restrict__ nor,
                     float *__restrict__ lapv, float *__restrict__ s_out, float *__restrict__ part, int P, int n_lap,
                     int F, int B) {
    __shared__ float sm[3 * 32];
    const int64_t n_v = (int64_t)B * P, n_f = (int64_t)B * F;
    float acc[3] = {0.f, 0.f, 0.f};                      // l1, lap, normal
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_v + n_f; w += (int64_t)gridDim.x * blockDim.x) {
        if (w < n_v) {
            const int b = (int)(w / P), p = (int)(w - (int64_t)b * P);
            const float *o = out + (size_t)b * P * 3, *g = gt + (size_t)b * P * 3;
            acc[0] += fabsf(g[p * 3] - o[p * 3]) + fabsf(g[p * 3 + 1] - o[p * 3 + 1]) + fabsf(g[p * 3 + 2] - o[p * 3 + 2]);
            const int l = __ldg(lap_last + p);
            float lx = 0.f, ly = 0.f, lz = 0.f;
            for (int n = 0; n < l; ++n) {
                const int q = __ldg(lap_idx + (size_t)p * n_lap + n);
                if (q == P) continue;
                const float c = __ldg(lap_coef + (size_t)p * n_lap + n);
                lx = fmaf(c, g[q * 3] - o[q * 3], lx);
                ly = fmaf(c, g[q * 3 + 1] - o[q * 3 + 1], ly);
                lz = fmaf(c, g[q * 3 + 2] - o[q * 3 + 2], lz);
            }
            float *d = lapv + (size_t)w * 3;
            d[0] = lx; d[1] = ly; d[2] = lz;
            acc[1] += lx * lx + ly * ly + lz * lz;
        } else {
            const int64_t wf = w - n_v;
            const int b = (int)(wf / F), f = (int)(wf - (int64_t)b * F);
            const float *o = out + (size_t)b * P * 3, *g = gt + (size_t)b * P * 3;
            float tx = 0.f, ty = 0.f, tz = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int q = __ldg(face_idx + (size_t)f * 3 + k);
                if (q == P) continue;
                tx += o[q * 3] - g[q * 3];
                ty += o[q * 3 + 1] - g[q * 3 + 1];
                tz += o[q * 3 + 2] - g[q * 3 + 2];
            }
            const float *nn = nor + (size_t)wf * 3;
            const float s = tx / 3.0f * __ldg(nn) + ty / 3.0f * __ldg(nn + 1) + tz / 3.0f * __ldg(nn + 2);
            s_out[wf] = s;
            acc[2] += s * s;
        }
    }
    block_sum<3>(acc, sm);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 3] = acc[0];
        part[blockIdx.x * 3 + 1] = acc[1];
        part[blockIdx.x * 3 + 2] = acc[2];
    }
}

// total = w_pose l1 + w_lap lap + w_nor nor, parts = {l1, lap, nor}
__global__ void loss_head_final_kernel(const float *__restrict__ part, int n_part, float *__restrict__ total,
                                       float *__restrict__ parts, float inv_l1,
                                       float inv_lap, float inv_nor, float w_pose, float w_lap, float w_nor) {
    __shared__ float sm[3 * 32];
    float acc[3] = {0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) {
        acc[0] += part[i * 3];
        acc[1] += part[i * 3 + 1];
        acc[2] += part[i * 3 + 2];
    }
    block_sum<3>(acc, sm);
    if (threadIdx.x == 0) {
        const float l1 = acc[0] * inv_l1, lap = acc[1] * inv_lap, nr = acc[2] * inv_nor;
        total[0] = w_pose * l1 + w_lap * lap + w_nor * nr;
        parts[0] = l1;
        parts[1] = lap;
        parts[2] = nr;
    }
}

// ------------------------------------------------------------------ loss head, backward ------
// d_out[b,q] = g * ( -w_pose sign(e) / (3BP)
//                    - 2 w_lap / (BP) * sum_{(p,n): idx[p,n]=q} coef[p,n] lv[b,p]
//                    + 2 w_nor / (3BF) * sum_{f contains q} s[b,f] nor[b,f] )
__global__ void __launch_bounds__(kLossThreads)
loss_head_bwd_kernel(const int32_t *__restrict__ lap_rev_ptr, const int32_t *__restrict__ lap_rev_slot,
                     const float *__restrict__ lap_coef, const int32_t *__restrict__ face_rev_ptr,
                     const int32_t *__restrict__ face_rev_slot, const float *__restrict__ out,
                     const float *__restrict__ gt, const float *__restrict__ nor, const float *__restrict__ lapv,
                     const float *__restrict__ s_in, const float *__restrict__ g_total, float *__restrict__ d_out, int P,
                     int n_lap, int F, int B, float c_l1, float c_lap, float c_nor) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (int64_t)B * P) return;
    const int b = (int)(w / P), q = (int)(w - (int64_t)b * P);
    const float gs = __ldg(g_total);
    const float *o = out + (size_t)w * 3, *g = gt + (size_t)w * 3;
    float r[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float e = g[c] - o[c];
        r[c] = -c_l1 * (float)((e > 0.f) - (e < 0.f));
    }
    {
        const float *lv = lapv + (size_t)b * P * 3;
        float ax = 0.f, ay = 0.f, az = 0.f;
        for (int e = __ldg(lap_rev_ptr + q), e1 = __ldg(lap_rev_ptr + q + 1); e < e1; ++e) {
            const int slot = __ldg(lap_rev_slot + e), p = slot / n_lap;
            const float c = __ldg(lap_coef + slot);
            ax = fmaf(c, lv[p * 3], ax);
            ay = fmaf(c, lv[p * 3 + 1], ay);
            az = fmaf(c, lv[p * 3 + 2], az);
        }
        r[0] -= c_lap * ax; r[1] -= c_lap * ay; r[2] -= c_lap * az;
    }
    {
        const float *sb = s_in + (size_t)b * F, *nb = nor + (size_t)b * F * 3;
        float ax = 0.f, ay = 0.f, az = 0.f;
        for (int e = __ldg(face_rev_ptr + q), e1 = __ldg(face_rev_ptr + q + 1); e < e1; ++e) {
            const int f = __ldg(face_rev_slot + e) / 3;
            const float s = sb[f];
            ax = fmaf(s, __ldg(nb + f * 3), ax);
            ay = fmaf(s, __ldg(nb + f * 3 + 1), ay);
            az = fmaf(s, __ldg(nb + f * 3 + 2), az);
        }
        r[0] += c_nor * ax; r[1] += c_nor * ay; r[2] += c_nor * az;
    }
    float *d = d_out + (size_t)w * 3;
    d[0] = gs * r[0]; d[1] = gs * r[1]; d[2] = gs * r[2];
}

static int loss_grid(int64_t items) {
    const int64_t want = ceil_div(items, kLossThreads);
    const int64_t cap = std::min<int64_t>(kLossMaxParts, (int64_t)sm_count() * 8);
    return (int)std::max<int64_t>(1, std::min(want, cap));
}

}  // namespace vcm

using namespace vcm;

extern "C" {

// one 1024-thread CTA per 4096 elements (a single CTA walking a 50k-element latent takes 26 us), at most one per SM
static int reparam_grid(size_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div((int64_t)n, 4096), sm_count())); }

size_t vcm_reparam_kl_workspace(size_t n) { return reparam_grid(n) > 1 ? (size_t)reparam_grid(n) : 0; }

int vcm_reparam_kl_fwd(const float *raw_mu, const float *raw_logstd, const float *eps, float *z, float *kl,
                       float *workspace, size_t n, float scale_mu, float scale_logstd, void *stream) {
    VCM_CHECK(n > 0, "empty latent");
    VCM_DEVPTR(raw_mu); VCM_DEVPTR(raw_logstd); VCM_DEVPTR(eps); VCM_DEVPTR(z); VCM_DEVPTR(kl);
    VCM_DEVPTR_OPT(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = reparam_grid(n);
    VCM_CHECK(grid == 1 || workspace != nullptr, "workspace required (%d partial sums)", grid);
    reparam_kl_fwd_kernel<<<grid, 1024, 0, st>>>(raw_mu, raw_logstd, eps, z, workspace, kl, n, scale_mu, scale_logstd);
    VCM_LAUNCH_CHECK();
    if (grid > 1) {
        reparam_kl_final_kernel<<<1, 256, 0, st>>>(workspace, grid, kl, n);
        VCM_LAUNCH_CHECK();
    }
    return VCM_OK;
}

int vcm_reparam_kl_bwd(const float *raw_mu, const float *raw_logstd, const float *eps, const float *dz,
                       const float *dkl, float *d_raw_mu, float *d_raw_logstd, size_t n, float scale_mu,
                       float scale_logstd, void *stream) {
    VCM_CHECK(n > 0, "empty latent");
    VCM_DEVPTR(raw_mu); VCM_DEVPTR(raw_logstd); VCM_DEVPTR(eps); VCM_DEVPTR(d_raw_mu); VCM_DEVPTR(d_raw_logstd);
    VCM_DEVPTR_OPT(dz); VCM_DEVPTR_OPT(dkl);
    const int grid = (int)std::min<int64_t>(ceil_div((int64_t)n, 256), (int64_t)sm_count() * 8);
    reparam_kl_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(raw_mu, raw_logstd, eps, dz, dkl, d_raw_mu,
                                                                   d_raw_logstd, n, scale_mu, scale_logstd);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

static int check_loss_tables(const Tables *lap, const Tables *face) {
    VCM_CHECK(lap != nullptr && face != nullptr, "tables handle is NULL");
    VCM_CHECK(lap->p_in == lap->p_out, "the Laplacian table must map the mesh onto itself");
    VCM_CHECK(face->p_in == lap->p_in && face->n_max == 3, "the face table must be [F, 1 + 2*3] over the same %d vertices",
              lap->p_in);
    return VCM_OK;
}

size_t vcm_loss_head_workspace(const vcm_tables *lap_h, const vcm_tables *face_h, int B) {
    const Tables *lap = as_tables(lap_h), *face = as_tables(face_h);
    if (lap == nullptr || face == nullptr || B <= 0) return 0;
    return (size_t)3 * loss_grid((int64_t)B * (lap->p_out + face->p_out));
}

int vcm_loss_head_fwd(const vcm_tables *lap_h, const vcm_tables *face_h, const float *lap_coef, const float *out,
                      const float *gt, const float *nor, float *lapv, float *s, float *total, float *parts,
                      float *workspace, int B, float w_pose, float w_lap, float w_nor, void *stream) {
    const Tables *lap = as_tables(lap_h), *face = as_tables(face_h);
    if (int e = check_loss_tables(lap, face)) return e;
    VCM_CHECK(B > 0 && lap->p_out > 0 && face->p_out > 0, "empty batch or mesh");
    VCM_DEVPTR(lap_coef); VCM_DEVPTR(out); VCM_DEVPTR(gt); VCM_DEVPTR(nor); VCM_DEVPTR(lapv); VCM_DEVPTR(s);
    VCM_DEVPTR(total); VCM_DEVPTR(parts); VCM_DEVPTR(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    const int P = lap->p_out, F = face->p_out;
    const int grid = loss_grid((int64_t)B * (P + F));
    loss_head_fwd_kernel<<<grid, kLossThreads, 0, st>>>(lap->idx, lap->last, lap_coef, face->idx, out, gt, nor, lapv, s,
                                                        workspace, P, lap->n_max, F, B);
    VCM_LAUNCH_CHECK();
    loss_head_final_kernel<<<1, 256, 0, st>>>(workspace, grid, total, parts, 1.f / ((float)B * P * 3.f), 1.f / ((float)B * P),
                                              1.f / ((float)B * F), w_pose, w_lap, w_nor);
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

int vcm_loss_head_bwd(const vcm_tables *lap_h, const vcm_tables *face_h, const float *lap_coef, const float *out,
                      const float *gt, const float *nor, const float *lapv, const float *s, const float *g_total,
                      float *d_out, int B, float w_pose, float w_lap, float w_nor, void *stream) {
    const Tables *lap = as_tables(lap_h), *face = as_tables(face_h);
    if (int e = check_loss_tables(lap, face)) return e;
    VCM_CHECK(B > 0 && lap->p_out > 0 && face->p_out > 0, "empty batch or mesh");
    VCM_DEVPTR(lap_coef); VCM_DEVPTR(out); VCM_DEVPTR(gt); VCM_DEVPTR(nor); VCM_DEVPTR(lapv); VCM_DEVPTR(s);
    VCM_DEVPTR(g_total); VCM_DEVPTR(d_out);
    const int P = lap->p_out, F = face->p_out;
    const float bp = (float)B * P, bf = (float)B * F;
    loss_head_bwd_kernel<<<(unsigned)ceil_div((int64_t)B * P, kLossThreads), kLossThreads, 0, (cudaStream_t)stream>>>(
        lap->rev_ptr, lap->rev_slot, lap_coef, face->rev_ptr, face->rev_slot, out, gt, nor, lapv, s, g_total, d_out, P,
        lap->n_max, F, B, w_pose / (3.f * bp), 2.f * w_lap / bp, 2.f * w_nor / (3.f * bf));
    VCM_LAUNCH_CHECK();
    return VCM_OK;
}

}  // extern "C"
