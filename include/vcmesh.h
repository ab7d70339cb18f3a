/* libvcmesh.so — the vc-conv / vd-pool / vd-unpool hot path of VCMeshConv as
 * hand-written sm_100a CUDA kernels behind a plain C ABI.
 *
 * The reference has no FFI layer: its hot path is the ATen op sequence inside
 * LASMConvssw.forward (GraphAutoEncoder/graphVAESSW.py:97-163) and the autograd
 * mirror of it. Each entry point below replaces one slice of that sequence;
 * the Python layer in vcmeshconv_b200/ (same nn.Module API as the reference)
 * binds them with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer argument of a compute call is a DEVICE pointer to
 *     contiguous row-major fp32 (or int32 where stated) on the current device;
 *   - every compute call takes the cudaStream_t it must enqueue on (as
 *     void*), is asynchronous, allocates nothing and holds no global mutable
 *     state (CUDA-graph capturable); the caller owns all buffers;
 *   - return value: 0 = ok, otherwise a VCM_E* code below or a cudaError_t
 *     (>0); vcm_last_error() returns the message of the last failure on the
 *     calling thread. Nothing throws, nothing exits;
 *   - there is NO CPU fallback: host pointers are rejected.
 *
 * Shapes:  B batch, P_in / P_out vertex counts, N padded neighbour count,
 *          M number of bases, I / O channels, R = B*P_out rows, K = M*I.
 *   x [B,P_in,I]   A [P_out,N,M]   z [B,P_out,M,I] = [R,K]
 *   W [M,O,I] (the reference's `weights` [M, O*I], native layout)
 *   y / out [B,P_out,O]   bias [P_out,O] or [O]
 */
#ifndef VCMESH_H
#define VCMESH_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define VCM_API __attribute__((visibility("default")))
#else
#define VCM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

enum {
    VCM_OK = 0,
    VCM_EINVAL = -1,      /* bad argument (shape, null pointer, host pointer) */
    VCM_EUNSUPPORTED = -2,/* valid request this build has no kernel for */
    VCM_ENOMEM = -3
};

/* GEMM arithmetic for the basis GEMMs (vcm_basis_*):
 *   VCM_PREC_FP32   fp32 FMA on CUDA cores: the reference's own arithmetic
 *   VCM_PREC_TF32   tcgen05 kind::tf32, fp32 accumulate in TMEM (1 pass)
 *   VCM_PREC_TF32X3 tcgen05, 3-pass hi/lo split: fp32-grade result on tensor cores
 * Shapes the tensor-core kernels do not cover fall back to VCM_PREC_FP32
 * kernels (still CUDA, never the CPU). */
enum { VCM_PREC_FP32 = 0, VCM_PREC_TF32 = 1, VCM_PREC_TF32X3 = 2 };

VCM_API const char *vcm_last_error(void);
VCM_API int vcm_version(void);
/* Diagnostics: number of CUDA kernels this library has launched in the process. */
VCM_API unsigned long long vcm_launch_count(void);
/* 1 when a tcgen05 kernel exists for this (rows, K, O) and precision. */
VCM_API int vcm_basis_uses_tensor_cores(int64_t rows, int M, int I, int O, int precision);

/* ---------------------------------------------------------------- tables --
 * Immutable per-layer connectivity on the device. Replaces the parsing done in
 * LASMConvssw.__init__ (graphVAESSW.py:44-57) and the per-call sentinel mask
 * (graphVAESSW.py:111-118).
 *
 * table_host: HOST int32 [P_out, cols], cols = 1 + 2N, rows as written by
 * GraphSampling (meshCNN.h:54-92): [cnt, id0, hop0, id1, hop1, ...], padding
 * id == P_in. Derived device tables:
 *   idx      int32 [P_out, N]   neighbour ids (sentinel P_in kept as is)
 *   last     int32 [P_out]      1 + position of the last non-sentinel entry
 *   perm     int32 [P_out]      output vertices sorted by `last` descending
 *                               (stable) = degree-bucketed processing order
 *   idx_sorted  int32 [P_out, N]  rows of idx in `perm` order, meta_sorted int32 [P_out, 2] = {vertex,
 *                               last} in `perm` order: what a CTA needs for work item s sits at
 *                               addresses that depend on s only
 *   rev_ptr  int32 [P_in + 1]   CSR row pointers of the transposed graph
 *   rev_slot int32 [nnz]        for input vertex q the slots p*N+n with
 *                               idx[p,n] == q, ascending (deterministic dX)
 */
typedef struct vcm_tables vcm_tables;

typedef struct vcm_tables_info {
    int32_t p_in, p_out, n_max;
    int32_t max_last;   /* max over rows of `last` */
    int64_t nnz;        /* non-sentinel entries */
    int32_t max_rev;    /* max in-degree of the transposed graph */
    int32_t device;
} vcm_tables_info;

enum { VCM_TAB_IDX = 0, VCM_TAB_LAST = 1, VCM_TAB_PERM = 2, VCM_TAB_REV_PTR = 3, VCM_TAB_REV_SLOT = 4,
       VCM_TAB_IDX_SORTED = 5, VCM_TAB_META_SORTED = 6 };

VCM_API int vcm_tables_create(const int32_t *table_host, int p_out, int cols, int p_in, vcm_tables **out);
VCM_API void vcm_tables_destroy(vcm_tables *t);
VCM_API int vcm_tables_get_info(const vcm_tables *t, vcm_tables_info *info);
/* Copies one derived table back to the host (tests: bit-exactness). */
VCM_API int vcm_tables_export(const vcm_tables *t, int which, int32_t *dst_host, size_t capacity_elems);
/* Device pointer of a derived table (for callers that fuse their own kernels). */
VCM_API const int32_t *vcm_tables_device_ptr(const vcm_tables *t, int which);

/* -------------------------------------------------- gather / contract (K1) --
 * z[b,p,m,i] = sum_n [idx[p,n] != P_in] * A[p,n,m] * x[b, idx[p,n], i]
 * Replaces graphVAESSW.py:111-123 (mask, zero-pad, index_select, einsum
 * 'pnm,bpni->bpmi'). With M = 1 and A = normalised |p_neighbors| it is also
 * the residual pool/unpool of graphVAESSW.py:152-159. */
VCM_API int vcm_gather_contract_fwd(const vcm_tables *t, const float *x, const float *A, float *z,
                                    int B, int I, int M, void *stream);

/* Same contraction with the arithmetic mode of the basis GEMMs: VCM_PREC_TF32 runs it on the tensor pipe
 * (mma.sync m16n8k8 tf32, fp32 accumulate) when M <= 24 and I % 16 == 0; every other case is the fp32
 * kernel above. */
VCM_API int vcm_gather_contract_fwd_ex(const vcm_tables *t, const float *x, const float *A, float *z, int B,
                                       int I, int M, int precision, void *stream);

/* Project-first order for narrow outputs (M*O <= 64; the 32 -> 3 layer on the full mesh). With
 * u[b,q,m,o] = sum_i W[m,o,i] x[b,q,i] computed by vcm_basis_fwd (bases stacked along the output axis):
 *   fwd: s[b,p,o]    = sum_{n,m} A[p,n,m] u[b, idx[p,n], m, o]
 *   bwd: dA[p,n,m]   = sum_{b,o} ds[b,p,o] u[b, idx[p,n], m, o]      (0 on padding; NULL to skip)
 *        du[b,q,m,o] = sum_{(p,n): idx[p,n]=q} A[p,n,m] ds[b,p,o]    (reverse CSR; NULL to skip)
 * Same result as graphVAESSW.py:121-133 (the op is bilinear), without materialising z. */
VCM_API int vcm_project_gather_fwd(const vcm_tables *t, const float *u, const float *A, float *s, int B, int M,
                                   int O, void *stream);
VCM_API int vcm_project_gather_bwd(const vcm_tables *t, const float *ds, const float *u, const float *A, float *dA,
                                   float *du, int B, int M, int O, void *stream);

/* Backward of the above w.r.t. A and the gathered rows (K5):
 *   dA[p,n,m]   = sum_{b,i} dz[b,p,m,i] * x[b, idx[p,n], i]     (0 on padding)
 *   dG[b,p,n,i] = sum_m A[p,n,m] * dz[b,p,m,i]                  (real edges only)
 * dG is [B,P_out,N,I]; padding slots are left untouched and never read. dG may
 * be NULL when the input gradient is not needed (first layer), dA may be NULL
 * when the coefficients are constants (Laplacian / face gathers of the losses).
 * workspace: fp32, at least vcm_coeff_grad_workspace() elements (may be NULL
 * when that returns 0). Deterministic: fixed reduction order, no atomics. */
VCM_API size_t vcm_coeff_grad_workspace(const vcm_tables *t, int B, int I, int M);
VCM_API int vcm_coeff_grad(const vcm_tables *t, const float *dz, const float *x, const float *A, float *dA,
                           float *dG, float *workspace, int B, int I, int M, void *stream);

/* Same gradients with the arithmetic mode of the basis GEMMs: in VCM_PREC_TF32 / VCM_PREC_TF32X3 both are small
 * per-vertex GEMMs on the tensor pipe (mma.sync m16n8k8 tf32; TF32X3 = hi/lo split, fp32-grade) when M <= 24,
 * I % 16 == 0 and N <= 64; every other case is the fp32 kernel above. Use the matching workspace query. */
/* Diagnostics: which kernel vcm_coeff_grad_ex runs for this shape: 2 = tensor-pipe kernel with TMA-staged dz tiles
 * and an mbarrier pipeline (coeff_grad_tma_kernel, the TF32-mode default), 1 = tensor-pipe kernel staged with
 * cp.async (coeff_grad_mma_kernel), 0 = fp32 CUDA-core kernel (coeff_grad_kernel). bench.py groups its timings by it. */
VCM_API int vcm_coeff_grad_variant(const vcm_tables *t, int B, int I, int M, int precision);
/* Same, knowing which outputs are asked for: 3 = coeff_grad_small_kernel (dA only, I <= 8 channels and B*I <= 256
 * columns per vertex, tensor-core modes: one warp per vertex, mma.sync fragments loaded straight from global memory -
 * the 3-channel input layer and the M = 1 residual paths over xyz), otherwise the value above. */
VCM_API int vcm_coeff_grad_variant_ex(const vcm_tables *t, int B, int I, int M, int precision, int want_dA,
                                      int want_dG);
/* K5 with the final reduction taken off the backward chain. When the TMA kernel splits a vertex's units over several
 * CTAs (few vertices: the coarse levels), every CTA row leaves one partial slab of dA [P_out,N,M] and a small kernel
 * adds the slabs in fixed order. Only dG feeds the next kernel of the chain (K6); vcm_coeff_grad_slabs returns the
 * number of slabs (0 = this shape does not split or is not served by that kernel), vcm_coeff_grad_slabs_launch runs
 * K5 and leaves `slabs` [n_slabs][P_out*N*M] (the vcm_coeff_grad_workspace_ex size) unreduced, and vcm_sum_slabs -
 * launched by the caller on a side stream - adds them: out[i] = slabs[0][i] + slabs[1][i] + ... in that order. */
VCM_API int vcm_coeff_grad_slabs(const vcm_tables *t, int B, int I, int M, int precision);
VCM_API int vcm_coeff_grad_slabs_launch(const vcm_tables *t, const float *dz, const float *x, const float *A,
                                        float *slabs, float *dG, int B, int I, int M, int precision, void *stream);
VCM_API int vcm_sum_slabs(const float *slabs, float *out, size_t n, int n_slabs, void *stream);
VCM_API size_t vcm_coeff_grad_workspace_ex(const vcm_tables *t, int B, int I, int M, int precision);
VCM_API int vcm_coeff_grad_ex(const vcm_tables *t, const float *dz, const float *x, const float *A, float *dA,
                              float *dG, float *workspace, int B, int I, int M, int precision, void *stream);

/* dX by the reverse (CSR-transposed) neighbour list (K6), no atomics:
 *   dx[b,q,i] = sum_{slot in rev(q)} dG[b, slot, i]  (+ dx[b,q,i] if accumulate)
 * Replaces the index_add_ scatter autograd runs for graphVAESSW.py:122. */
VCM_API int vcm_scatter_rev(const vcm_tables *t, const float *dG, float *dx, int B, int I, int accumulate,
                            void *stream);
/* K6 with one coefficient per edge (M = 1 gathers: residual pool / un-pool graphVAESSW.py:151-159, loss gathers):
 * dx[b,q,:] (+)= sum_{slot in rev(q)} coef[slot] * dy[b, slot / N, :], coef [P_out, N] (padding slots are never
 * read), dy [B, P_out, I]. The autograd mirror of index_select + weighted sum without materialising the per-edge
 * gradient. */
VCM_API int vcm_scatter_rev_scaled(const vcm_tables *t, const float *dy, const float *coef, float *dx, int B, int I,
                                   int accumulate, void *stream);

/* Residual edge weights, graphVAESSW.py:155-157:
 *   pn[p,n] = |p_nb[p,n]| * mask / (sum_n |p_nb[p,n]| * mask + 1e-8)
 * and its backward (dp_nb from dpn). */
VCM_API int vcm_edge_weights_fwd(const vcm_tables *t, const float *p_nb, float *pn, void *stream);
VCM_API int vcm_edge_weights_bwd(const vcm_tables *t, const float *p_nb, const float *dpn, float *dp_nb,
                                 void *stream);

/* ------------------------------------------------------ basis GEMMs (K2-K4) --
 * K2 forward, replaces graphVAESSW.py:125-138 and the blend of :161:
 *   t[r,o]   = sum_{m,i} z[r,m,i] * W[m,o,i] + bias
 *   y        = act ? ELU(t) : t                (stored to y_act when non-NULL)
 *   out      = y * scale_y + (res ? res * scale_res : 0)
 * rows r = b*P_out + p; bias is [P_out,O] when bias_per_point, [O] otherwise,
 * NULL for none. out may alias nothing else. workspace: fp32, at least
 * vcm_basis_fwd_workspace() elements (partial sums of a split reduction when
 * there are few rows and a long K); may be NULL when that returns 0. */
VCM_API size_t vcm_basis_fwd_workspace(int64_t rows, int M, int I, int O, int precision);
VCM_API int vcm_basis_fwd(const float *z, const float *W, const float *bias, int bias_per_point,
                          const float *res, float *y_act, float *out, float *workspace, int B, int P_out, int M,
                          int I, int O, int act, float scale_y, float scale_res, int precision, void *stream);

/* K1 + K2 in ONE launch (graphVAESSW.py:111-138,161): the contraction z = sum_n A[p,n,m] x[b,idx[p,n],i] is
 * produced tile by tile in shared memory in the tensor-core operand layout and consumed there, so z makes no
 * HBM round trip. `z` (fp32 [B,P_out,M*I]) is optional: pass a buffer when the backward pass needs it (written
 * once by TMA stores), NULL for inference. Everything else as vcm_basis_fwd. Available for precision
 * VCM_PREC_TF32, B % 16 == 0, B <= 128, M <= 17, I % 8 == 0, O % 32 == 0 (vcm_conv_fwd_fused_supported);
 * other shapes use vcm_gather_contract_fwd + vcm_basis_fwd.
 * The 3-channel layers (I = 3, O = 32 or 64, M <= 24, N <= 64; VCM_PREC_TF32 and VCM_PREC_TF32X3) are served by a
 * second kernel with one warp per vertex and mma.sync fragments loaded straight from global memory
 * (fused_fwd_small_kernel): vcm_conv_fwd_fused_kind returns 2 for them, 1 for the tcgen05 kernel, 0 = unsupported.
 * The nn.Module layer uses kind 2 by default (the forward chain loses a launch and the round trip of z). */
VCM_API int vcm_conv_fwd_fused_kind(const vcm_tables *t, int B, int M, int I, int O, int precision);
VCM_API int vcm_conv_fwd_fused_supported(const vcm_tables *t, int B, int M, int I, int O, int precision);
VCM_API size_t vcm_conv_fwd_fused_workspace(const vcm_tables *t, int B, int M, int I, int O, int precision);
VCM_API int vcm_conv_fwd_fused(const vcm_tables *t, const float *x, const float *A, const float *W,
                               const float *bias, int bias_per_point, const float *res, float *y_act, float *out,
                               float *z, float *workspace, int B, int M, int I, int O, int act, float scale_y,
                               float scale_res, int precision, void *stream);

/* The K2 epilogue as a stand-alone pass (used when the basis projection runs
 * before the gather, i.e. for layers whose M*O is narrower than I):
 *   out = (act ? ELU : id)(s + bias) * scale_y + (res ? res * scale_res : 0)
 * Its backward is vcm_act_bwd. */
VCM_API int vcm_bias_act_fwd(const float *s, const float *bias, int bias_per_point, const float *res,
                             float *y_act, float *out, int B, int P_out, int O, int act, float scale_y,
                             float scale_res, void *stream);

/* Activation backward + bias gradient in one pass:
 *   ds[r,o]     = dout[r,o] * scale_y * (act ? (y>0 ? 1 : y+1) : 1)
 *   dbias[p,o]  = sum_b ds[b,p,o]          (dbias may be NULL)
 *   dres[r,o]   = dout[r,o] * scale_res    (dres may be NULL)
 * dbias is always the per-point [P_out,O] sum; for a shared bias reduce it
 * with vcm_colsum. */
VCM_API int vcm_act_bwd(const float *dout, const float *y_act, float *ds, float *dbias, float *dres, int B,
                        int P_out, int O, int act, float scale_y, float scale_res, void *stream);
VCM_API int vcm_colsum(const float *src, float *dst, int rows, int cols, void *stream);

/* K3: dz[r,m,i] = sum_o ds[r,o] * W[m,o,i]  (+ dz if accumulate). workspace: fp32,
 * at least vcm_basis_bwd_dz_workspace() elements (the tensor-core kernel stages
 * the stacked bases Wt[M*I, O] there); may be NULL when that returns 0. */
VCM_API size_t vcm_basis_bwd_dz_workspace(int64_t rows, int M, int I, int O, int precision);
VCM_API int vcm_basis_bwd_dz(const float *ds, const float *W, float *dz, float *workspace, int64_t rows, int M,
                             int I, int O, int accumulate, int precision, void *stream);

/* The stacked bases Wt[M*I, O] (Wt[m*I+i][o] = W[m][o][i]; the north_star's [M*in, out] matrix) depend on
 * the parameters only, so a training step can build them once, off the backward chain (a second stream / a
 * parallel graph branch), and hand them to K3: vcm_basis_bwd_dz_stacked is vcm_basis_bwd_dz without the
 * in-call stacking. Valid where vcm_basis_bwd_dz_workspace() != 0 (tensor-core shapes), VCM_EUNSUPPORTED else;
 * Wt has that many elements. */
VCM_API int vcm_basis_stack(const float *W, float *Wt, int M, int I, int O, void *stream);
VCM_API int vcm_basis_bwd_dz_stacked(const float *ds, const float *Wt, float *dz, int64_t rows, int M, int I,
                                     int O, int precision, void *stream);

/* K4: dW[m,o,i] = sum_r ds[r,o] * z[r,m,i]. workspace: fp32, at least
 * vcm_basis_bwd_dw_workspace() elements. Deterministic split reduction. */
VCM_API size_t vcm_basis_bwd_dw_workspace(int64_t rows, int M, int I, int O, int precision);
VCM_API int vcm_basis_bwd_dw(const float *ds, const float *z, float *dW, float *workspace, int64_t rows,
                             int M, int I, int O, int precision, void *stream);

/* ------------------------------------------------------------- optimizer --
 * Fused Adam over one flat fp32 buffer; replaces torch.optim.Adam(lr,
 * betas=(0.9, 0.999)) of graphVAE_train.py:252-262 / :316 (no weight decay, no
 * amsgrad). `step` counts from 1. g is multiplied by grad_scale first (1 for
 * the reference's SUM-over-replicas semantics). */
VCM_API int vcm_adam_step(float *p, const float *g, float *m, float *v, size_t n, float lr, float beta1,
                          float beta2, float eps, int step, float grad_scale, void *stream);

/* Same update with the step counter and learning rate resident on the device, so
 * the call can be captured once in a CUDA graph and replayed: state is 4 floats
 * {step, lr, 1-beta1^step, sqrt(1-beta2^step)}; the call increments step first.
 * The host changes the learning rate (ExponentialLR, graphVAE_train.py:281,353)
 * by writing state[1]. */
VCM_API int vcm_adam_step_graph(float *p, const float *g, float *m, float *v, size_t n, float *state, float beta1,
                                float beta2, float eps, float grad_scale, void *stream);
/* The two halves of vcm_adam_step_graph, so that one step can update the flat buffer range by range (a range
 * whose gradients are final is updated while the rest of the backward pass still runs): advance the state once
 * per step, then apply it to any number of disjoint ranges. */
VCM_API int vcm_adam_advance(float *state, float beta1, float beta2, void *stream);
VCM_API int vcm_adam_apply(float *p, const float *g, float *m, float *v, size_t n, const float *state, float beta1,
                           float beta2, float eps, float grad_scale, void *stream);

/* ------------------------------------------------------------- loss head --
 * The autoencoder's loss glue (GraphAutoEncoder/graphVAE_train.py:199-225) as four
 * launches instead of ~110 few-microsecond ATen kernels between the decoder's last
 * kernel and the first backward kernel. Deterministic (fixed-order
 * partial sums, reverse-CSR gathers, no atomics).
 *
 * Reparameterisation + KL term (graphVAE_train.py:199-206; the two encoder heads are
 * scaled by scale_mu / scale_logstd as MCEnc.forward does, graphVAESSW.py:286-288):
 *   mu = raw_mu*scale_mu, ls = raw_logstd*scale_logstd, z = mu + exp(ls)*eps,
 *   kl[0] = mean(-0.5 - ls + 0.5 mu^2 + 0.5 exp(2 ls)).  All arrays have n elements.
 * workspace: fp32, vcm_reparam_kl_workspace(n) elements (NULL when that is 0).
 * Backward: dz [n] and dkl [1] (device scalars; either may be NULL = zero). */
VCM_API size_t vcm_reparam_kl_workspace(size_t n);
VCM_API int vcm_reparam_kl_fwd(const float *raw_mu, const float *raw_logstd, const float *eps, float *z, float *kl,
                               float *workspace, size_t n, float scale_mu, float scale_logstd, void *stream);
VCM_API int vcm_reparam_kl_bwd(const float *raw_mu, const float *raw_logstd, const float *eps, const float *dz,
                               const float *dkl, float *d_raw_mu, float *d_raw_logstd, size_t n, float scale_mu,
                               float scale_logstd, void *stream);

/* Geometric losses on out / gt [B,P,3] and face normals nor [B,F,3]:
 *   l1  = mean |gt - out|                                  graphVAESSW.py:416-421
 *   lap = mean_{b,p} |L(gt - out)[b,p]|^2, L(e)[p] = sum_n lap_coef[p,n] e[idx[p,n]]
 *         over the layer-0 table `lap` (coef = cnt_p for slot 0, -1 elsewhere)
 *                                                          graphVAESSW.py:507-542
 *   nor = mean_{b,f} ((sum_k (out - gt)[face[f,k]]) / 3 . nor[b,f])^2 over the
 *         one-row-per-face table `faces` ([F, 1+2*3])      graphVAE_train.py:212-217
 *   total[0] = w_pose*l1 + w_lap*lap + w_nor*nor  (:225 without KL), parts[3] = {l1, lap, nor}
 * The forward also stores lapv [B,P,3] = L(gt - out) and s [B,F] (the per-face dot
 * products) for the backward, which writes d_out = g_total[0] * d(total)/d(out).
 * workspace: fp32, vcm_loss_head_workspace() elements. */
VCM_API size_t vcm_loss_head_workspace(const vcm_tables *lap, const vcm_tables *faces, int B);
VCM_API int vcm_loss_head_fwd(const vcm_tables *lap, const vcm_tables *faces, const float *lap_coef, const float *out,
                              const float *gt, const float *nor, float *lapv, float *s, float *total, float *parts,
                              float *workspace, int B, float w_pose, float w_lap, float w_nor, void *stream);
VCM_API int vcm_loss_head_bwd(const vcm_tables *lap, const vcm_tables *faces, const float *lap_coef, const float *out,
                              const float *gt, const float *nor, const float *lapv, const float *s,
                              const float *g_total, float *d_out, int B, float w_pose, float w_lap, float w_nor,
                              void *stream);

/* ------------------------------------------------- data-parallel reduction --
 * Gradient SUM over the ranks of one node fused with the optimizer, over NVLink peer memory (replaces the
 * reduce-to-GPU-0 + Adam of nn.DataParallel, graphVAE_train.py:277, 315-316). grad_peers / param_peers / flag_peers
 * are HOST arrays of `world` DEVICE pointers: the base of every rank's flat gradient, flat parameter and flag buffer
 * as mapped into THIS process (symmetric memory). For the range [offset, offset + n) one kernel does: barrier; for
 * the chunks this rank owns (chunk c = elements [c*chunk, (c+1)*chunk) of the flat buffers belongs to rank
 * c % world) sum the ranks' gradients in rank order, apply Adam (exp_avg / exp_avg_sq are this rank's own buffers,
 * live on its chunks only; opt_state = {step, lr, 1-b1^t, sqrt(1-b2^t)} on the device, see vcm_adam_advance) and
 * store the new parameters into every rank's parameter buffer; barrier. Flag buffer: uint32, zero-initialised, at
 * least 8 * (flag_slot + vcm_reduce_adam_blocks(n, world)) entries; concurrent launches must use disjoint slots and
 * every rank must issue the same launches in the same order. */
VCM_API int vcm_reduce_adam_chunk(void);
VCM_API int vcm_reduce_adam_blocks(size_t n, int world);
VCM_API int vcm_reduce_adam_bcast(const float *const *grad_peers, float *const *param_peers, uint32_t *const *flag_peers,
                                  float *exp_avg, float *exp_avg_sq, size_t offset, size_t n, const float *opt_state,
                                  float beta1, float beta2, float eps, int rank, int world, int flag_slot, void *stream);

#ifdef __cplusplus
}
#endif
#endif
