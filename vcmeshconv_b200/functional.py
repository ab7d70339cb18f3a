"""torch.autograd.Functions over the C ABI of libvcmesh.so.

PyTorch is plumbing here: it owns the device buffers, the stream and the
autograd tape. Every arithmetic step of the layer is one of our kernels:

  gather_contract   K1 fwd / K5 + K6 bwd       graphVAESSW.py:111-123
  basis_gemm        K2 fwd / act_bwd+K3+K4 bwd graphVAESSW.py:125-138, :161
  edge_weights      residual |p| normalisation graphVAESSW.py:155-157
  reparam_kl        z = mu + std*eps, KL term    graphVAE_train.py:199-206
  loss_head         l1 + laplace + normal losses graphVAE_train.py:212-225
  vcconv            the whole layer            graphVAESSW.py:97-163

Tensors must be CUDA fp32; CPU tensors raise (no fallback).
"""
import math

import numpy as np
import torch

from . import _lib

# Default: fp32-grade results on the tensor cores (3-pass TF32 hi/lo split, <= 1e-5 vs the reference; shapes the
# tcgen05 kernels do not take run the fp32 CUDA-core kernels). 'tf32' is the fast mode (stated tolerance 3e-3),
# 'fp32' forces CUDA-core FMA everywhere.
_precision = _lib.PRECISIONS[__import__("os").environ.get("VCM_PRECISION", "tf32x3")]


def set_precision(name):
    """'tf32x3' (default: tcgen05 with hi/lo split, fp32-grade), 'tf32' (tcgen05, single pass) or 'fp32'
    (CUDA-core FMA, the reference's own arithmetic); see include/vcmesh.h. Returns the previous mode."""
    global _precision
    prev = {v: k for k, v in _lib.PRECISIONS.items()}[_precision]
    _precision = _lib.PRECISIONS[name]
    return prev


def get_precision():
    return {v: k for k, v in _lib.PRECISIONS.items()}[_precision]


def _req(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: vcmeshconv_b200 has no CPU path" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32, got %s" % (name, t.dtype))
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the current device's stream and the tables live on the tensor's device
        raise RuntimeError("%s is on cuda:%d but the current device is cuda:%d (wrap the call in "
                           "torch.cuda.device(...))" % (name, t.device.index, torch.cuda.current_device()))
    return t.contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _grad_target(t, needed=True):
    """Slice of the trainer's flat gradient buffer registered for parameter `t` (or for the parameter `t`
    is a whole, order-preserving view of), if this is the first gradient written to it this step. The
    backward kernels then write there directly and autograd sees no gradient for the input: no temporary,
    no accumulate kernel, and the buffer NCCL reduces / Adam reads is filled in place. `needed` is the
    Function's `ctx.needs_input_grad[k]` of that input: False under no_grad at the call site (grad mode is always off
    INSIDE Function.forward, so it cannot be asked there), and then nothing is claimed."""
    if t is None or not needed:
        return None
    base = t._base if t._base is not None else t
    buf = getattr(base, "_vcm_grad_buf", None)
    if buf is None or not t.is_contiguous() or t.numel() != base.numel() or getattr(base, "_vcm_grad_written", False):
        return None
    base._vcm_grad_written = True
    return buf.view(t.shape)


# M = 1 gathers (residual pool / un-pool, loss gathers): dx without the per-edge gradient (VCM_M1_DIRECT=0: K5 + K6)
_m1_direct = __import__("os").environ.get("VCM_M1_DIRECT", "1") == "1"


def begin_step(params):
    """Called by the trainer before every (captured or eager) backward: each parameter may be written
    directly once; a second use falls back to autograd accumulation."""
    for p in params:
        p._vcm_grad_written = False


class _GatherContract(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, A, tables):
        x, A = _req(x, "x"), _req(A, "coefficients")
        B, p_in, I = x.shape
        if p_in != tables.p_in:
            raise RuntimeError("x has %d vertices, table expects %d" % (p_in, tables.p_in))
        if A.dim() != 3 or A.shape[0] != tables.p_out or A.shape[1] != tables.n_max:
            raise RuntimeError("coefficients must be [%d, %d, M], got %s" % (tables.p_out, tables.n_max, tuple(A.shape)))
        M = A.shape[2]
        z = torch.empty(B, tables.p_out, M * I, device=x.device, dtype=torch.float32)
        if B:
            nnz = tables.nnz
            _lib.call("vcm_gather_contract_fwd_ex", tables.handle, _ptr(x), _ptr(A), _ptr(z), B, I, M, _precision,
                      _lib.stream_ptr(),
                      nbytes=4 * (x.numel() + nnz * M + nnz + z.numel()), flops=2 * B * nnz * M * I,
                      tag="P%d>%d I%d M%d" % (p_in, tables.p_out, I, M))
        ctx.save_for_backward(x, A)
        ctx.tables = tables
        ctx.precision = _precision
        ctx.dA_target = _grad_target(A, ctx.needs_input_grad[1])
        return z

    @staticmethod
    def backward(ctx, dz):
        x, A = ctx.saved_tensors
        need_dx, need_dA = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx, dA = _gather_backward(ctx.tables, x, A, dz, need_dx, need_dA, ctx.dA_target, ctx.precision)
        return dx, dA, None


_defer_slab_sum = __import__("os").environ.get("VCM_DEFER_SLAB_SUM", "1") == "1"
# VCM_SPLIT_K5=1 (trainer): the dA half of K5 as a parallel branch, for layers whose dz is at most this many MB (the
# split reads dz twice: it can only pay where K5 is latency-bound, not on the layers that carry the bytes)
_split_k5_max_bytes = int(float(__import__("os").environ.get("VCM_SPLIT_K5_MAX_MB", "1e9")) * (1 << 20))


def _gather_backward(t, x, A, dz, need_dx, need_dA, dA_target, precision):
    """K5 (+ K6): (dx, dA) of z = gather_contract(x, A); dA is None when it was written into `dA_target`."""
    B, p_in, I = x.shape
    M = A.shape[2]
    if not (need_dx or need_dA) or B == 0:
        return (torch.zeros_like(x) if need_dx else None), (torch.zeros_like(A) if need_dA else None)
    dz = _req(dz, "dz")
    L = _lib.lib()
    if M == 1 and need_dx and _m1_direct:
        # one coefficient per edge: dx comes straight from dz through the reverse list (vcm_scatter_rev_scaled), the
        # per-edge gradient dG = A * dz is never written; K5 only produces dA
        st = _lib.stream_ptr()
        nnz = t.nnz
        tag = "P%d>%d I%d M1 direct" % (p_in, t.p_out, I)
        dx = torch.empty_like(x)
        _lib.call("vcm_scatter_rev_scaled", t.handle, _ptr(dz), _ptr(A), _ptr(dx), B, I, 0, st,
                  nbytes=4 * (dz.numel() + 2 * nnz + (p_in + 1) + x.numel()), flops=2 * B * nnz * I, tag=tag)
        if not need_dA:
            return dx, None
        _, dA = _gather_backward(t, x, A, dz, False, True, dA_target, precision)
        return dx, dA
    direct_dA = need_dA and dA_target is not None
    dA = (dA_target if direct_dA else torch.empty_like(A)) if need_dA else None
    dG = torch.empty(B, t.p_out, t.n_max, I, device=x.device, dtype=torch.float32) if need_dx else None
    n_ws = L.vcm_coeff_grad_workspace_ex(t.handle, B, I, M, precision) if need_dA else 0
    split = need_dx and direct_dA and _da_stream is not None and 4 * dz.numel() <= _split_k5_max_bytes
    ws = torch.empty(n_ws, device=x.device, dtype=torch.float32) if n_ws and not split else None
    st = _lib.stream_ptr()
    nnz = t.nnz
    # the tag names the layer shape and the kernel that serves it (bench.py groups timings by kernel)
    tag = "P%d>%d I%d M%d %s" % (p_in, t.p_out, I, M,
                                 ("fp32", "mma", "tma", "small")[L.vcm_coeff_grad_variant_ex(
                                     t.handle, B, I, M, precision, int(need_dA), int(need_dx))])
    gb = 4 * B * nnz * I
    if split:
        # Only dG is on the backward chain (dG -> K6 -> dx -> next layer); dA goes straight into the flat gradient
        # buffer and nothing downstream reads it. Two launches of K5 (compile-time dG-only / dA-only variants): the
        # chain waits for the cheaper half, the dA half runs as a parallel branch next to it (dz is read twice).
        _lib.call("vcm_coeff_grad_ex", t.handle, _ptr(dz), _ptr(x), _ptr(A), None, _ptr(dG), None, B, I, M,
                  precision, st, nbytes=4 * (dz.numel() + nnz * M + nnz) + gb, flops=2 * B * nnz * M * I, tag=tag)
        cur = torch.cuda.current_stream()
        _da_stream.wait_stream(cur)
        with torch.cuda.stream(_da_stream):
            ws = torch.empty(n_ws, device=x.device, dtype=torch.float32) if n_ws else None
            _lib.call("vcm_coeff_grad_ex", t.handle, _ptr(dz), _ptr(x), _ptr(A), _ptr(dA), None, _ptr(ws), B, I, M,
                      precision, _lib.stream_ptr(), nbytes=4 * (dz.numel() + x.numel() + 2 * nnz * M + nnz),
                      flops=2 * B * nnz * M * I, tag=tag)
        _side_keep.append((dz, x, ws))
    elif (need_dx and direct_dA and _side_stream is not None and _defer_slab_sum
          and L.vcm_coeff_grad_slabs(t.handle, B, I, M, precision) > 1):
        # few vertices: K5 leaves one partial slab of dA per CTA row. Only dG continues down the chain (K6), so the
        # fixed-order sum of the slabs into the flat gradient buffer runs on the side stream next to it.
        n_slabs = L.vcm_coeff_grad_slabs(t.handle, B, I, M, precision)
        _lib.call("vcm_coeff_grad_slabs_launch", t.handle, _ptr(dz), _ptr(x), _ptr(A), _ptr(ws), _ptr(dG), B, I, M,
                  precision, st,
                  nbytes=4 * (dz.numel() + x.numel() + 2 * nnz * M + nnz) + gb,
                  flops=4 * B * nnz * M * I, tag=tag)
        _side_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(_side_stream):
            _lib.call("vcm_sum_slabs", _ptr(ws), _ptr(dA), dA.numel(), n_slabs, _lib.stream_ptr(),
                      nbytes=4 * dA.numel() * (n_slabs + 1), flops=dA.numel() * n_slabs, tag=tag)
        _side_keep.append((ws,))
    else:
        _lib.call("vcm_coeff_grad_ex", t.handle, _ptr(dz), _ptr(x), _ptr(A), _ptr(dA), _ptr(dG), _ptr(ws), B, I, M,
                  precision, st,
                  nbytes=4 * (dz.numel() + (x.numel() if need_dA else 0) + (2 if need_dA else 1) * nnz * M + nnz) +
                  (gb if need_dx else 0),
                  flops=2 * (int(need_dx) + int(need_dA)) * B * nnz * M * I, tag=tag)
    dx = None
    if need_dx:
        dx = torch.empty_like(x)
        _lib.call("vcm_scatter_rev", t.handle, _ptr(dG), _ptr(dx), B, I, 0, st,
                  nbytes=gb + 4 * nnz + 4 * (p_in + 1) + 4 * x.numel(), flops=B * nnz * I, tag=tag)
    return dx, (dA if need_dA and not direct_dA else None)


class _ProjectGather(torch.autograd.Function):
    """s[b,p,o] = sum_{n,m} A[p,n,m] u[b, idx[p,n], m, o] (project-first order, M*O <= 64)."""

    @staticmethod
    def forward(ctx, u, A, tables, M, O):
        u, A = _req(u, "u"), _req(A, "coefficients")
        B = u.shape[0]
        if tuple(u.shape) != (B, tables.p_in, M * O) or tuple(A.shape) != (tables.p_out, tables.n_max, M):
            raise RuntimeError("project-first gather shape mismatch: u %s A %s" % (tuple(u.shape), tuple(A.shape)))
        s = torch.empty(B, tables.p_out, O, device=u.device, dtype=torch.float32)
        nnz = tables.nnz
        _lib.call("vcm_project_gather_fwd", tables.handle, _ptr(u), _ptr(A), _ptr(s), B, M, O, _lib.stream_ptr(),
                  nbytes=4 * (B * nnz * M * O + nnz * M + nnz + s.numel()), flops=2 * B * nnz * M * O,
                  tag="P%d>%d MO%d" % (tables.p_in, tables.p_out, M * O))
        ctx.save_for_backward(u, A)
        ctx.tables, ctx.cfg = tables, (M, O)
        ctx.dA_target = _grad_target(A, ctx.needs_input_grad[1])
        return s

    @staticmethod
    def backward(ctx, ds):
        u, A = ctx.saved_tensors
        t, (M, O) = ctx.tables, ctx.cfg
        B = u.shape[0]
        need_du, need_dA = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        ds = _req(ds, "ds")
        direct = need_dA and ctx.dA_target is not None
        dA = (ctx.dA_target if direct else torch.empty_like(A)) if need_dA else None
        du = torch.empty_like(u) if need_du else None
        nnz = t.nnz
        tag = "P%d>%d MO%d" % (t.p_in, t.p_out, M * O)
        if direct and need_du and _side_stream is not None:
            # du feeds the rest of the backward pass, dA goes straight into the flat gradient buffer: the dA kernel
            # runs next to the chain on the side stream (the C call launches one kernel per requested gradient)
            _lib.call("vcm_project_gather_bwd", t.handle, _ptr(ds), _ptr(u), _ptr(A), None, _ptr(du), B, M, O,
                      _lib.stream_ptr(), nbytes=4 * (B * nnz * M * O + nnz * M + ds.numel() + du.numel()),
                      flops=2 * B * nnz * M * O, tag=tag)
            _side_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(_side_stream):
                _lib.call("vcm_project_gather_bwd", t.handle, _ptr(ds), _ptr(u), _ptr(A), _ptr(dA), None, B, M, O,
                          _lib.stream_ptr(), nbytes=4 * (B * nnz * M * O + 2 * nnz * M + ds.numel()),
                          flops=2 * B * nnz * M * O, tag=tag)
            _side_keep.append((ds, u))
        else:
            _lib.call("vcm_project_gather_bwd", t.handle, _ptr(ds), _ptr(u), _ptr(A), _ptr(dA), _ptr(du), B, M, O,
                      _lib.stream_ptr(),
                      nbytes=4 * (B * nnz * M * O + 2 * nnz * M + ds.numel() + (du.numel() if need_du else 0)),
                      flops=2 * (int(need_du) + int(need_dA)) * B * nnz * M * O, tag=tag)
        return du, (None if direct or not need_dA else dA), None, None, None


class _BasisGemm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, W, bias, res, M, I, O, act, scale_y, scale_res, precision):
        z, W = _req(z, "z"), _req(W, "weights")
        B, p_out, K = z.shape
        if K != M * I or W.numel() != M * O * I:
            raise RuntimeError("basis GEMM shape mismatch: z %s, weights %s, M=%d I=%d O=%d" %
                               (tuple(z.shape), tuple(W.shape), M, I, O))
        per_point = 0
        if bias is not None:
            bias = _req(bias, "bias")
            if bias.dim() == 2:
                if tuple(bias.shape) != (p_out, O):
                    raise RuntimeError("per-point bias must be [%d, %d]" % (p_out, O))
                per_point = 1
            elif tuple(bias.shape) != (O,):
                raise RuntimeError("shared bias must be [%d]" % O)
        if res is not None:
            res = _req(res, "residual")
            if tuple(res.shape) != (B, p_out, O):
                raise RuntimeError("residual must be [%d, %d, %d], got %s" % (B, p_out, O, tuple(res.shape)))
        out = torch.empty(B, p_out, O, device=z.device, dtype=torch.float32)
        # the activated value is what ELU' needs; it is `out` itself unless a blend follows
        y_act = None
        if act and (res is not None or scale_y != 1.0):
            y_act = torch.empty_like(out)
        if B and p_out:
            nb = 4 * (z.numel() + W.numel() + (bias.numel() if bias is not None else 0) + out.numel() +
                      (res.numel() if res is not None else 0) + (y_act.numel() if y_act is not None else 0))
            n_ws = _lib.lib().vcm_basis_fwd_workspace(B * p_out, M, I, O, precision)
            ws = torch.empty(n_ws, device=z.device, dtype=torch.float32) if n_ws else None
            _lib.call("vcm_basis_fwd", _ptr(z), _ptr(W), _ptr(bias), per_point, _ptr(res), _ptr(y_act), _ptr(out),
                      _ptr(ws), B, p_out, M, I, O, int(act), scale_y, scale_res, precision, _lib.stream_ptr(),
                      nbytes=nb, flops=2 * B * p_out * M * I * O, tag="R%d K%d O%d" % (B * p_out, M * I, O))
        y_saved = (y_act if y_act is not None else out) if act else None
        ctx.save_for_backward(z, W, y_saved)
        ctx.cfg = (M, I, O, int(act), scale_y, scale_res, precision, per_point, bias is not None, res is not None)
        ctx.dW_target = _grad_target(W, ctx.needs_input_grad[1])
        ctx.db_target = _grad_target(bias, ctx.needs_input_grad[2])
        ctx.Wt = None
        if _prep_stream is not None and ctx.needs_input_grad[0] and B and p_out:
            # the dz GEMM's stacked-bases operand depends on W only: built now, on the trainer's preparation
            # stream (a parallel branch from the start of the step), instead of in front of K3 on the backward chain
            n_wt = _lib.lib().vcm_basis_bwd_dz_workspace(B * p_out, M, I, O, precision)
            if n_wt:
                with torch.cuda.stream(_prep_stream):
                    ctx.Wt = torch.empty(n_wt, device=z.device, dtype=torch.float32)
                    _lib.call("vcm_basis_stack", _ptr(W), _ptr(ctx.Wt), M, I, O, _lib.stream_ptr(),
                              nbytes=8 * W.numel(), flops=0, tag="M%d I%d O%d" % (M, I, O))
        return out

    @staticmethod
    def backward(ctx, dout):
        z, W, y = ctx.saved_tensors
        has_bias, has_res = ctx.cfg[8], ctx.cfg[9]
        need = (ctx.needs_input_grad[0], ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2],
                has_res and ctx.needs_input_grad[3])
        return _basis_backward(z, W, y, dout, ctx.cfg, need, ctx.dW_target, ctx.db_target, ctx.Wt) + (None,) * 7


def _basis_backward(z, W, y, dout, cfg, need, dW_target, db_target, Wt=None):
    """act_bwd + K3 + K4: (dz, dW, dbias, dres) of out = basis_gemm(z, W, bias, res); gradients written into a
    `*_target` come back as None."""
    M, I, O, act, sy, sr, precision, per_point, has_bias, has_res = cfg
    need_dz, need_dW, need_db, need_dres = need
    B, p_out, K = z.shape
    dout = _req(dout, "dout")
    L = _lib.lib()
    st = _lib.stream_ptr()
    dev = z.device
    if B == 0 or p_out == 0:
        return (torch.zeros_like(z), torch.zeros_like(W),
                (torch.zeros(p_out, O, device=dev) if per_point else torch.zeros(O, device=dev)) if need_db else None,
                torch.zeros_like(dout) if need_dres else None)
    direct_db = need_db and db_target is not None
    db_pp = None
    if need_db:
        db_pp = db_target if (direct_db and per_point) else torch.empty(p_out, O, device=dev, dtype=torch.float32)
    dres = torch.empty_like(dout) if need_dres else None
    ny = dout.numel()
    tag = "R%d K%d O%d" % (B * p_out, K, O)
    if not act and sy == 1.0 and not need_db and not need_dres:
        ds = dout                                # plain projection (residual 1x1, project-first bases): ds is dout itself
    else:
        ds = torch.empty_like(dout)
        _lib.call("vcm_act_bwd", _ptr(dout), _ptr(y), _ptr(ds), _ptr(db_pp), _ptr(dres), B, p_out, O, act, sy, sr, st,
                  nbytes=4 * (ny * (2 + (1 if act else 0) + (1 if need_dres else 0)) + (p_out * O if need_db else 0)),
                  flops=2 * ny, tag=tag)
    dbias = None
    if need_db:
        if per_point:
            dbias = db_pp
        else:
            dbias = db_target if direct_db else torch.empty(O, device=dev, dtype=torch.float32)
            _lib.call("vcm_colsum", _ptr(db_pp), _ptr(dbias), p_out, O, st, nbytes=4 * (p_out * O + O),
                      flops=p_out * O, tag=tag)
        if direct_db:
            dbias = None
    R = B * p_out
    dz = dW = None
    if need_dz:
        dz = torch.empty_like(z)
        if Wt is not None:                       # stacked bases prepared during the forward pass
            _lib.call("vcm_basis_bwd_dz_stacked", _ptr(ds), _ptr(Wt), _ptr(dz), R, M, I, O, precision, st,
                      nbytes=4 * (ny + W.numel() + z.numel()), flops=2 * R * K * O, tag=tag)
        else:
            n_ws = L.vcm_basis_bwd_dz_workspace(R, M, I, O, precision)
            ws = torch.empty(n_ws, device=dev, dtype=torch.float32) if n_ws else None
            _lib.call("vcm_basis_bwd_dz", _ptr(ds), _ptr(W), _ptr(dz), _ptr(ws), R, M, I, O, 0, precision, st,
                      nbytes=4 * (ny + W.numel() + z.numel()), flops=2 * R * K * O, tag=tag)
    if need_dW:
        direct_dW = dW_target is not None
        dW = dW_target if direct_dW else torch.empty_like(W)
        n_ws = L.vcm_basis_bwd_dw_workspace(R, M, I, O, precision)
        if direct_dW and _side_stream is not None:
            # dW depends on (ds, z) only and nothing downstream of this layer reads it: it runs as a parallel branch
            # (second stream; a parallel branch of the captured graph) next to the dz -> K5 -> K6 chain
            cur = torch.cuda.current_stream()
            _side_stream.wait_stream(cur)
            with torch.cuda.stream(_side_stream):
                ws = torch.empty(max(n_ws, 1), device=dev, dtype=torch.float32)
                _lib.call("vcm_basis_bwd_dw", _ptr(ds), _ptr(z), _ptr(dW), _ptr(ws), R, M, I, O, precision,
                          _lib.stream_ptr(), nbytes=4 * (ny + W.numel() + z.numel()), flops=2 * R * K * O, tag=tag)
            _side_keep.append((ds, z, ws))            # their memory must not be recycled before the join
        else:
            ws = torch.empty(max(n_ws, 1), device=dev, dtype=torch.float32)
            _lib.call("vcm_basis_bwd_dw", _ptr(ds), _ptr(z), _ptr(dW), _ptr(ws), R, M, I, O, precision, st,
                      nbytes=4 * (ny + W.numel() + z.numel()), flops=2 * R * K * O, tag=tag)
        if direct_dW:
            dW = None
    return dz, dW, dbias, dres


# Optional second stream for the dW GEMMs of a backward pass (set by the trainer, which also joins it).
_side_stream = None
_side_keep = []


def set_side_stream(stream):
    global _side_stream
    _side_stream = stream


# Optional preparation stream: work that depends on the parameters only (the stacked bases of the dz GEMMs). The
# trainer forks it from the main stream at the start of a step and joins it before the backward pass.
_prep_stream = None


def set_prep_stream(stream):
    global _prep_stream
    _prep_stream = stream


# Optional stream for the dA half of K5 (see _gather_backward); joined together with the dW stream.
_da_stream = None


def set_da_stream(stream):
    global _da_stream
    _da_stream = stream


# Optional stream for the residual branch of every layer (forward; autograd replays a node's backward on the
# stream of its forward, so the branch's backward overlaps the main chain as well). Set by the trainer.
_residual_stream = None


def set_residual_stream(stream):
    global _residual_stream
    _residual_stream = stream


def residual_stream():
    return _residual_stream


# Optional stream for the second encoder head (mu and std share their input and stage 1; their basis GEMMs are a few
# CTAs each and independent, forward and backward). Set by the trainer.
_head_stream = None


def set_head_stream(stream):
    global _head_stream
    _head_stream = stream


def head_stream():
    return _head_stream


def join_side_stream():
    """The current stream waits for the branch; buffers parked for it are released."""
    for st in (_side_stream, _da_stream):
        if st is not None:
            torch.cuda.current_stream().wait_stream(st)
    _side_keep.clear()


class _FusedConv(torch.autograd.Function):
    """out = epi(gather_contract(x, A) @ stacked(W)) in one launch (vcm_conv_fwd_fused): z is produced on chip and
    only written to HBM (once) when a backward pass will need it."""

    @staticmethod
    def forward(ctx, x, A, W, bias, res, tables, M, I, O, act, scale_y, scale_res, precision):
        x, A, W = _req(x, "x"), _req(A, "coefficients"), _req(W, "weights")
        B, p_in, _ = x.shape
        p_out = tables.p_out
        per_point = 0
        if bias is not None:
            bias = _req(bias, "bias")
            per_point = 1 if bias.dim() == 2 else 0
        if res is not None:
            res = _req(res, "residual")
        out = torch.empty(B, p_out, O, device=x.device, dtype=torch.float32)
        y_act = torch.empty_like(out) if act and (res is not None or scale_y != 1.0) else None
        z = torch.empty(B, p_out, M * I, device=x.device, dtype=torch.float32) if any(ctx.needs_input_grad[:5]) else None
        n_ws = _lib.lib().vcm_conv_fwd_fused_workspace(tables.handle, B, M, I, O, precision)
        ws = torch.empty(n_ws, device=x.device, dtype=torch.float32) if n_ws else None
        nnz = tables.nnz
        _lib.call("vcm_conv_fwd_fused", tables.handle, _ptr(x), _ptr(A), _ptr(W), _ptr(bias), per_point, _ptr(res),
                  _ptr(y_act), _ptr(out), _ptr(z), _ptr(ws), B, M, I, O, int(act), scale_y, scale_res, precision,
                  _lib.stream_ptr(),
                  nbytes=4 * (x.numel() + nnz * M + nnz + W.numel() + out.numel() * (2 if res is not None else 1) +
                              (z.numel() if z is not None else 0) + (y_act.numel() if y_act is not None else 0)),
                  flops=2 * B * nnz * M * I + 2 * B * p_out * M * I * O,
                  tag="P%d>%d I%d O%d" % (p_in, p_out, I, O))
        y_saved = (y_act if y_act is not None else out) if act else None
        ctx.save_for_backward(x, A, W, z, y_saved)
        ctx.tables = tables
        ctx.cfg = (M, I, O, int(act), scale_y, scale_res, precision, per_point, bias is not None, res is not None)
        ctx.dA_target = _grad_target(A, ctx.needs_input_grad[1])
        ctx.dW_target = _grad_target(W, ctx.needs_input_grad[2])
        ctx.db_target = _grad_target(bias, ctx.needs_input_grad[3])
        return out

    @staticmethod
    def backward(ctx, dout):
        x, A, W, z, y = ctx.saved_tensors
        has_bias, has_res = ctx.cfg[8], ctx.cfg[9]
        need_dx, need_dA = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need = (need_dx or need_dA, ctx.needs_input_grad[2], has_bias and ctx.needs_input_grad[3],
                has_res and ctx.needs_input_grad[4])
        dz, dW, dbias, dres = _basis_backward(z, W, y, dout, ctx.cfg, need, ctx.dW_target, ctx.db_target)
        dx, dA = _gather_backward(ctx.tables, x, A, dz, need_dx, need_dA, ctx.dA_target, ctx.cfg[6])
        return (dx, dA, dW, dbias, dres) + (None,) * 8


class _BiasAct(torch.autograd.Function):
    """out = (act ? ELU : id)(s + bias) * scale_y + res * scale_res, the K2 epilogue on its own."""

    @staticmethod
    def forward(ctx, s, bias, res, act, scale_y, scale_res):
        s = _req(s, "s")
        B, p_out, O = s.shape
        per_point = 0
        if bias is not None:
            bias = _req(bias, "bias")
            per_point = 1 if bias.dim() == 2 else 0
            if tuple(bias.shape) not in ((p_out, O), (O,)):
                raise RuntimeError("bias must be [%d, %d] or [%d]" % (p_out, O, O))
        if res is not None:
            res = _req(res, "residual")
            if tuple(res.shape) != tuple(s.shape):
                raise RuntimeError("residual shape mismatch")
        out = torch.empty_like(s)
        y_act = torch.empty_like(s) if act and (res is not None or scale_y != 1.0) else None
        if s.numel():
            _lib.call("vcm_bias_act_fwd", _ptr(s), _ptr(bias), per_point, _ptr(res), _ptr(y_act), _ptr(out), B, p_out,
                      O, int(act), scale_y, scale_res, _lib.stream_ptr(), nbytes=4 * s.numel() * 3, flops=4 * s.numel(),
                      tag="R%d O%d" % (B * p_out, O))
        ctx.save_for_backward((y_act if y_act is not None else out) if act else None)
        ctx.cfg = (int(act), scale_y, scale_res, per_point, bias is not None, res is not None, O)
        ctx.db_target = _grad_target(bias, ctx.needs_input_grad[1])
        return out

    @staticmethod
    def backward(ctx, dout):
        (y,) = ctx.saved_tensors
        act, sy, sr, per_point, has_bias, has_res, O = ctx.cfg
        dout = _req(dout, "dout")
        B, p_out, _ = dout.shape
        need_db = has_bias and ctx.needs_input_grad[1]
        need_dres = has_res and ctx.needs_input_grad[2]
        ds = torch.empty_like(dout)
        direct_db = need_db and ctx.db_target is not None
        db_pp = None
        if need_db:
            db_pp = ctx.db_target if (direct_db and per_point) else torch.empty(p_out, O, device=dout.device,
                                                                               dtype=torch.float32)
        dres = torch.empty_like(dout) if need_dres else None
        st = _lib.stream_ptr()
        _lib.call("vcm_act_bwd", _ptr(dout), _ptr(y), _ptr(ds), _ptr(db_pp), _ptr(dres), B, p_out, O, act, sy, sr, st,
                  nbytes=4 * dout.numel() * 3, flops=2 * dout.numel(), tag="R%d O%d" % (B * p_out, O))
        dbias = db_pp
        if need_db and not per_point:
            dbias = ctx.db_target if direct_db else torch.empty(O, device=dout.device, dtype=torch.float32)
            _lib.call("vcm_colsum", _ptr(db_pp), _ptr(dbias), p_out, O, st, nbytes=4 * (p_out * O + O), flops=p_out * O)
        if direct_db:
            dbias = None
        return ds, dbias, dres, None, None, None


class _EdgeWeights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p_nb, tables):
        p_nb = _req(p_nb, "p_neighbors")
        if tuple(p_nb.shape) != (tables.p_out, tables.n_max):
            raise RuntimeError("p_neighbors must be [%d, %d]" % (tables.p_out, tables.n_max))
        pn = torch.empty_like(p_nb)
        _lib.call("vcm_edge_weights_fwd", tables.handle, _ptr(p_nb), _ptr(pn), _lib.stream_ptr(),
                  nbytes=12 * p_nb.numel(), flops=3 * p_nb.numel())
        ctx.save_for_backward(p_nb)
        ctx.tables = tables
        ctx.dp_target = _grad_target(p_nb, ctx.needs_input_grad[0])
        return pn

    @staticmethod
    def backward(ctx, dpn):
        (p_nb,) = ctx.saved_tensors
        dpn = _req(dpn, "dpn")
        direct = ctx.dp_target is not None
        dp = ctx.dp_target if direct else torch.empty_like(p_nb)
        _lib.call("vcm_edge_weights_bwd", ctx.tables.handle, _ptr(p_nb), _ptr(dpn), _ptr(dp), _lib.stream_ptr(),
                  nbytes=16 * p_nb.numel(), flops=6 * p_nb.numel())
        return (None if direct else dp), None


class _ReparamKL(torch.autograd.Function):
    """(z, kl) of graphVAE_train.py:199-206 from the two UNSCALED encoder heads (MCEnc scales them by 0.1 / 0.01,
    graphVAESSW.py:286-288): one launch forward, one backward."""

    @staticmethod
    def forward(ctx, raw_mu, raw_logstd, eps, scale_mu, scale_logstd):
        raw_mu, raw_logstd, eps = _req(raw_mu, "mu head"), _req(raw_logstd, "log-std head"), _req(eps, "eps")
        if raw_mu.shape != raw_logstd.shape or eps.shape != raw_mu.shape:
            raise RuntimeError("mu, log-std and eps must have one shape, got %s %s %s" %
                               (tuple(raw_mu.shape), tuple(raw_logstd.shape), tuple(eps.shape)))
        n = raw_mu.numel()
        z = torch.empty_like(raw_mu)
        kl = torch.empty((), device=raw_mu.device, dtype=torch.float32)
        n_ws = _lib.lib().vcm_reparam_kl_workspace(n)
        ws = torch.empty(n_ws, device=raw_mu.device, dtype=torch.float32) if n_ws else None
        _lib.call("vcm_reparam_kl_fwd", _ptr(raw_mu), _ptr(raw_logstd), _ptr(eps), _ptr(z), _ptr(kl), _ptr(ws), n,
                  scale_mu, scale_logstd, _lib.stream_ptr(), nbytes=16 * n, flops=12 * n, tag="n%d" % n)
        ctx.save_for_backward(raw_mu, raw_logstd, eps)
        ctx.scales = (scale_mu, scale_logstd)
        return z, kl

    @staticmethod
    def backward(ctx, dz, dkl):
        raw_mu, raw_logstd, eps = ctx.saved_tensors
        n = raw_mu.numel()
        dz = _req(dz, "dz") if dz is not None else None
        dkl = _req(dkl, "dkl") if dkl is not None else None
        d_mu, d_ls = torch.empty_like(raw_mu), torch.empty_like(raw_logstd)
        _lib.call("vcm_reparam_kl_bwd", _ptr(raw_mu), _ptr(raw_logstd), _ptr(eps), _ptr(dz), _ptr(dkl), _ptr(d_mu),
                  _ptr(d_ls), n, ctx.scales[0], ctx.scales[1], _lib.stream_ptr(), nbytes=24 * n, flops=12 * n,
                  tag="n%d" % n)
        return d_mu, d_ls, None, None, None


class _LossHead(torch.autograd.Function):
    """(total = w_pose*l1 + w_lap*laplace + w_nor*normal, parts = [l1, laplace, normal]) of
    graphVAE_train.py:212-225 (KL excluded) in two launches; the backward is one launch. Only total is
    differentiable."""

    @staticmethod
    def forward(ctx, out_pc, gt_pc, nor, lap_coef, lap_tables, face_tables, w_pose, w_lap, w_nor):
        out_pc, gt_pc, nor = _req(out_pc, "out_pc"), _req(gt_pc, "gt_pc"), _req(nor, "face normals")
        lap_coef = _req(lap_coef, "laplacian coefficients")
        B, P, C = out_pc.shape
        F = face_tables.p_out
        if C != 3 or tuple(gt_pc.shape) != (B, P, 3) or tuple(nor.shape) != (B, F, 3):
            raise RuntimeError("loss head needs out/gt [B,%d,3] and normals [B,%d,3], got %s %s %s" %
                               (P, F, tuple(out_pc.shape), tuple(gt_pc.shape), tuple(nor.shape)))
        if lap_tables.p_out != P or lap_coef.numel() != P * lap_tables.n_max:
            raise RuntimeError("laplacian table / coefficients do not match %d vertices" % P)
        dev = out_pc.device
        lapv = torch.empty(B, P, 3, device=dev, dtype=torch.float32)
        s = torch.empty(B, F, device=dev, dtype=torch.float32)
        total = torch.empty((), device=dev, dtype=torch.float32)
        parts = torch.empty(3, device=dev, dtype=torch.float32)
        ws = torch.empty(_lib.lib().vcm_loss_head_workspace(lap_tables.handle, face_tables.handle, B), device=dev,
                         dtype=torch.float32)
        nb = 4 * (2 * B * P * 3 * 2 + lap_tables.nnz * 2 + B * F * 4 + B * P * 3 + B * F)
        _lib.call("vcm_loss_head_fwd", lap_tables.handle, face_tables.handle, _ptr(lap_coef), _ptr(out_pc), _ptr(gt_pc),
                  _ptr(nor), _ptr(lapv), _ptr(s), _ptr(total), _ptr(parts), _ptr(ws), B, w_pose, w_lap, w_nor, _lib.stream_ptr(),
                  nbytes=nb, flops=B * (6 * lap_tables.nnz + 20 * F + 12 * P), tag="P%d F%d" % (P, F))
        ctx.save_for_backward(out_pc, gt_pc, nor, lap_coef, lapv, s)
        ctx.cfg = (lap_tables, face_tables, w_pose, w_lap, w_nor)
        ctx.mark_non_differentiable(parts)
        return total, parts

    @staticmethod
    def backward(ctx, g_total, _g_parts):
        out_pc, gt_pc, nor, lap_coef, lapv, s = ctx.saved_tensors
        lap_tables, face_tables, w_pose, w_lap, w_nor = ctx.cfg
        B, P, _ = out_pc.shape
        g_total = _req(g_total, "d loss")
        d_out = torch.empty_like(out_pc)
        nb = 4 * (B * P * 3 * 4 + lap_tables.nnz * 2 + B * face_tables.nnz * 4)
        _lib.call("vcm_loss_head_bwd", lap_tables.handle, face_tables.handle, _ptr(lap_coef), _ptr(out_pc), _ptr(gt_pc),
                  _ptr(nor), _ptr(lapv), _ptr(s), _ptr(g_total), _ptr(d_out), B, w_pose, w_lap, w_nor,
                  _lib.stream_ptr(), nbytes=nb, flops=B * (6 * lap_tables.nnz + 6 * face_tables.nnz + 12 * P),
                  tag="P%d F%d" % (P, face_tables.p_out))
        d_gt = -d_out if ctx.needs_input_grad[1] else None
        return d_out, d_gt, None, None, None, None, None, None, None


def reparam_kl(raw_mu, raw_logstd, eps, scale_mu=1.0, scale_logstd=1.0):
    """z = mu + exp(logstd)*eps and the KL term mean(-0.5 - logstd + 0.5 mu^2 + 0.5 exp(2 logstd)) with
    mu = raw_mu*scale_mu, logstd = raw_logstd*scale_logstd. Returns (z, kl scalar)."""
    return _ReparamKL.apply(raw_mu, raw_logstd, eps, float(scale_mu), float(scale_logstd))


def loss_head(out_pc, gt_pc, nor, lap_coef, lap_tables, face_tables, w_pose, w_lap, w_nor):
    """(w_pose*l1 + w_lap*laplace + w_nor*normal, [l1, laplace, normal]) on out/gt [B,P,3], face normals [B,F,3]."""
    return _LossHead.apply(out_pc, gt_pc, nor, lap_coef, lap_tables, face_tables, float(w_pose), float(w_lap),
                           float(w_nor))


def gather_contract(x, coeffs, tables):
    """z[b,p,m*I+i] = sum_n coeffs[p,n,m] * x[b, idx[p,n], i] over real neighbours."""
    return _GatherContract.apply(x, coeffs, tables)


def project_gather(u, coeffs, tables, M, O):
    """s[b,p,o] = sum_{n,m} coeffs[p,n,m] * u[b, idx[p,n], m*O+o] (needs M*O <= 64)."""
    return _ProjectGather.apply(u, coeffs, tables, M, O)


def basis_gemm(z, weights, bias, res, M, I, O, act, scale_y=1.0, scale_res=0.0, precision=None):
    """out = (act ? ELU : id)(z @ stacked(weights) + bias) * scale_y + res * scale_res."""
    return _BasisGemm.apply(z, weights, bias, res, M, I, O, bool(act), float(scale_y), float(scale_res),
                            _precision if precision is None else precision)


# Opt-in (VCM_FUSED_FWD=1): on the C2 training step the fused kernel is still slower than K1 + K2 (its CUDA-core
# producer runs at one CTA per SM); see DESIGN.md section 7. The C entry point itself is always available.
fused_conv_enabled = __import__("os").environ.get("VCM_FUSED_FWD", "0") == "1"


def fused_conv_kind(tables, B, M, I, O, precision=None):
    """0 = K1 + K2, 1 = tcgen05 fused kernel (opt-in), 2 = the 3-channel kernel (used by default)."""
    prec = _precision if precision is None else precision
    return int(_lib.lib().vcm_conv_fwd_fused_kind(tables.handle, B, M, I, O, prec))


def fused_conv_supported(tables, B, M, I, O, precision=None):
    prec = _precision if precision is None else precision
    return bool(_lib.lib().vcm_conv_fwd_fused_supported(tables.handle, B, M, I, O, prec))


def fused_conv(x, coeffs, weights, bias, res, tables, M, I, O, act, scale_y=1.0, scale_res=0.0, precision=None):
    """basis_gemm(gather_contract(x, coeffs, tables), ...) in one launch; check fused_conv_supported first."""
    return _FusedConv.apply(x, coeffs, weights, bias, res, tables, M, I, O, bool(act), float(scale_y),
                            float(scale_res), _precision if precision is None else precision)


def bias_act(s, bias, res, act, scale_y=1.0, scale_res=0.0):
    return _BiasAct.apply(s, bias, res, bool(act), float(scale_y), float(scale_res))


def edge_weights(p_neighbors, tables):
    return _EdgeWeights.apply(p_neighbors, tables)


def blend_scales(residual_rate):
    """sqrt(1-r), sqrt(r) as the reference applies them to fp32 tensors
    (np.sqrt of a python float, graphVAESSW.py:161)."""
    return float(np.float32(np.sqrt(1 - residual_rate))), float(np.float32(np.sqrt(residual_rate)))


def vcconv(x, coeffs, weights, bias, tables, *, is_final_layer=False, residual_rate=0.0, p_neighbors=None,
           weight_res=None, precision=None):
    """One vc-conv / vd-pool / vd-unpool layer (LASMConvssw.forward,
    graphVAESSW.py:97-163) on the CUDA kernels."""
    B, p_in, I = x.shape
    M = weights.shape[0]
    O = weights.shape[1] // I
    z = gather_contract(x, coeffs, tables)
    act = not is_final_layer
    if residual_rate == 0:
        return basis_gemm(z, weights, bias, None, M, I, O, act, 1.0, 0.0, precision)
    sy, sr = blend_scales(residual_rate)
    xr = x
    if I != O:                                   # :144-145, 1x1 projection = basis GEMM with one basis, no bias
        xr = basis_gemm(x, weight_res, None, None, 1, I, O, False, 1.0, 0.0, precision)
    if p_in == tables.p_out:                     # :148-149 identity skip
        res = xr
    else:                                        # :151-159 weighted pool / unpool of the projected rows
        pn = edge_weights(p_neighbors, tables)
        res = gather_contract(xr, pn.unsqueeze(-1), tables)
    return basis_gemm(z, weights, bias, res, M, I, O, act, sy, sr, precision)
