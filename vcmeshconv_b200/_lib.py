"""ctypes binding of libvcmesh.so (include/vcmesh.h).

The library is the product: if it cannot be loaded the import of any op fails
loudly — there is no eager / PyTorch / CPU fallback anywhere in this package.
"""
import ctypes as C
import os
import re

from . import build as _build

_lib = None

PREC_FP32, PREC_TF32, PREC_TF32X3 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "tf32x3": PREC_TF32X3}
TAB_IDX, TAB_LAST, TAB_PERM, TAB_REV_PTR, TAB_REV_SLOT, TAB_IDX_SORTED, TAB_META_SORTED = range(7)

_p, _i, _f, _i64, _sz = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t


class TablesInfo(C.Structure):
    _fields_ = [("p_in", C.c_int32), ("p_out", C.c_int32), ("n_max", C.c_int32), ("max_last", C.c_int32),
                ("nnz", C.c_int64), ("max_rev", C.c_int32), ("device", C.c_int32)]


# name -> (restype, argtypes); must list every VCM_API symbol of include/vcmesh.h
SIGNATURES = {
    "vcm_last_error": (C.c_char_p, []),
    "vcm_version": (_i, []),
    "vcm_launch_count": (C.c_ulonglong, []),
    "vcm_basis_uses_tensor_cores": (_i, [_i64, _i, _i, _i, _i]),
    "vcm_tables_create": (_i, [_p, _i, _i, _i, C.POINTER(_p)]),
    "vcm_tables_destroy": (None, [_p]),
    "vcm_tables_get_info": (_i, [_p, C.POINTER(TablesInfo)]),
    "vcm_tables_export": (_i, [_p, _i, _p, _sz]),
    "vcm_tables_device_ptr": (_p, [_p, _i]),
    "vcm_gather_contract_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "vcm_gather_contract_fwd_ex": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "vcm_project_gather_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "vcm_project_gather_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "vcm_coeff_grad_workspace": (_sz, [_p, _i, _i, _i]),
    "vcm_coeff_grad": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "vcm_coeff_grad_variant": (_i, [_p, _i, _i, _i, _i]),
    "vcm_coeff_grad_variant_ex": (_i, [_p, _i, _i, _i, _i, _i, _i]),
    "vcm_coeff_grad_slabs": (_i, [_p, _i, _i, _i, _i]),
    "vcm_coeff_grad_slabs_launch": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "vcm_sum_slabs": (_i, [_p, _p, _sz, _i, _p]),
    "vcm_coeff_grad_workspace_ex": (_sz, [_p, _i, _i, _i, _i]),
    "vcm_coeff_grad_ex": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "vcm_scatter_rev": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "vcm_scatter_rev_scaled": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "vcm_edge_weights_fwd": (_i, [_p, _p, _p, _p]),
    "vcm_edge_weights_bwd": (_i, [_p, _p, _p, _p, _p]),
    "vcm_basis_fwd_workspace": (_sz, [_i64, _i, _i, _i, _i]),
    "vcm_basis_fwd": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _f, _i, _p]),
    "vcm_conv_fwd_fused_supported": (_i, [_p, _i, _i, _i, _i, _i]),
    "vcm_conv_fwd_fused_kind": (_i, [_p, _i, _i, _i, _i, _i]),
    "vcm_conv_fwd_fused_workspace": (_sz, [_p, _i, _i, _i, _i, _i]),
    "vcm_conv_fwd_fused": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _i, _p]),
    "vcm_bias_act_fwd": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p]),
    "vcm_act_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p]),
    "vcm_colsum": (_i, [_p, _p, _i, _i, _p]),
    "vcm_basis_bwd_dz_workspace": (_sz, [_i64, _i, _i, _i, _i]),
    "vcm_basis_bwd_dz": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _p]),
    "vcm_basis_stack": (_i, [_p, _p, _i, _i, _i, _p]),
    "vcm_basis_bwd_dz_stacked": (_i, [_p, _p, _p, _i64, _i, _i, _i, _i, _p]),
    "vcm_basis_bwd_dw_workspace": (_sz, [_i64, _i, _i, _i, _i]),
    "vcm_basis_bwd_dw": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i, _i, _p]),
    "vcm_adam_step": (_i, [_p, _p, _p, _p, _sz, _f, _f, _f, _f, _i, _f, _p]),
    "vcm_adam_step_graph": (_i, [_p, _p, _p, _p, _sz, _p, _f, _f, _f, _f, _p]),
    "vcm_adam_advance": (_i, [_p, _f, _f, _p]),
    "vcm_adam_apply": (_i, [_p, _p, _p, _p, _sz, _p, _f, _f, _f, _f, _p]),
    "vcm_reduce_adam_chunk": (_i, []),
    "vcm_reduce_adam_blocks": (_i, [_sz, _i]),
    "vcm_reduce_adam_bcast": (_i, [_p, _p, _p, _p, _p, _sz, _sz, _p, _f, _f, _f, _i, _i, _i, _p]),
    "vcm_reparam_kl_workspace": (_sz, [_sz]),
    "vcm_reparam_kl_fwd": (_i, [_p, _p, _p, _p, _p, _p, _sz, _f, _f, _p]),
    "vcm_reparam_kl_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _sz, _f, _f, _p]),
    "vcm_loss_head_workspace": (_sz, [_p, _p, _i]),
    "vcm_loss_head_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _f, _f, _f, _p]),
    "vcm_loss_head_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _f, _f, _f, _p]),
}


def declared_symbols():
    """Every VCM_API function include/vcmesh.h declares (used by the CPU tests)."""
    hdr = open(os.path.join(_build.ROOT, "include", "vcmesh.h")).read()
    return sorted(set(re.findall(r"VCM_API[^;(]*?\b(vcm_\w+)\s*\(", hdr)))


class VcmError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("VCM_LIBRARY") or _build.LIB_CUDA      # override: A/B runs of two builds
        if not os.path.exists(path):
            _build.build_cuda()          # raises if nvcc is unavailable: no fallback
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)        # AttributeError = ABI drift, fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code):
    if code != 0:
        raise VcmError("libvcmesh: %s (code %d)" % (lib().vcm_last_error().decode(), code))


class Profiler:
    """Per-call CUDA-event timing of the C-ABI compute calls (bench.py roofline).
    Events are recorded on torch's current stream, which is the stream every
    call is enqueued on. Use as a context manager; results after a synchronize."""

    def __init__(self):
        self.records = []          # (name, tag, bytes, flops, start_event, end_event)

    def __enter__(self):
        global _profiler
        _profiler = self
        return self

    def __exit__(self, *exc):
        global _profiler
        _profiler = None

    def summary(self, pair_overhead_ms=0.0):
        """`pair_overhead_ms`: what an event pair around NOTHING measures on this stream (the record commands
        themselves); it is subtracted from every call's time."""
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, tag, nbytes, flops, e0, e1 in self.records:
            d = out.setdefault(name, {"calls": 0, "ms": 0.0, "bytes": 0, "flops": 0, "by_tag": {}})
            ms = max(e0.elapsed_time(e1) - pair_overhead_ms, 1e-4)
            d["calls"] += 1
            d["ms"] += ms
            d["bytes"] += nbytes
            d["flops"] += flops
            t = d["by_tag"].setdefault(tag, {"calls": 0, "ms": 0.0, "bytes": 0, "flops": 0})
            t["calls"] += 1
            t["ms"] += ms
            t["bytes"] += nbytes
            t["flops"] += flops
        return out


_profiler = None


def call(name, *args, nbytes=0, flops=0, tag=""):
    """Invoke one compute entry point; raises VcmError on a non-zero status."""
    fn = getattr(lib(), name)
    if _profiler is None:
        check(fn(*args))
        return
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args))
    e1.record()
    _profiler.records.append((name, tag, nbytes, flops, e0, e1))


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
