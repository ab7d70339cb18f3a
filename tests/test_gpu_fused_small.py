"""K1 + K2 in one launch for the 3-channel layers (vcm_conv_fwd_fused, kind 2: fused_fwd_small_kernel) against fp64.

TF32 mode: stated tolerance 3e-3 norm-wise, and <= 2e-5 against the fp64 product of TF32-truncated operands (pins
both fragment mappings and the z re-layout independently of rounding). tf32x3 mode: fp32-grade, <= 1e-5.
"""
import numpy as np
import pytest
import torch

from test_gpu_fused import random_table, rel_err, trunc_tf32
from vcmeshconv_b200 import _lib, functional as F_
from vcmeshconv_b200.connectivity import DeviceTables

pytestmark = pytest.mark.gpu

CASES = [  # B, p_in, p_out, n_max, M, O        (I = 3)
    (16, 500, 130, 35, 17, 32),         # the input layer of the autoencoder
    (64, 300, 90, 21, 17, 32),          # C3 batch: four blocks of 16 batch entries
    (5, 80, 70, 58, 17, 64),            # partial batch block, > 32 neighbours, O = 64
    (20, 60, 61, 9, 24, 32),            # batch not a multiple of 16, M*I = 72 (nine k-steps of stage 2)
    (16, 60, 60, 9, 1, 32),
]


def fp64_layer(x, A, ids, W, bias, res, M, O, p_in, act, sy, sr, trunc=False):
    B, _, I = x.shape
    tr = trunc_tf32 if trunc else (lambda v: v)
    xp = torch.cat([tr(x).double(), torch.zeros(B, 1, I, dtype=torch.float64, device=x.device)], 1)
    idt = torch.from_numpy(ids).to(x.device)
    real = (idt != p_in).double()
    z = torch.einsum("pnm,bpni->bpmi", tr(A).double() * real[:, :, None], xp[:, idt])
    zz = tr(z.float()).double() if trunc else z
    s = torch.einsum("bpmi,moi->bpo", zz, tr(W).view(M, O, I).double())
    if bias is not None:
        s = s + bias.double()
    y = torch.where(s > 0, s, torch.expm1(s)) if act else s
    out = y * sy + (res.double() * sr if res is not None else 0)
    return z.reshape(B, -1, M * I), y, out


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_P%d-%d_N%d_M%d_O%d" % c)
@pytest.mark.parametrize("prec", ["tf32", "tf32x3"])
@pytest.mark.parametrize("variant", ["act_res_ppbias", "plain"])
def test_fused_small_forward(case, prec, variant):
    B, p_in, p_out, n_max, M, O = case
    I = 3
    tab, ids = random_table(p_in, p_out, n_max, seed=B + p_out)
    t = DeviceTables(tab, p_in)
    L = _lib.lib()
    P = _lib.PRECISIONS[prec]
    assert L.vcm_conv_fwd_fused_kind(t.handle, B, M, I, O, P) == 2
    assert L.vcm_conv_fwd_fused_kind(t.handle, B, M, I, O, _lib.PRECISIONS["fp32"]) == 0
    assert L.vcm_conv_fwd_fused_workspace(t.handle, B, M, I, O, P) == 0
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(B, p_in, I, device="cuda", generator=g)
    A = torch.randn(p_out, n_max, M, device="cuda", generator=g)
    A[torch.from_numpy(ids == p_in).cuda()] = 1e3                          # padding coefficients must not leak
    W = torch.randn(M, O * I, device="cuda", generator=g) / np.sqrt(M * I)
    full = variant == "act_res_ppbias"
    bias = torch.randn(p_out, O, device="cuda", generator=g) if full else None
    res = torch.randn(B, p_out, O, device="cuda", generator=g) if full else None
    sy, sr = (0.3, 0.9) if full else (1.0, 0.0)
    ptr = lambda a: a.data_ptr() if a is not None else None
    z64, y64, o64 = fp64_layer(x, A, ids, W, bias, res, M, O, p_in, full, sy, sr)
    for with_z in (True, False):
        out = torch.full((B, p_out, O), float("nan"), device="cuda")
        y = torch.full((B, p_out, O), float("nan"), device="cuda") if full else None
        z = torch.full((B, p_out, M * I), float("nan"), device="cuda") if with_z else None
        _lib.check(L.vcm_conv_fwd_fused(t.handle, ptr(x), ptr(A), ptr(W), ptr(bias), 1, ptr(res), ptr(y), ptr(out),
                                        ptr(z), None, B, M, I, O, int(full), sy, sr, P, None))
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        tol = 3e-3 if prec == "tf32" else 1e-5
        assert rel_err(out.double(), o64) < tol
        if y is not None:
            assert rel_err(y.double(), y64) < tol
        if z is not None:
            assert torch.isfinite(z).all()
            assert rel_err(z.double(), z64) < tol
        if prec == "tf32" and z is not None:
            zt, _, _ = fp64_layer(x, A, ids, W, bias, res, M, O, p_in, full, sy, sr, trunc=True)
            assert rel_err(z.double(), zt) < 2e-5                          # stage 1: truncated x and coefficients
            # stage 2 on the kernel's own z (truncating a different rounding of z would flip low bits)
            s = torch.einsum("bpmi,moi->bpo", trunc_tf32(z).view(B, p_out, M, I).double(),
                             trunc_tf32(W).view(M, O, I).double())
            if bias is not None:
                s = s + bias.double()
            yt = torch.where(s > 0, s, torch.expm1(s)) if full else s
            ot = yt * sy + (res.double() * sr if res is not None else 0)
            assert rel_err(out.double(), ot) < 2e-5


@pytest.mark.parametrize("prec", ["tf32", "tf32x3"])
def test_fused_small_autograd_matches_two_kernel_path(prec):
    B, p_in, p_out, n_max, M, I, O = 16, 200, 77, 30, 17, 3, 32
    tab, ids = random_table(p_in, p_out, n_max, seed=3)
    t = DeviceTables(tab, p_in)
    P = _lib.PRECISIONS[prec]
    assert F_.fused_conv_kind(t, B, M, I, O, P) == 2
    g = torch.Generator(device="cuda").manual_seed(11)
    mk = lambda *s: torch.randn(*s, device="cuda", generator=g).requires_grad_()
    x, A, W, bias, res = mk(B, p_in, I), mk(p_out, n_max, M), mk(M, O * I), mk(p_out, O), mk(B, p_out, O)
    dout = torch.randn(B, p_out, O, device="cuda", generator=g)
    o1 = F_.fused_conv(x, A, W, bias, res, t, M, I, O, True, 0.3, 0.9, P)
    g1 = torch.autograd.grad(o1, (x, A, W, bias, res), dout)
    o2 = F_.basis_gemm(F_.gather_contract(x, A, t), W, bias, res, M, I, O, True, 0.3, 0.9, P)
    g2 = torch.autograd.grad(o2, (x, A, W, bias, res), dout)
    tol = 3e-3 if prec == "tf32" else 1e-5
    assert rel_err(o1, o2) < tol
    for a, b in zip(g1, g2):
        # tf32: the two forwards round differently (the two-kernel path is fp32 FMA on this shape) and every gradient
        # then goes through TF32 products again
        assert rel_err(a, b) < (1e-2 if prec == "tf32" else tol)
