"""K5 on the tensor pipe (vcm_coeff_grad_ex, TF32 and 3-pass modes) against an fp64 reference.

  dA[p,n,m] = sum_{b,i} dz[b,p,m,i] * x[b, idx[p,n], i]      dG[b,p,n,i] = sum_m A[p,n,m] * dz[b,p,m,i]

TF32 mode (the bench default; the tensor-pipe kernel is the default for <= 32 neighbours): stated tolerance 3e-3
norm-wise, and <= 2e-5 against the fp64 product of TF32-truncated operands (pins staging layout, swizzle, fragment
mapping independently of rounding). tf32x3 mode: fp32-grade, <= 1e-5. Padding slots must come back exactly 0.
"""
import numpy as np
import pytest
import torch

from test_gpu_fused import random_table, rel_err, trunc_tf32
from vcmeshconv_b200 import _lib
from vcmeshconv_b200.connectivity import DeviceTables

pytestmark = pytest.mark.gpu

CASES = [  # B, p_in, p_out, n_max, M, I
    (16, 60, 45, 7, 17, 32),
    (16, 344, 300, 12, 17, 128),        # unpool-like, several tiles per warp, split over columns
    (16, 200, 200, 20, 17, 64),         # two 16-neighbour blocks
    (16, 100, 20, 31, 17, 256),         # few vertices: column split + fixed-order sum of the splits
    (3, 50, 40, 9, 17, 32),             # odd batch
    (1, 50, 40, 9, 17, 64),
    (16, 90, 33, 58, 17, 64),           # > 32 neighbours: default rule keeps the fp32 kernel, forced mode uses 4 blocks
    (8, 64, 64, 6, 9, 96),              # M = 9
    (16, 64, 64, 6, 24, 32),            # M = 24 fills the three basis tiles
]


def run(t, dz, x, A, B, I, M, prec, want_dA=True, want_dG=True):
    L = _lib.lib()
    dA = torch.full_like(A, float("nan")) if want_dA else None
    dG = torch.full((B, t.p_out, t.n_max, I), float("nan"), device="cuda") if want_dG else None
    ws = torch.empty(max(L.vcm_coeff_grad_workspace_ex(t.handle, B, I, M, prec), 1), device="cuda")
    p = lambda a: a.data_ptr() if a is not None else None
    _lib.check(L.vcm_coeff_grad_ex(t.handle, p(dz), p(x), p(A), p(dA), p(dG), p(ws), B, I, M, prec, None))
    torch.cuda.synchronize()
    return dA, dG


def reference(dz, x, A, ids, p_in):
    B, P, M, I = dz.shape
    xp = torch.cat([x.double(), torch.zeros(B, 1, I, dtype=torch.float64, device=x.device)], 1)
    idt = torch.from_numpy(ids).to(x.device)
    g = xp[:, idt]                                                        # [B, P, N, I]
    real = (idt != p_in)
    dA = torch.einsum("bpmi,bpni->pnm", dz.double(), g) * real[:, :, None]
    dG = torch.einsum("pnm,bpmi->bpni", A.double(), dz.double()) * real[None, :, :, None]
    return dA, dG, real


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_P%d-%d_N%d_M%d_I%d" % c)
@pytest.mark.parametrize("prec", ["tf32", "tf32x3"])
def test_coeff_grad_tensor_pipe(case, prec):
    B, p_in, p_out, n_max, M, I = case
    tab, ids = random_table(p_in, p_out, n_max, seed=p_out + n_max)
    t = DeviceTables(tab, p_in)
    g = torch.Generator(device="cuda").manual_seed(5)
    dz = torch.randn(B, p_out, M, I, device="cuda", generator=g)
    x = torch.randn(B, p_in, I, device="cuda", generator=g)
    A = torch.randn(p_out, n_max, M, device="cuda", generator=g)
    dA, dG = run(t, dz.view(B, p_out, M * I), x, A, B, I, M, _lib.PRECISIONS[prec])
    wA, wG, real = reference(dz, x, A, ids, p_in)
    # dG is only defined on real neighbour slots (K6 reads nothing else); dA padding slots are exactly 0
    dGm = torch.where(real[None, :, :, None], dG, torch.zeros_like(dG))
    assert torch.isfinite(dA).all() and torch.isfinite(dGm).all()
    assert (dA[~real] == 0).all()
    tol = 3e-3 if prec == "tf32" else 1e-5
    assert rel_err(dA, wA) < tol
    assert rel_err(dGm, wG) < tol
    if prec == "tf32" and n_max <= 64 and M >= 2 and I % 32 == 0:
        # the tensor-pipe kernel ran: it must match the product of truncated operands to accumulation round-off
        tA, tG, _ = reference(trunc_tf32(dz), trunc_tf32(x), trunc_tf32(A), ids, p_in)
        assert rel_err(dA, tA) < 2e-5
        assert rel_err(dGm, tG) < 2e-5


def test_optional_outputs_and_determinism():
    B, p_in, p_out, n_max, M, I = 16, 120, 77, 9, 17, 64
    tab, ids = random_table(p_in, p_out, n_max, seed=9)
    t = DeviceTables(tab, p_in)
    g = torch.Generator(device="cuda").manual_seed(6)
    dz = torch.randn(B, p_out, M * I, device="cuda", generator=g)
    x = torch.randn(B, p_in, I, device="cuda", generator=g)
    A = torch.randn(p_out, n_max, M, device="cuda", generator=g)
    tf32 = _lib.PRECISIONS["tf32"]
    dA, dG = run(t, dz, x, A, B, I, M, tf32)
    dA2, _ = run(t, dz, x, A, B, I, M, tf32, want_dG=False)
    _, dG2 = run(t, dz, x, A, B, I, M, tf32, want_dA=False)
    dA3, dG3 = run(t, dz, x, A, B, I, M, tf32)
    real = torch.from_numpy(ids != p_in).cuda()
    assert torch.equal(dA, dA2) and torch.equal(dA, dA3)                   # fixed-order reductions: bitwise reproducible
    assert torch.equal(dG[:, real], dG2[:, real]) and torch.equal(dG[:, real], dG3[:, real])


SMALL_CASES = [  # B, p_in, p_out, n_max, M, I   (dA only: coeff_grad_small_kernel)
    (16, 500, 130, 35, 17, 3),          # the 3-channel input layer: three neighbour blocks
    (16, 130, 500, 12, 1, 3),           # residual path over xyz: M = 1, one block
    (64, 300, 90, 21, 17, 3),           # C3 batch: 192 columns, 24 k-steps
    (5, 80, 70, 58, 17, 3),             # 15 columns (partial k-step), four neighbour blocks
    (7, 60, 60, 9, 24, 1),              # one channel, M = 24
    (3, 60, 60, 16, 8, 8),              # I = 8, M = 8 (one basis block)
]


@pytest.mark.parametrize("case", SMALL_CASES, ids=lambda c: "B%d_P%d-%d_N%d_M%d_I%d" % c)
@pytest.mark.parametrize("prec", ["tf32", "tf32x3"])
def test_coeff_grad_small_channels(case, prec):
    B, p_in, p_out, n_max, M, I = case
    tab, ids = random_table(p_in, p_out, n_max, seed=p_out + n_max)
    t = DeviceTables(tab, p_in)
    L = _lib.lib()
    assert L.vcm_coeff_grad_variant_ex(t.handle, B, I, M, _lib.PRECISIONS[prec], 1, 0) == 3
    assert L.vcm_coeff_grad_variant_ex(t.handle, B, I, M, _lib.PRECISIONS[prec], 1, 1) != 3
    assert L.vcm_coeff_grad_variant_ex(t.handle, B, I, M, _lib.PRECISIONS["fp32"], 1, 0) == 0
    g = torch.Generator(device="cuda").manual_seed(11)
    dz = torch.randn(B, p_out, M, I, device="cuda", generator=g)
    x = torch.randn(B, p_in, I, device="cuda", generator=g)
    A = torch.randn(p_out, n_max, M, device="cuda", generator=g)
    dA, _ = run(t, dz.view(B, p_out, M * I), x, A, B, I, M, _lib.PRECISIONS[prec], want_dG=False)
    wA, _, real = reference(dz, x, A, ids, p_in)
    assert torch.isfinite(dA).all()
    assert (dA[~real] == 0).all()                                          # padding slots: exact zeros
    assert rel_err(dA, wA) < (3e-3 if prec == "tf32" else 1e-5)
    if prec == "tf32":
        tA, _, _ = reference(trunc_tf32(dz), trunc_tf32(x), A, ids, p_in)
        assert rel_err(dA, tA) < 2e-5                                      # pins the fragment mapping
    dA2, _ = run(t, dz.view(B, p_out, M * I), x, A, B, I, M, _lib.PRECISIONS[prec], want_dG=False)
    assert torch.equal(dA, dA2)
    # and the fp32 kernel agrees (the same call in fp32 mode)
    dA3, _ = run(t, dz.view(B, p_out, M * I), x, A, B, I, M, _lib.PRECISIONS["fp32"], want_dG=False)
    assert rel_err(dA3, wA) < 1e-5
