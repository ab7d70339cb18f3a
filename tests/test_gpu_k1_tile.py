"""K1 with shared-memory staged neighbour rows on the tensor pipe (csrc/gather_fwd_tile.cu; the TF32-mode default of
vcm_gather_contract_fwd_ex): z[b,p,m,i] = sum_n A[p,n,m] x[b,idx[p,n],i]  (graphVAESSW.py:111-123).

Exact (<= 2e-5) against the fp64 contraction of TF32-truncated operands -- pins the staging layout, the permuted k
slots and the fragment mapping independently of rounding --, within the stated TF32 tolerance (3e-3) of the fp32
kernel, bitwise reproducible, no NaN from padding rows, every element of z written."""
import numpy as np
import pytest
import torch

from test_gpu_fused import random_table, rel_err, trunc_tf32
from vcmeshconv_b200 import _lib
from vcmeshconv_b200.connectivity import DeviceTables

pytestmark = pytest.mark.gpu
TF32 = _lib.PRECISIONS["tf32"]

CASES = [  # B, p_in, p_out, n_max, M, I
    (16, 60, 45, 7, 17, 32),            # one k-step
    (16, 344, 300, 12, 17, 128),        # unpool-like, two k-steps, several units per warp
    (16, 200, 200, 20, 17, 64),         # three k-steps, per-vertex trip counts
    (16, 100, 20, 31, 17, 256),         # few vertices, many units per vertex
    (3, 50, 40, 9, 17, 32),             # odd batch
    (1, 50, 40, 9, 17, 64),
    (16, 90, 33, 58, 17, 64),           # pooling-like: > 24 neighbours keep the fp32 kernel
    (8, 64, 64, 6, 9, 64),              # M = 9: one basis tile
    (16, 64, 64, 6, 24, 32),            # M = 24 fills the second basis tile
    (2, 30, 25, 5, 16, 512),            # M = 16, wide rows
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_P%d-%d_N%d_M%d_I%d" % c)
def test_gather_contract_tile_kernel(case):
    B, p_in, p_out, n_max, M, I = case
    tab, ids = random_table(p_in, p_out, n_max, seed=p_out + n_max)
    t = DeviceTables(tab, p_in)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(B, p_in, I, device="cuda", generator=g)
    A = torch.randn(p_out, n_max, M, device="cuda", generator=g)
    A[torch.from_numpy(ids == p_in).cuda()] = 1e3                   # padded coefficient slots hold arbitrary parameters
    L = _lib.lib()
    z = torch.full((B, p_out, M * I), float("nan"), device="cuda")
    z2 = torch.full_like(z, float("nan"))
    z32 = torch.empty_like(z)
    for out in (z, z2):
        _lib.check(L.vcm_gather_contract_fwd_ex(t.handle, x.data_ptr(), A.data_ptr(), out.data_ptr(), B, I, M, TF32,
                                                _lib.stream_ptr()))
    _lib.check(L.vcm_gather_contract_fwd(t.handle, x.data_ptr(), A.data_ptr(), z32.data_ptr(), B, I, M, _lib.stream_ptr()))
    torch.cuda.synchronize()
    idt = torch.from_numpy(ids).cuda()
    real = (idt != p_in)
    A0 = torch.where(real[:, :, None], A, torch.zeros_like(A))
    xpad = torch.cat([trunc_tf32(x).double(), torch.zeros(B, 1, I, device="cuda", dtype=torch.float64)], 1)
    want = torch.einsum("pnm,bpni->bpmi", trunc_tf32(A0).double(), xpad[:, idt]).reshape(B, p_out, M * I)
    assert torch.isfinite(z).all()                                   # every element written, no NaN from padding
    assert torch.equal(z, z2)                                        # bitwise reproducible
    if n_max <= 24:                                                  # the tile kernel served it (wider tables: fp32 kernel)
        assert rel_err(z, want) < 2e-5
    assert rel_err(z, z32) < 3e-3


@pytest.mark.parametrize("case", [(16, 120, 77, 9, 64), (4, 300, 40, 47, 256), (16, 60, 200, 12, 3)],
                         ids=lambda c: "B%d_P%d-%d_N%d_I%d" % c)
def test_scaled_reverse_scatter_equals_coeff_grad_plus_scatter(case):
    """M = 1 gathers (residual pool / un-pool): dx[b,q,:] = sum_{slot in rev(q)} coef[slot] * dy[b, slot // N, :]
    (vcm_scatter_rev_scaled, no per-edge gradient) against the fp64 reference and against K5 (dG) + K6."""
    B, p_in, p_out, n_max, I = case
    tab, ids = random_table(p_in, p_out, n_max, seed=3 * p_out + n_max)
    t = DeviceTables(tab, p_in)
    g = torch.Generator(device="cuda").manual_seed(21)
    dy = torch.randn(B, p_out, I, device="cuda", generator=g)
    coef = torch.randn(p_out, n_max, device="cuda", generator=g)
    x = torch.randn(B, p_in, I, device="cuda", generator=g)
    L = _lib.lib()
    dx = torch.full((B, p_in, I), float("nan"), device="cuda")
    dx2 = torch.full_like(dx, float("nan"))
    _lib.check(L.vcm_scatter_rev_scaled(t.handle, dy.data_ptr(), coef.data_ptr(), dx.data_ptr(), B, I, 0, _lib.stream_ptr()))
    dG = torch.empty(B, p_out, n_max, I, device="cuda")
    _lib.check(L.vcm_coeff_grad(t.handle, dy.data_ptr(), x.data_ptr(), coef.data_ptr(), None, dG.data_ptr(), None, B, I, 1,
                                _lib.stream_ptr()))
    _lib.check(L.vcm_scatter_rev(t.handle, dG.data_ptr(), dx2.data_ptr(), B, I, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    idt = torch.from_numpy(ids).cuda()
    real = (idt != p_in)
    want = torch.zeros(B, p_in + 1, I, dtype=torch.float64, device="cuda")
    contrib = (coef.double() * real)[None, :, :, None] * dy.double()[:, :, None, :]          # [B, P, N, I]
    want.index_add_(1, idt.reshape(-1), contrib.reshape(B, p_out * n_max, I))
    want = want[:, :p_in]
    assert torch.isfinite(dx).all()
    assert rel_err(dx, want) < 1e-6
    assert rel_err(dx, dx2) < 1e-6
