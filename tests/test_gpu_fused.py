"""K1 + K2 fused in one kernel (vcm_conv_fwd_fused) against the two-kernel path and an fp64 reference.

z written by the fused kernel must equal the stand-alone gather/contract (same fp32 FMA chain per element);
out is checked against the fp64 product of the TF32-truncated operands (pins the smem operand layout the
producer warps write, the (i, m, o) tensor map and the k ordering) and, norm-wise, against full precision.
"""
import numpy as np
import pytest
import torch

from helpers import rel_err as _rel_err
from vcmeshconv_b200 import _lib, functional as F_
from vcmeshconv_b200.connectivity import DeviceTables

pytestmark = pytest.mark.gpu
TF32 = _lib.PREC_TF32
TF32_TOL = 3e-3
EXACT_TOL = 2e-5


def rel_err(a, b):
    return _rel_err(a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy())


def trunc_tf32(t):
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def random_table(p_in, p_out, n_max, seed, holes=True):
    rng = np.random.RandomState(seed)
    tab = np.zeros((p_out, 1 + 2 * n_max), np.int32)
    ids = np.full((p_out, n_max), p_in, np.int64)
    for p in range(p_out):
        k = rng.randint(0 if holes else 1, n_max + 1)
        ids[p, :k] = rng.randint(0, p_in, size=k)
        if holes and k > 2 and rng.rand() < 0.3:
            ids[p, rng.randint(0, k - 1)] = p_in          # a padding entry in the middle of the list
        tab[p, 0] = int((ids[p] != p_in).sum())
    tab[:, 1::2] = ids
    return tab, ids


def reference(x, A, ids, W, M, I, O, p_in):
    """fp64: z and the stacked-basis product, optionally with TF32-truncated GEMM operands."""
    B = x.shape[0]
    xp = torch.cat([x.double(), torch.zeros(B, 1, I, dtype=torch.float64, device=x.device)], 1)
    g = xp[:, torch.from_numpy(ids).to(x.device)]                          # [B, P, N, I]
    z = torch.einsum("pnm,bpni->bpmi", A.double(), g)                     # [B, P, M, I]
    return z


CASES = [  # B, p_in, p_out, n_max, M, I, O
    (16, 40, 23, 7, 17, 32, 64),
    (16, 300, 344, 9, 17, 64, 128),       # several tiles, tail tile
    (16, 51, 10, 12, 17, 256, 256),       # few tiles, long reduction: split + second-pass epilogue
    (16, 100, 90, 20, 17, 16, 32),
    (16, 64, 64, 6, 17, 64, 512),         # two N tiles
    (32, 50, 21, 7, 17, 32, 64),          # 4 vertices per tile
    (64, 30, 9, 5, 17, 32, 32),
    (128, 20, 5, 4, 17, 8, 64),
    (48, 33, 11, 6, 17, 24, 96),          # 2 vertices per tile, 32 idle rows; O = 96
    (16, 80, 70, 8, 9, 32, 64),           # M = 9: 3 k-steps per chunk, tail of one basis
    (16, 80, 70, 8, 4, 40, 64),           # M = 4: a single full k-step
    (16, 80, 70, 8, 1, 32, 32),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_P%d-%d_N%d_M%d_I%d_O%d" % c)
@pytest.mark.parametrize("variant", ["act_res_ppbias", "plain"])
def test_fused_forward(case, variant):
    B, p_in, p_out, n_max, M, I, O = case
    tab, ids = random_table(p_in, p_out, n_max, seed=B + p_out)
    t = DeviceTables(tab, p_in)
    L = _lib.lib()
    if not L.vcm_conv_fwd_fused_supported(t.handle, B, M, I, O, TF32):
        pytest.skip("fused kernel does not cover this shape")
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(B, p_in, I, device="cuda", generator=g)
    A = torch.randn(p_out, n_max, M, device="cuda", generator=g)
    W = torch.randn(M, O * I, device="cuda", generator=g) / np.sqrt(M * I)
    full = variant == "act_res_ppbias"
    bias = torch.randn(p_out, O, device="cuda", generator=g) if full else None
    res = torch.randn(B, p_out, O, device="cuda", generator=g) if full else None
    sy, sr = (0.3, 0.9) if full else (1.0, 0.0)
    out = torch.full((B, p_out, O), float("nan"), device="cuda")
    y = torch.full((B, p_out, O), float("nan"), device="cuda") if full else None
    z = torch.full((B, p_out, M * I), float("nan"), device="cuda")
    ws = torch.empty(max(L.vcm_conv_fwd_fused_workspace(t.handle, B, M, I, O, TF32), 1), device="cuda")
    ptr = lambda a: a.data_ptr() if a is not None else None
    for zbuf in (z, None):                                                # with and without the z store
        out.fill_(float("nan"))
        _lib.check(L.vcm_conv_fwd_fused(t.handle, ptr(x), ptr(A), ptr(W), ptr(bias), 1, ptr(res), ptr(y), ptr(out),
                                        ptr(zbuf), ptr(ws), B, M, I, O, int(full), sy, sr, TF32, None))
        torch.cuda.synchronize()
        z64 = reference(x, A, ids, W, M, I, O, p_in)
        Wk = W.view(M, O, I)
        def product(zz, ww):
            s = torch.einsum("bpmi,moi->bpo", zz, ww)
            if bias is not None:
                s = s + bias.double()
            if full:
                s = torch.where(s > 0, s, torch.expm1(s))
            return s * sy + (res.double() * sr if res is not None else 0)
        exact = product(z64, Wk.double())
        assert torch.isfinite(out).all()
        assert rel_err(out.double(), exact) < TF32_TOL
        if zbuf is not None:
            # the stand-alone K1 kernel runs the same fp32 FMA chain
            z1 = torch.empty_like(z)
            _lib.check(L.vcm_gather_contract_fwd(t.handle, ptr(x), ptr(A), ptr(z1), B, I, M, None))
            torch.cuda.synchronize()
            assert torch.isfinite(z).all()
            assert rel_err(z.double(), z64.reshape(B, p_out, M * I)) < 1e-6
            assert torch.equal(z, z1)
            trunc = product(trunc_tf32(z).view(B, p_out, M, I).double(), trunc_tf32(Wk).double())
            assert rel_err(out.double(), trunc) < EXACT_TOL
            if y is not None:
                pre = (out.double() - res.double() * sr) / sy
                assert rel_err(y.double(), pre) < 1e-5


def test_fused_autograd_matches_two_kernel_path():
    """Gradients through the fused forward (z stored by the fused kernel) equal those of gather_contract + basis_gemm."""
    B, p_in, p_out, n_max, M, I, O = 16, 120, 77, 9, 17, 32, 64
    tab, ids = random_table(p_in, p_out, n_max, seed=3)
    t = DeviceTables(tab, p_in)
    if not F_.fused_conv_supported(t, B, M, I, O, TF32):
        pytest.skip("fused kernel does not cover this shape")
    g = torch.Generator(device="cuda").manual_seed(11)
    mk = lambda *s: torch.randn(*s, device="cuda", generator=g).requires_grad_()
    x, A, W, bias, res = mk(B, p_in, I), mk(p_out, n_max, M), mk(M, O * I), mk(p_out, O), mk(B, p_out, O)
    dout = torch.randn(B, p_out, O, device="cuda", generator=g)
    o1 = F_.fused_conv(x, A, W, bias, res, t, M, I, O, True, 0.3, 0.9, TF32)
    g1 = torch.autograd.grad(o1, (x, A, W, bias, res), dout)
    o2 = F_.basis_gemm(F_.gather_contract(x, A, t), W, bias, res, M, I, O, True, 0.3, 0.9, TF32)
    g2 = torch.autograd.grad(o2, (x, A, W, bias, res), dout)
    assert rel_err(o1, o2) < EXACT_TOL
    for a, b in zip(g1, g2):
        # out differs by accumulation order; downstream of it dz, dA and dx go through TF32 products again
        assert rel_err(a, b) < TF32_TOL
