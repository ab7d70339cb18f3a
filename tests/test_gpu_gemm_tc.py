"""tcgen05 / TMEM basis GEMMs (VCM_PREC_TF32) through the C ABI.

Two references per case:
  * exact:     fp64 product of the fp32 inputs            -> stated TF32 tolerance 3e-3 (norm-wise)
  * truncated: fp64 product of the inputs cut to TF32 (the tensor core ignores the low 13 mantissa
               bits of each fp32 operand) -> must agree to fp32-accumulation round-off (2e-5),
               which pins descriptors / swizzles / tile bookkeeping independently of TF32 rounding.
"""
import numpy as np
import pytest
import torch

from helpers import rel_err
from vcmeshconv_b200 import _lib

pytestmark = pytest.mark.gpu
TF32_TOL = 3e-3
EXACT_TOL = 2e-5
TF32 = _lib.PREC_TF32


def trunc_tf32(t):
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def _case(R, M, I, O, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    z = torch.randn(R, M * I, device="cuda", generator=g)
    W = torch.randn(M, O * I, device="cuda", generator=g)
    ds = torch.randn(R, O, device="cuda", generator=g)
    return z, W, ds


SHAPES = [  # R, M, I, O
    (256, 1, 32, 32),
    (128, 17, 32, 64),
    (1000, 17, 64, 128),
    (28976, 17, 32, 64),       # enc layer 1 of C2 (16 x 1811 rows)
    (5504, 17, 128, 256),
    (816, 17, 256, 512),       # two N tiles
    (816, 17, 512, 64),        # few rows, K = 8704: split reduction + second-pass epilogue
    (77, 3, 64, 96),           # rows < one tile, O not a power of two
    (300, 5, 96, 32),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "R%d_M%d_I%d_O%d" % s)
def test_forward_gemm(shape):
    R, M, I, O = shape
    if not _lib.lib().vcm_basis_uses_tensor_cores(R, M, I, O, TF32):
        pytest.skip("no tensor-core kernel for this shape")
    z, W, _ = _case(R, M, I, O, 1)
    P = R // 4 if R % 4 == 0 else R
    B = R // P
    bias = torch.randn(P, O, device="cuda")
    res = torch.randn(R, O, device="cuda")
    out = torch.full((R, O), float("nan"), device="cuda")
    y = torch.full((R, O), float("nan"), device="cuda")
    ws = torch.empty(max(_lib.lib().vcm_basis_fwd_workspace(R, M, I, O, TF32), 1), device="cuda")
    _lib.check(_lib.lib().vcm_basis_fwd(z.data_ptr(), W.data_ptr(), bias.data_ptr(), 1, res.data_ptr(), y.data_ptr(),
                                        out.data_ptr(), ws.data_ptr(), B, P, M, I, O, 1, 0.5, 2.0, TF32,
                                        _lib.stream_ptr()))
    torch.cuda.synchronize()
    Wk = W.view(M, O, I).permute(0, 2, 1).reshape(M * I, O)

    def ref(zz, ww):
        t = zz.double() @ ww.double() + bias.double().repeat(B, 1)
        yy = torch.nn.functional.elu(t)
        return yy, yy * 0.5 + res.double() * 2.0

    y_t, o_t = ref(trunc_tf32(z), trunc_tf32(Wk))
    y_e, o_e = ref(z, Wk)
    assert not torch.isnan(out).any() and not torch.isnan(y).any()
    assert rel_err(y.cpu().numpy(), y_t.cpu().numpy()) < EXACT_TOL
    assert rel_err(out.cpu().numpy(), o_t.cpu().numpy()) < EXACT_TOL
    assert rel_err(out.cpu().numpy(), o_e.cpu().numpy()) < TF32_TOL


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "R%d_M%d_I%d_O%d" % s)
def test_dz_gemm(shape):
    R, M, I, O = shape
    L = _lib.lib()
    if not L.vcm_basis_uses_tensor_cores(R, M, I, O, TF32):
        pytest.skip("no tensor-core kernel for this shape")
    _, W, ds = _case(R, M, I, O, 2)
    dz = torch.full((R, M * I), float("nan"), device="cuda")
    ws = torch.empty(max(L.vcm_basis_bwd_dz_workspace(R, M, I, O, TF32), 1), device="cuda")
    _lib.check(L.vcm_basis_bwd_dz(ds.data_ptr(), W.data_ptr(), dz.data_ptr(), ws.data_ptr(), R, M, I, O, 0, TF32,
                                  _lib.stream_ptr()))
    torch.cuda.synchronize()
    Wk = W.view(M, O, I).permute(0, 2, 1).reshape(M * I, O)
    want_t = trunc_tf32(ds).double() @ trunc_tf32(Wk).double().t()
    want_e = ds.double() @ Wk.double().t()
    assert not torch.isnan(dz).any()
    assert rel_err(dz.cpu().numpy(), want_t.cpu().numpy()) < EXACT_TOL
    assert rel_err(dz.cpu().numpy(), want_e.cpu().numpy()) < TF32_TOL


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "R%d_M%d_I%d_O%d" % s)
def test_dw_gemm(shape):
    R, M, I, O = shape
    L = _lib.lib()
    if not L.vcm_basis_uses_tensor_cores(R, M, I, O, TF32):
        pytest.skip("no tensor-core kernel for this shape")
    z, _, ds = _case(R, M, I, O, 3)
    dW = torch.full((M, O * I), float("nan"), device="cuda")
    ws = torch.empty(max(L.vcm_basis_bwd_dw_workspace(R, M, I, O, TF32), 1), device="cuda")
    _lib.check(L.vcm_basis_bwd_dw(ds.data_ptr(), z.data_ptr(), dW.data_ptr(), ws.data_ptr(), R, M, I, O, TF32,
                                  _lib.stream_ptr()))
    torch.cuda.synchronize()

    def ref(zz, dd):
        return (zz.double().t() @ dd.double()).view(M, I, O).permute(0, 2, 1).reshape(M, O * I)

    assert not torch.isnan(dW).any()
    assert rel_err(dW.cpu().numpy(), ref(trunc_tf32(z), trunc_tf32(ds)).cpu().numpy()) < EXACT_TOL
    assert rel_err(dW.cpu().numpy(), ref(z, ds).cpu().numpy()) < TF32_TOL
    # deterministic split reduction
    dW2 = torch.empty_like(dW)
    _lib.check(L.vcm_basis_bwd_dw(ds.data_ptr(), z.data_ptr(), dW2.data_ptr(), ws.data_ptr(), R, M, I, O, TF32,
                                  _lib.stream_ptr()))
    assert torch.equal(dW, dW2)


def test_layer_in_tf32_mode_meets_stated_tolerance():
    """Whole layer fwd+bwd with the tensor-core GEMMs vs the fp64 CPU oracle."""
    from oracle import vcconv_oracle as orc
    from vcmeshconv_b200 import functional as F_
    from vcmeshconv_b200 import graph_sampling as gs
    from vcmeshconv_b200 import graphVAESSW as vae
    from vcmeshconv_b200 import synthetic
    pts, faces = synthetic.sphere(1500, seed=1)
    tabs, nums = gs.build_tables(faces, 1500, [(1, 1, 1), (2, 2, 2)])
    table, p_in, cin, cout, m = tabs["pool1"], nums[1], 32, 64, 17
    gen = torch.Generator().manual_seed(1)
    layer_o = orc.Layer(table, cin, cout, p_in, 0.9)
    params = orc.init_layer_params(table, cin, cout, m, p_in, True, 0.9, gen, torch.float64)
    coeffs = orc.init_coeffs(table, m, gen, torch.float64) * 4
    x = torch.randn(4, p_in, cin, generator=gen, dtype=torch.float64) * 0.3
    gy = torch.randn(4, table.shape[0], cout, generator=gen, dtype=torch.float64)
    for t in [x, coeffs] + list(params.values()):
        t.requires_grad_(True)
    y_ref = layer_o(x, coeffs, params)
    y_ref.backward(gy)
    layer = vae.LASMConvssw(cin, cout, m, p_in, table, True, 0.9)
    layer.load_state_dict({**{k: v.detach().float() for k, v in params.items()},
                           "neighbor_num_lst": layer.neighbor_num_lst, "neighbor_id_lstlst": layer.neighbor_id_lstlst})
    layer = layer.cuda()
    prev = F_.set_precision("tf32")
    try:
        xg = x.detach().float().cuda().requires_grad_(True)
        Ag = coeffs.detach().float().cuda().requires_grad_(True)
        y = layer(xg, Ag)
        y.backward(gy.float().cuda())
    finally:
        F_.set_precision(prev)
    assert rel_err(y.detach().cpu().numpy(), y_ref.detach().numpy()) < TF32_TOL
    assert rel_err(xg.grad.cpu().numpy(), x.grad.numpy()) < TF32_TOL
    assert rel_err(Ag.grad.cpu().numpy(), coeffs.grad.numpy()) < TF32_TOL
    for k, p in layer.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), params[k].grad.numpy()) < TF32_TOL, k


@pytest.mark.parametrize("cfg", [("pool1", 32, 17, 4), ("unpool1", 64, 17, 3), ("pool0", 128, 17, 2), ("pool1", 48, 5, 2),
                                 ("unpool1", 16, 24, 1)], ids=lambda c: "%s_I%d_M%d_B%d" % c)
def test_gather_contract_on_tensor_pipe(cfg):
    """K1 in TF32 mode (mma.sync m16n8k8): exact vs the fp64 contraction of TF32-truncated operands, and within
    the stated TF32 tolerance of the fp32 kernel."""
    from vcmeshconv_b200 import connectivity
    from vcmeshconv_b200 import graph_sampling as gs
    from vcmeshconv_b200 import synthetic
    import os
    if os.environ.get("VCM_K1_MMA", "0") == "0":
        pytest.skip("tensor-pipe K1 is opt-in (VCM_K1_MMA=1); the fp32 kernel is faster in the full step")
    which, I, M, B = cfg
    pts, faces = synthetic.sphere(1500, seed=2)
    tabs, nums = gs.build_tables(faces, 1500, [(1, 1, 1), (2, 2, 2)])
    table = tabs[which]
    p_in = nums[2] if which.startswith("unpool") else nums[int(which[-1])]
    t = connectivity.DeviceTables(table, p_in)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, p_in, I, device="cuda", generator=g)
    A = torch.randn(t.p_out, t.n_max, M, device="cuda", generator=g)
    z_tc = torch.full((B, t.p_out, M * I), float("nan"), device="cuda")
    z_32 = torch.empty_like(z_tc)
    L = _lib.lib()
    _lib.check(L.vcm_gather_contract_fwd_ex(t.handle, x.data_ptr(), A.data_ptr(), z_tc.data_ptr(), B, I, M, TF32,
                                            _lib.stream_ptr()))
    _lib.check(L.vcm_gather_contract_fwd(t.handle, x.data_ptr(), A.data_ptr(), z_32.data_ptr(), B, I, M,
                                         _lib.stream_ptr()))
    torch.cuda.synchronize()
    ids = torch.from_numpy(table[:, 1::2].astype("int64")).cuda()
    mask = (ids != p_in).double()
    xpad = torch.cat([trunc_tf32(x).double(), torch.zeros(B, 1, I, device="cuda", dtype=torch.float64)], 1)
    want = torch.einsum("pnm,bpni->bpmi", trunc_tf32(A).double() * mask[:, :, None], xpad[:, ids]).reshape(B, t.p_out, M * I)
    assert not torch.isnan(z_tc).any()
    assert rel_err(z_tc.cpu().numpy(), want.cpu().numpy()) < EXACT_TOL
    assert rel_err(z_tc.cpu().numpy(), z_32.cpu().numpy()) < TF32_TOL


X3 = _lib.PRECISIONS["tf32x3"]
X3_TOL = 1e-5        # the fp32 bar of the north_star, met on tensor cores by the 3-pass hi/lo split


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "R%d_M%d_I%d_O%d" % s)
def test_tf32x3_gemms_are_fp32_grade(shape):
    R, M, I, O = shape
    L = _lib.lib()
    if not L.vcm_basis_uses_tensor_cores(R, M, I, O, X3):
        pytest.skip("no tensor-core kernel for this shape")
    z, W, ds = _case(R, M, I, O, 7)
    Wk = W.view(M, O, I).permute(0, 2, 1).reshape(M * I, O).double()
    st = _lib.stream_ptr()
    out = torch.full((R, O), float("nan"), device="cuda")
    ws = torch.empty(max(L.vcm_basis_fwd_workspace(R, M, I, O, X3), L.vcm_basis_bwd_dz_workspace(R, M, I, O, X3),
                         L.vcm_basis_bwd_dw_workspace(R, M, I, O, X3), 1), device="cuda")
    _lib.check(L.vcm_basis_fwd(z.data_ptr(), W.data_ptr(), None, 0, None, None, out.data_ptr(), ws.data_ptr(), 1, R, M, I,
                               O, 0, 1.0, 0.0, X3, st))
    assert rel_err(out.cpu().numpy(), (z.double() @ Wk).cpu().numpy()) < X3_TOL
    dz = torch.full((R, M * I), float("nan"), device="cuda")
    _lib.check(L.vcm_basis_bwd_dz(ds.data_ptr(), W.data_ptr(), dz.data_ptr(), ws.data_ptr(), R, M, I, O, 0, X3, st))
    assert rel_err(dz.cpu().numpy(), (ds.double() @ Wk.t()).cpu().numpy()) < X3_TOL
    dW = torch.full((M, O * I), float("nan"), device="cuda")
    _lib.check(L.vcm_basis_bwd_dw(ds.data_ptr(), z.data_ptr(), dW.data_ptr(), ws.data_ptr(), R, M, I, O, X3, st))
    want = (z.double().t() @ ds.double()).view(M, I, O).permute(0, 2, 1).reshape(M, O * I)
    assert rel_err(dW.cpu().numpy(), want.cpu().numpy()) < X3_TOL


def test_layer_in_tf32x3_mode_meets_the_fp32_bar():
    from oracle import vcconv_oracle as orc
    from vcmeshconv_b200 import functional as F_
    from vcmeshconv_b200 import graph_sampling as gs
    from vcmeshconv_b200 import graphVAESSW as vae
    from vcmeshconv_b200 import synthetic
    pts, faces = synthetic.sphere(1500, seed=1)
    tabs, nums = gs.build_tables(faces, 1500, [(1, 1, 1), (2, 2, 2)])
    table, p_in, cin, cout, m = tabs["unpool1"], nums[2], 64, 32, 17
    gen = torch.Generator().manual_seed(3)
    layer_o = orc.Layer(table, cin, cout, p_in, 0.9)
    params = orc.init_layer_params(table, cin, cout, m, p_in, True, 0.9, gen, torch.float64)
    coeffs = orc.init_coeffs(table, m, gen, torch.float64) * 4
    x = torch.randn(3, p_in, cin, generator=gen, dtype=torch.float64) * 0.3
    gy = torch.randn(3, table.shape[0], cout, generator=gen, dtype=torch.float64)
    for t in [x, coeffs] + list(params.values()):
        t.requires_grad_(True)
    y_ref = layer_o(x, coeffs, params)
    y_ref.backward(gy)
    layer = vae.LASMConvssw(cin, cout, m, p_in, table, True, 0.9)
    layer.load_state_dict({**{k: v.detach().float() for k, v in params.items()},
                           "neighbor_num_lst": layer.neighbor_num_lst, "neighbor_id_lstlst": layer.neighbor_id_lstlst})
    layer = layer.cuda()
    prev = F_.set_precision("tf32x3")
    try:
        xg = x.detach().float().cuda().requires_grad_(True)
        Ag = coeffs.detach().float().cuda().requires_grad_(True)
        y = layer(xg, Ag)
        y.backward(gy.float().cuda())
    finally:
        F_.set_precision(prev)
    assert rel_err(y.detach().cpu().numpy(), y_ref.detach().numpy()) < X3_TOL
    assert rel_err(xg.grad.cpu().numpy(), x.grad.numpy()) < X3_TOL
    assert rel_err(Ag.grad.cpu().numpy(), coeffs.grad.numpy()) < X3_TOL
    for k, p in layer.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), params[k].grad.numpy()) < X3_TOL, k
