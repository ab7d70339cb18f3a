"""Parity of the CUDA path (through the C ABI) against the golden vectors of
the reference and against the CPU oracle. Run on the B200 box: -m gpu.

Tolerances (north_star): index tables and gather results bit-exact; fp32
forward/backward <= 1e-5 relative (norm-wise, vs the fp64 gold).
"""
import types

import numpy as np
import pytest
import torch

from helpers import layer_cases, load, oracle_layer_from_golden, rel_err
from oracle import vcconv_oracle as orc
from vcmeshconv_b200 import _lib, connectivity, functional as F_
from vcmeshconv_b200 import graph_sampling as gs
from vcmeshconv_b200 import graphVAESSW as vae
from vcmeshconv_b200 import synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda"


def _module_from_golden(g):
    per_point = g["param.bias"].ndim == 2
    layer = vae.LASMConvssw(int(g["in_channel"]), int(g["out_channel"]), int(g["weight_num"]), int(g["in_point_num"]),
                            g["table"], per_point, float(g["residual_rate"]))
    layer.load_state_dict({k[len("param."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")})
    return layer.to(DEV)


def test_library_is_the_cuda_one():
    assert _lib.lib().vcm_version() >= 100
    assert F_.get_precision() == "tf32x3"          # default: fp32-grade on tensor cores; meets the 1e-5 bar below


@pytest.mark.parametrize("case", ["layer_pool_plain", "layer_conv_res_same", "layer_tets_pool_res"])
def test_layer_in_cuda_core_fp32_mode(case):
    """Same golden comparison with the GEMMs forced onto the fp32 CUDA-core kernels."""
    prev = F_.set_precision("fp32")
    try:
        test_layer_forward_backward_vs_reference_gold(case)
    finally:
        F_.set_precision(prev)


@pytest.mark.parametrize("case", layer_cases())
def test_device_tables_bit_exact(case):
    g = load(case)
    p_in = int(g["in_point_num"])
    t = connectivity.DeviceTables(g["table"], p_in)
    want = connectivity.derive_host(g["table"], p_in)
    assert np.array_equal(t.export(_lib.TAB_IDX).reshape(want["idx"].shape), want["idx"])
    assert np.array_equal(t.export(_lib.TAB_IDX).reshape(want["idx"].shape), g["param.neighbor_id_lstlst"])
    assert np.array_equal(t.export(_lib.TAB_LAST), want["last"])
    assert np.array_equal(t.export(_lib.TAB_PERM), want["perm"])
    assert np.array_equal(t.export(_lib.TAB_REV_PTR), want["rev_ptr"])
    assert np.array_equal(t.export(_lib.TAB_REV_SLOT), want["rev_slot"])
    assert np.array_equal(t.export(_lib.TAB_IDX_SORTED).reshape(want["idx"].shape), want["idx_sorted"])
    assert np.array_equal(t.export(_lib.TAB_META_SORTED).reshape(-1, 2), want["meta_sorted"])
    assert t.nnz == int((g["param.neighbor_id_lstlst"] != p_in).sum())


@pytest.mark.parametrize("case", ["layer_conv_res_same", "layer_pool_res", "layer_unpool_final", "layer_conv_3to32"])
def test_gather_is_bit_exact(case):
    """With one-hot coefficients the contraction kernel is a pure gather: the
    result must equal the reference's zero-padded index_select bit for bit."""
    g = load(case)
    p_in = int(g["in_point_num"])
    ids = g["param.neighbor_id_lstlst"]
    t = connectivity.DeviceTables(g["table"], p_in)
    x = torch.from_numpy(g["x"]).to(DEV)
    xpad = np.concatenate([g["x"], np.zeros_like(g["x"][:, :1])], 1)
    for k in sorted({0, 1, ids.shape[1] // 2, ids.shape[1] - 1}):
        A = torch.zeros(ids.shape[0], ids.shape[1], 1, device=DEV)
        A[:, k, 0] = 1.0
        z = F_.gather_contract(x, A, t).cpu().numpy()
        assert np.array_equal(z, xpad[:, ids[:, k]]), k


@pytest.mark.parametrize("case", layer_cases())
def test_layer_forward_backward_vs_reference_gold(case):
    g = load(case)
    layer = _module_from_golden(g)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    A = torch.from_numpy(g["coeffs"]).to(DEV).requires_grad_(True)
    y = layer(x, A, is_final_layer=bool(g["is_final_layer"]))
    assert y.shape == g["y64"].shape and y.dtype == torch.float32 and y.is_contiguous()
    assert rel_err(y.detach().cpu().numpy(), g["y64"]) < TOL
    y.backward(torch.from_numpy(g["gy"]).to(DEV))
    assert rel_err(x.grad.cpu().numpy(), g["grad.x"]) < TOL
    assert rel_err(A.grad.cpu().numpy(), g["grad.coeffs"]) < TOL
    for k, p in layer.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), g["grad." + k]) < TOL, k
    # padded coefficient slots: exactly zero gradient (Adam then never moves them)
    pad = g["param.neighbor_id_lstlst"] == int(g["in_point_num"])
    assert np.all(A.grad.cpu().numpy()[pad] == 0)
    if "p_neighbors" in dict(layer.named_parameters()):
        assert np.all(layer.p_neighbors.grad.cpu().numpy()[pad] == 0)


@pytest.mark.parametrize("case", ["layer_pool_res", "layer_tets_pool_res", "layer_unpool_res_same"])
def test_backward_is_bitwise_deterministic(case):
    g = load(case)
    layer = _module_from_golden(g)
    outs = []
    for _ in range(3):
        layer.zero_grad(set_to_none=True)
        x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
        A = torch.from_numpy(g["coeffs"]).to(DEV).requires_grad_(True)
        layer(x, A).backward(torch.from_numpy(g["gy"]).to(DEV))
        outs.append([x.grad.clone(), A.grad.clone()] + [p.grad.clone() for p in layer.parameters()])
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert torch.equal(a, b)


def _random_case(seed, n_vert, levels, which, cin, cout, m, batch, resid, per_point, final, kind="sphere"):
    pts, faces = (synthetic.sphere if kind == "sphere" else synthetic.tets)(n_vert, seed=seed)
    tabs, nums = gs.build_tables(faces, n_vert, levels)
    table = tabs[which]
    lvl = int("".join(c for c in which if c.isdigit()))
    p_in = nums[lvl + 1] if which.startswith("unpool") else nums[lvl]
    gen = torch.Generator().manual_seed(seed)
    layer = orc.Layer(table, cin, cout, p_in, resid)
    params = orc.init_layer_params(table, cin, cout, m, p_in, per_point, resid, gen, torch.float64)
    params["bias"] = torch.randn(params["bias"].shape, generator=gen, dtype=torch.float64) * 0.1
    coeffs = orc.init_coeffs(table, m, gen, torch.float64) * 4
    x = torch.randn(batch, p_in, cin, generator=gen, dtype=torch.float64) * 0.3
    gy = torch.randn(batch, table.shape[0], cout, generator=gen, dtype=torch.float64)
    return table, p_in, layer, params, coeffs, x, gy


RANDOM_CASES = [
    # seed n_vert levels               table      I    O   M  B  resid per_pt final kind
    (1, 1500, [(1, 1, 1), (2, 2, 2)], "pool1", 32, 64, 17, 4, 0.9, True, False, "sphere"),
    (2, 1500, [(1, 1, 1), (2, 2, 2)], "unpool1", 64, 32, 17, 3, 0.9, True, False, "sphere"),
    (3, 1200, [(1, 1, 1)], "pool0", 3, 32, 17, 16, 0.0, True, False, "sphere"),
    (4, 900, [(1, 1, 1), (2, 2, 2)], "unpool1", 32, 3, 17, 5, 0.9, True, True, "sphere"),
    (5, 700, [(1, 1, 1), (2, 2, 2), (1, 1, 1)], "pool2", 128, 128, 17, 2, 0.9, True, False, "sphere"),
    (6, 400, [(1, 1, 1)], "pool0", 5, 7, 2, 1, 0.0, False, False, "tets"),       # odd sizes, scalar path
    (7, 400, [(1, 1, 1), (2, 2, 2)], "pool1", 6, 10, 23, 2, 0.9, True, False, "tets"),   # M > 17, N ~ 100
    (8, 300, [(1, 1, 1), (2, 1, 1)], "unpool1", 16, 24, 1, 3, 0.9, False, False, "sphere"),  # M = 1, rows with 1 nbr
    (9, 2500, [(1, 1, 1), (2, 2, 2), (1, 1, 1), (2, 2, 2)], "pool3", 256, 512, 17, 2, 0.0, True, False, "sphere"),
]


@pytest.mark.parametrize("cfg", RANDOM_CASES, ids=lambda c: "seed%d_%s_%dto%d_M%d" % (c[0], c[3], c[4], c[5], c[6]))
def test_layer_vs_cpu_oracle(cfg):
    seed, n_vert, levels, which, cin, cout, m, batch, resid, per_point, final, kind = cfg
    table, p_in, olayer, params, coeffs, x, gy = _random_case(seed, n_vert, levels, which, cin, cout, m, batch, resid,
                                                              per_point, final, kind)
    leaves = [x, coeffs] + list(params.values())
    for t in leaves:
        t.requires_grad_(True)
    y_ref = olayer(x, coeffs, params, is_final_layer=final)
    y_ref.backward(gy)

    layer = vae.LASMConvssw(cin, cout, m, p_in, table, per_point, resid)
    layer.load_state_dict({**{k: v.detach().float() for k, v in params.items()},
                           "neighbor_num_lst": layer.neighbor_num_lst, "neighbor_id_lstlst": layer.neighbor_id_lstlst})
    layer = layer.to(DEV)
    xg = x.detach().float().to(DEV).requires_grad_(True)
    Ag = coeffs.detach().float().to(DEV).requires_grad_(True)
    y = layer(xg, Ag, is_final_layer=final)
    y.backward(gy.float().to(DEV))
    assert rel_err(y.detach().cpu().numpy(), y_ref.detach().numpy()) < TOL
    assert rel_err(xg.grad.cpu().numpy(), x.grad.numpy()) < TOL
    assert rel_err(Ag.grad.cpu().numpy(), coeffs.grad.numpy()) < TOL
    for k, p in layer.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), params[k].grad.numpy()) < TOL, k


def test_edge_cases_empty_rows_and_middle_sentinels():
    # row 1 has no neighbour at all, row 2 has a sentinel between two real entries
    p_in = 5
    table = np.array([[3, 0, 0, 1, 1, 4, 1],
                      [0, 5, -1, 5, -1, 5, -1],
                      [2, 2, 0, 5, -1, 3, 1],
                      [1, 4, 0, 5, -1, 5, -1]], np.int32)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(2, p_in, 8, generator=gen, dtype=torch.float64)
    A = torch.randn(4, 3, 17, generator=gen, dtype=torch.float64)
    W = torch.randn(17, 8 * 8, generator=gen, dtype=torch.float64)
    b = torch.randn(4, 8, generator=gen, dtype=torch.float64)
    for t in (x, A, W, b):
        t.requires_grad_(True)
    _, ids = orc.parse_table(table)
    y_ref = orc.vcconv_forward(x, A, W, b, ids, p_in)
    y_ref.sum().backward()
    t = connectivity.DeviceTables(table, p_in)
    xg, Ag, Wg, bg = [v.detach().float().to(DEV).requires_grad_(True) for v in (x, A, W, b)]
    y = F_.vcconv(xg, Ag, Wg, bg, t)
    y.sum().backward()
    assert rel_err(y.detach().cpu().numpy(), y_ref.detach().numpy()) < TOL
    for got, want in ((xg, x), (Ag, A), (Wg, W), (bg, b)):
        assert rel_err(got.grad.cpu().numpy(), want.grad.numpy()) < TOL
    assert np.all(Ag.grad.cpu().numpy()[ids.numpy() == p_in] == 0)
    # a row without neighbours is ELU(bias) (expm1f on the device vs ATen's CPU exp: 1 ulp apart)
    want = torch.nn.functional.elu(b[1].detach()).numpy()
    assert np.allclose(y.detach().cpu().numpy()[:, 1], want[None], rtol=1e-6, atol=1e-7)


def test_shape_and_device_errors():
    g = load("layer_conv_res_same")
    layer = _module_from_golden(g)
    x = torch.from_numpy(g["x"]).to(DEV)
    A = torch.from_numpy(g["coeffs"]).to(DEV)
    with pytest.raises(RuntimeError):
        layer(x[:, :-1], A)                       # wrong vertex count
    with pytest.raises(RuntimeError):
        layer(x, A[:, :-1])                       # wrong coefficient table
    with pytest.raises(RuntimeError):
        layer(x.double(), A)                      # fp32 only
    with pytest.raises(RuntimeError):
        layer(x.cpu(), A)                         # no CPU path
    # host pointers are refused by the C ABI itself
    t = layer.tables(x.device)
    host = np.zeros(16, np.float32)
    rc = _lib.lib().vcm_gather_contract_fwd(t.handle, host.ctypes.data, A.data_ptr(), x.data_ptr(), 1, 4, 1, None)
    assert rc == -1 and b"device pointer" in _lib.lib().vcm_last_error()


def _stack_modules(g):
    param = types.SimpleNamespace(residual_rate=float(g["residual_rate"]), conv_max=0,
                                  connection_layer_fn_lst_enc=[g["table." + str(n)] for n in g["enc_names"]],
                                  connection_layer_fn_lst_dec=[g["table." + str(n)] for n in g["dec_names"]])
    t0 = g["table.pool0"]
    param.neighbor_id_lstlst = t0[:, 1:].reshape(t0.shape[0], -1, 2)[:, :, 0]
    param.neighbor_num_lst = t0[:, 0]
    m = int(g["weight_num"])
    s_enc = vae.MCStructure(param, int(g["n_vert"]), m, bDec=False)
    vce, enc = vae.MCVcoeffs(s_enc, m), vae.MCEnc(s_enc, [int(c) for c in g["enc_channels"]], m)
    s_dec = vae.MCStructure(param, enc.out_nrpts, m, bDec=True)
    vcd, dec = vae.MCVcoeffs(s_dec, m), vae.MCDec(s_dec, [int(c) for c in g["dec_channels"]], m)
    mods = {"enc": enc, "dec": dec, "vce": vce, "vcd": vcd}
    for tag, mod in mods.items():
        sd = mod.state_dict()
        mod.load_state_dict({k[len(tag) + 1:]: torch.from_numpy(v).to(sd[k[len(tag) + 1:]].dtype)
                             for k, v in g.items() if k.startswith(tag + ".")})
        mod.to(DEV)
    return mods, vae.MCLoss(param).to(DEV)


def test_autoencoder_stack_vs_reference_gold():
    """MCEnc -> reparameterisation -> MCDec -> losses (graphVAE_train.py:199-225)
    against the values and gradients the reference's own modules produced."""
    g = load("stack_sphere300")
    mods, loss_mod = _stack_modules(g)
    x = torch.from_numpy(g["x"]).float().to(DEV)
    nor = torch.from_numpy(g["normals"]).float().to(DEV)
    eps = torch.from_numpy(g["eps"]).float().to(DEV)
    faces = torch.from_numpy(g["faces"]).long().to(DEV)
    mu, logstd = mods["enc"](x, mods["vce"])
    z = mu + logstd.exp() * eps
    kl = torch.mean(-0.5 - logstd + 0.5 * mu ** 2 + 0.5 * torch.exp(2 * logstd))
    out = mods["dec"](z, mods["vcd"])[:, :, 0:3]
    dif = out - x
    l_nor = ((dif[:, faces[:, 0]] + dif[:, faces[:, 1]] + dif[:, faces[:, 2]]) / 3.0 * nor).sum(2).pow(2).mean()
    l_l1 = loss_mod.compute_geometric_loss_l1(x, out)
    l_lap = loss_mod.compute_laplace_loss_l2(x, out)
    loss = l_l1 * 1.0 + l_lap * 100.0 + kl * 1e-5 + l_nor * 10.0
    loss.backward()
    assert rel_err(mu.detach().cpu().numpy(), g["mu"]) < TOL
    assert rel_err(logstd.detach().cpu().numpy(), g["logstd"]) < TOL
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < TOL
    for got, key in ((loss, "loss"), (l_l1, "loss_l1"), (l_lap, "loss_laplace"), (kl, "loss_kl"), (l_nor, "loss_normal")):
        assert abs(got.item() - float(g[key])) <= 2e-5 * max(abs(float(g[key])), 1e-3), key
    worst = 0.0
    for tag, mod in mods.items():
        for k, p in mod.named_parameters():
            worst = max(worst, rel_err(p.grad.cpu().numpy(), g["grad.%s.%s" % (tag, k)]))
    assert worst < 5 * TOL, worst     # 12 chained layers through a Laplacian-weighted (x100) loss


def test_full_size_properties_c1():
    """BASELINE config C1 shape (6890-vertex SMPL-topology sphere, 3 -> 32, M = 17,
    B = 16): too slow for the fp64 oracle per test run, so check size-independent
    properties: linearity of the un-activated layer in x, batch independence
    (bit-exact) and exact zero response to the zero input."""
    pts, faces = synthetic.sphere(6890, seed=0)
    tabs, _ = gs.build_tables(faces, 6890, [(1, 1, 1)])
    torch.manual_seed(0)
    layer = vae.LASMConvssw(3, 32, 17, 6890, tabs["pool0"], True, 0.0).to(DEV)
    A = (torch.randn(6890, layer.max_neighbor_num, 17) / (7 * 17)).to(DEV)
    x1 = (torch.randn(16, 6890, 3) * 0.3).to(DEV)
    x2 = (torch.randn(16, 6890, 3) * 0.3).to(DEV)
    with torch.no_grad():
        f = lambda x: layer(x, A, is_final_layer=True)
        y1, y2, y12 = f(x1), f(x2), f(2.0 * x1 - 3.0 * x2)
        assert rel_err(y12.cpu().numpy(), (2.0 * y1 - 3.0 * y2).cpu().numpy()) < TOL
        assert torch.equal(f(x1[3:4]), y1[3:4])
        assert torch.count_nonzero(f(torch.zeros_like(x1))) == 0
    # a one-vertex perturbation only reaches the vertices whose table row lists it
    ids = layer.neighbor_id_lstlst.cpu().numpy()
    xp = x1.clone()
    xp[:, 1234] += 1.0
    with torch.no_grad():
        changed = (layer(xp, A) != layer(x1, A)).any(0).any(1).cpu().numpy()
    assert set(np.nonzero(changed)[0]) <= set(np.nonzero((ids == 1234).any(1))[0])
