"""Pins oracle/vcconv_oracle.py against the golden vectors recorded from the
UNMODIFIED reference (oracle/gen_golden.py). CPU only."""
import numpy as np
import pytest
import torch

from helpers import layer_cases, load, oracle_layer_from_golden, rel_err
from oracle import vcconv_oracle as orc


@pytest.mark.parametrize("case", layer_cases())
def test_layer_forward_backward_fp64(case):
    g = load(case)
    layer, A, params, x, gy = oracle_layer_from_golden(g, torch.float64)
    x.requires_grad_(True)
    A.requires_grad_(True)
    for p in params.values():
        p.requires_grad_(True)
    y = layer(x, A, params, is_final_layer=bool(g["is_final_layer"]))
    assert rel_err(y.detach().numpy(), g["y64"]) < 1e-13
    y.backward(gy)
    assert rel_err(x.grad.numpy(), g["grad.x"]) < 1e-12
    assert rel_err(A.grad.numpy(), g["grad.coeffs"]) < 1e-12
    for k, p in params.items():
        assert rel_err(p.grad.numpy(), g["grad." + k]) < 1e-12, k
    # padded coefficient slots get exactly zero gradient
    pad = layer.ids.numpy() == layer.in_point_num
    assert np.all(A.grad.numpy()[pad] == 0)


@pytest.mark.parametrize("case", layer_cases())
def test_layer_fp32_matches_reference_fp32(case):
    g = load(case)
    layer, A, params, x, _ = oracle_layer_from_golden(g, torch.float32)
    with torch.no_grad():
        y = layer(x, A, params, is_final_layer=bool(g["is_final_layer"]))
    # same ATen op sequence as the reference: agreement to fp32 round-off of the reductions
    assert rel_err(y.numpy(), g["y32_forward"]) < 2e-6
    assert rel_err(y.numpy(), g["y64"]) < 1e-5


@pytest.mark.parametrize("case", [c for c in layer_cases() if float(load(c)["residual_rate"]) == 0.0])
def test_second_formulation(case):
    """forward2 (per-edge kernels) is an independent formulation of the same maths."""
    g = load(case)
    layer, A, params, x, _ = oracle_layer_from_golden(g, torch.float64)
    y2 = orc.vcconv_forward_edge_kernels(x, A, params["weights"], params["bias"], layer.ids, layer.in_point_num,
                                         is_final_layer=bool(g["is_final_layer"]))
    assert rel_err(y2.numpy(), g["y64_forward2"]) < 1e-13
    assert rel_err(y2.numpy(), g["y64"]) < 1e-12


def test_parse_table_matches_reference_buffers():
    for case in layer_cases():
        g = load(case)
        cnt, ids = orc.parse_table(g["table"])
        assert np.array_equal(ids.numpy(), g["param.neighbor_id_lstlst"])
        assert np.array_equal(cnt.numpy(), g["param.neighbor_num_lst"])
        assert ids.dtype == torch.int64 and cnt.dtype == torch.float32


def _stack_from_golden(g, dtype=torch.float64):
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)
    res = float(g["residual_rate"])

    def build(tag, vtag, names, channels, n_in, with_std):
        layers, coeffs, params = [], [], []
        for l, nm in enumerate(names):
            tab = g["table." + str(nm)]
            layers.append(orc.Layer(tab, int(channels[l]), int(channels[l + 1]), n_in, res))
            coeffs.append(t(g["%s.vcoeffs_list.%d" % (vtag, l)]))
            pre = "%s.layer_lst.%d." % (tag, l)
            params.append({k[len(pre):]: t(v) for k, v in g.items() if k.startswith(pre) and "neighbor_" not in k})
            n_in = layers[-1].out_point_num
        std = None
        if with_std:
            pre = "%s.stdlayer." % tag
            std = {k[len(pre):]: t(v) for k, v in g.items() if k.startswith(pre) and "neighbor_" not in k}
        return layers, coeffs, params, std, n_in

    el, ec, ep, sp, n_lat = build("enc", "vce", g["enc_names"], g["enc_channels"], int(g["n_vert"]), True)
    dl, dc, dp, _, _ = build("dec", "vcd", g["dec_names"], g["dec_channels"], n_lat, False)
    return (el, ec, ep, sp), (dl, dc, dp)


def test_stack_and_losses_fp64():
    g = load("stack_sphere300")
    enc, dec = _stack_from_golden(g)
    leaves = list(enc[1]) + list(dec[1]) + [p for d in enc[2] + dec[2] + [enc[3]] for p in d.values()]
    for p in leaves:
        p.requires_grad_(True)
    x = torch.from_numpy(g["x"])
    cnt0, ids0 = orc.parse_table(g["table.pool0"])
    loss, l1, lap, kl, nor, out = orc.autoencoder_losses(
        x, torch.from_numpy(g["normals"]), torch.from_numpy(g["faces"]).long(), torch.from_numpy(g["eps"]),
        enc, dec, ids0, cnt0)
    mu, logstd = orc.encoder_forward(x, *enc)
    assert rel_err(mu.detach().numpy(), g["mu"]) < 1e-12
    assert rel_err(logstd.detach().numpy(), g["logstd"]) < 1e-12
    assert rel_err(out.detach().numpy(), g["out"]) < 1e-12
    for got, key in ((loss, "loss"), (l1, "loss_l1"), (lap, "loss_laplace"), (kl, "loss_kl"), (nor, "loss_normal")):
        assert abs(got.item() - float(g[key])) <= 1e-12 * max(1.0, abs(float(g[key]))), key
    loss.backward()
    assert rel_err(enc[1][0].grad.numpy(), g["grad.vce.vcoeffs_list.0"]) < 1e-11
    assert rel_err(dec[1][2].grad.numpy(), g["grad.vcd.vcoeffs_list.2"]) < 1e-11
    assert rel_err(enc[2][1]["weights"].grad.numpy(), g["grad.enc.layer_lst.1.weights"]) < 1e-11
    assert rel_err(enc[3]["weights"].grad.numpy(), g["grad.enc.stdlayer.weights"]) < 1e-11
    assert rel_err(dec[2][0]["p_neighbors"].grad.numpy(), g["grad.dec.layer_lst.0.p_neighbors"]) < 1e-11
    assert rel_err(dec[2][2]["weight_res"].grad.numpy(), g["grad.dec.layer_lst.2.weight_res"]) < 1e-11
