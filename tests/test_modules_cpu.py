"""Host-side mirror of the reference module API: constructor signatures,
attributes, state_dict layout, initialiser parity, and the 'no CPU fallback'
rule. CPU only."""
import numpy as np
import pytest
import torch

from helpers import layer_cases, load
from vcmeshconv_b200 import graphVAESSW as vae


def _make(g, seed=None):
    if seed is not None:
        torch.manual_seed(seed)
    per_point = g["param.bias"].ndim == 2
    return vae.LASMConvssw(int(g["in_channel"]), int(g["out_channel"]), int(g["weight_num"]), int(g["in_point_num"]),
                           g["table"], per_point, float(g["residual_rate"]))


@pytest.mark.parametrize("case", layer_cases())
def test_state_dict_layout_matches_reference(case):
    g = load(case)
    layer = _make(g)
    sd = layer.state_dict()
    want = {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}
    assert list(sd.keys()) == list(want.keys())                 # same names, same registration order
    for k, v in want.items():
        assert tuple(sd[k].shape) == v.shape, k
        assert str(sd[k].dtype).replace("torch.", "") == str(v.dtype), k
    assert np.array_equal(sd["neighbor_id_lstlst"].numpy(), want["neighbor_id_lstlst"])
    assert np.array_equal(sd["neighbor_num_lst"].numpy(), want["neighbor_num_lst"])
    # a reference checkpoint loads as is
    layer.load_state_dict({k: torch.from_numpy(v) for k, v in want.items()})
    for attr in ("in_channel", "out_channel", "weight_num", "in_point_num", "out_point_num", "max_neighbor_num",
                 "avg_neighbor_num", "residual_rate", "relu"):
        assert hasattr(layer, attr)


SEEDS = {"layer_conv_3to32": 10, "layer_conv_res_same": 11, "layer_conv_res_proj": 12, "layer_pool_res": 13,
         "layer_pool_plain": 14, "layer_unpool_res_same": 15, "layer_unpool_final": 16,
         "layer_unpool_sharedbias": 17, "layer_tets_conv": 18, "layer_tets_pool_res": 19}


@pytest.mark.parametrize("case", layer_cases())
def test_same_seed_same_initial_parameters_as_reference(case):
    g = load(case)
    layer = _make(g, SEEDS[case])
    for k in ("weights", "p_neighbors", "weight_res"):
        if "param." + k in g:
            assert np.array_equal(getattr(layer, k).detach().numpy(), g["param." + k]), k


def test_cpu_tensors_are_rejected():
    g = load("layer_conv_res_same")
    layer = _make(g)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.from_numpy(g["x"]), torch.from_numpy(g["coeffs"]))
    with pytest.raises(NotImplementedError):
        layer(torch.from_numpy(g["x"]), torch.from_numpy(g["coeffs"]), b_max_pool=True)


def test_stack_containers():
    import types
    g = load("stack_sphere300")
    param = types.SimpleNamespace(residual_rate=float(g["residual_rate"]), conv_max=0,
                                  connection_layer_fn_lst_enc=[g["table." + str(n)] for n in g["enc_names"]],
                                  connection_layer_fn_lst_dec=[g["table." + str(n)] for n in g["dec_names"]])
    m = int(g["weight_num"])
    s_enc = vae.MCStructure(param, int(g["n_vert"]), m, bDec=False)
    vc_enc = vae.MCVcoeffs(s_enc, m)
    enc = vae.MCEnc(s_enc, [int(c) for c in g["enc_channels"]], m)
    s_dec = vae.MCStructure(param, enc.out_nrpts, m, bDec=True)
    vc_dec = vae.MCVcoeffs(s_dec, m)
    dec = vae.MCDec(s_dec, [int(c) for c in g["dec_channels"]], m)
    for tag, mod in (("enc", enc), ("dec", dec), ("vce", vc_enc), ("vcd", vc_dec)):
        want = {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + ".")}
        sd = mod.state_dict()
        assert list(sd.keys()) == list(want.keys()), tag
        for k, v in want.items():
            assert tuple(sd[k].shape) == v.shape, (tag, k)
        mod.load_state_dict({k: torch.from_numpy(v).to(sd[k].dtype) for k, v in want.items()})
    assert s_enc.ptnum_list[0] == 300 and enc.out_nrchs == int(g["enc_channels"][-1])
