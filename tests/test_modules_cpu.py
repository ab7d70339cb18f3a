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
         "layer_unpool_sharedbias": 17, "layer_tets_conv": 18, "layer_tets_pool_res": 19, "layer_conv_tc_res": 21}


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


def test_trainer_segments_follow_gradient_completion_order():
    """DataParallelTrainer lays the flat buffers out in the order in which gradients become final during the
    backward pass (decoder side | late encoder layers + heads | first encoder layers): every trainable parameter
    in exactly one segment, relative order inside a segment unchanged, and the hooked encoder layer index valid."""
    from vcmeshconv_b200 import graph_sampling, synthetic
    from vcmeshconv_b200 import graphVAE_train as T
    pts, faces = synthetic.sphere(600, seed=0)
    tabs, _ = graph_sampling.build_tables(faces, 600, synthetic.alternating_levels(5))
    enc = [tabs["pool%d" % l] for l in range(1, 4)]
    dec = [tabs["unpool%d" % l] for l in range(3, 0, -1)]
    model = T.Net_autoenc(T.make_param(enc, dec, tabs["pool0"], 600, 0.9), faces, enc_channels=[3, 16, 32, 8],
                          dec_channels=[8, 32, 16, 3], weight_num=9)
    params = model.trainable_parameters()
    segs, k = T.DataParallelTrainer._split_params(model)
    assert len(segs) == 3 and k == 2
    flat = [p for s in segs for p in s]
    assert sorted(map(id, flat)) == sorted(map(id, params)) and len(set(map(id, flat))) == len(params)
    order = {id(p): i for i, p in enumerate(params)}
    for s in segs:
        assert [order[id(p)] for p in s] == sorted(order[id(p)] for p in s)
    dec_ids = {id(p) for p in model.net_geodec.parameters()} | {id(p) for p in model.mcvcoeffsdec.parameters()}
    assert {id(p) for p in segs[0]} == dec_ids
    enc_m = model.net_geoenc
    tail = {id(p) for i in range(k) for p in enc_m.layer_lst[i].parameters()} | \
           {id(model.mcvcoeffsenc.vcoeffs_list[i]) for i in range(k)}
    assert {id(p) for p in segs[2]} == tail
    late = {id(p) for p in segs[1]}
    assert {id(p) for p in enc_m.stdlayer.parameters()} <= late and id(model.mcvcoeffsenc.vcoeffs_list[k]) in late
    # a model without the autoencoder structure is one segment
    class Plain(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(3))

        def trainable_parameters(self):
            return [self.w]
    segs1, k1 = T.DataParallelTrainer._split_params(Plain())
    assert len(segs1) == 1 and k1 is None
