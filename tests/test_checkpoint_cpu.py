"""Optimizer-state interchange with the reference's checkpoints (SURVEY §8f N-4,
graphVAE_train.py:284-291, 339-347), host side only: the Adam state travels in
torch.optim.Adam's own state_dict layout over the getoptim parameter order.

The fixture tests/golden/ref_checkpoint_sphere300.weight was written by the
UNMODIFIED reference modules + torch.optim.Adam (oracle/gen_golden.py)."""
import itertools
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load
from vcmeshconv_b200 import graphVAE_train as T


def model_from_fixture(g):
    enc = [g["table." + str(n)] for n in g["enc_names"]]
    dec = [g["table." + str(n)] for n in g["dec_names"]]
    param = T.make_param(enc, dec, g["table.pool0"], int(g["n_vert"]), float(g["residual_rate"]))
    torch.manual_seed(123)
    return T.Net_autoenc(param, g["faces"], [int(c) for c in g["enc_channels"]], [int(c) for c in g["dec_channels"]],
                         int(g["weight_num"]))


def getoptim_order(model):
    """graphVAE_train.py:252-262."""
    return [p for p in itertools.chain(model.net_geodec.parameters(), model.net_geoenc.parameters(),
                                       model.mcvcoeffsdec.parameters(), model.mcvcoeffsenc.parameters())
            if p.requires_grad]


def test_reference_checkpoint_loads_with_its_adam_state():
    g = load("ref_checkpoint_sphere300")
    model = model_from_fixture(g)
    tr = T.DataParallelTrainer(model, lr=1e-4)
    ck = tr.load(os.path.join(GOLDEN, "ref_checkpoint_sphere300.weight"))
    ref_opt = ck["optimizer_state_dict"]
    assert tr.step_count == 2 and abs(tr.lr - float(g["lr"])) < 1e-12
    order = getoptim_order(model)
    assert [id(p) for p in order] == [id(p) for p in model.trainable_parameters()]
    by_id = {id(p): o for p, o in zip(tr.flat.params, tr.flat.offsets)}
    n_checked = 0
    for i, p in enumerate(order):
        st = ref_opt["state"][i]
        o, n = by_id[id(p)], p.numel()
        assert torch.equal(tr.exp_avg[o:o + n].view(p.shape), st["exp_avg"])
        assert torch.equal(tr.exp_avg_sq[o:o + n].view(p.shape), st["exp_avg_sq"])
        n_checked += n
    assert n_checked == sum(p.numel() for p in order)
    # module parameters came over as well
    for (k, v), (_, w) in zip(model.net_geoenc.state_dict().items(), ck["encgeo_state_dict"].items()):
        assert torch.equal(v, w), k


def test_our_checkpoint_is_accepted_by_torch_adam_like_the_reference(tmp_path):
    """What the reference does on resume (graphVAE_train.py:280-291): build Adam over getoptim's parameters, then
    optimizer.load_state_dict(checkpoint['optimizer_state_dict']). Our file must survive exactly that and carry
    the same moments / step / lr back."""
    g = load("ref_checkpoint_sphere300")
    model = model_from_fixture(g)
    tr = T.DataParallelTrainer(model, lr=3e-4)
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        tr.exp_avg.copy_(torch.randn(tr.exp_avg.shape, generator=gen))
        tr.exp_avg_sq.copy_(torch.rand(tr.exp_avg_sq.shape, generator=gen))
        tr.opt_state[0] = 41.0
    path = tr.save(str(tmp_path), 41)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"encgeo_state_dict", "decgeo_state_dict", "mcvcoeffsdec_dict", "mcvcoeffsenc_dict",
                       "optimizer_state_dict"}
    model2 = model_from_fixture(g)
    model2.net_geoenc.load_state_dict(ck["encgeo_state_dict"])
    model2.net_geodec.load_state_dict(ck["decgeo_state_dict"])
    model2.mcvcoeffsenc.load_state_dict(ck["mcvcoeffsenc_dict"])
    model2.mcvcoeffsdec.load_state_dict(ck["mcvcoeffsdec_dict"])
    params2 = getoptim_order(model2)
    opt = torch.optim.Adam(params2, lr=1e-4, betas=(0.9, 0.999))
    opt.load_state_dict(ck["optimizer_state_dict"])
    assert opt.param_groups[0]["lr"] == pytest.approx(3e-4)
    by_id = {id(p): o for p, o in zip(tr.flat.params, tr.flat.offsets)}
    for p_ours, p2 in zip(getoptim_order(model), params2):
        st = opt.state[p2]
        o, n = by_id[id(p_ours)], p_ours.numel()
        assert float(st["step"]) == 41.0
        assert torch.equal(st["exp_avg"].reshape(-1), tr.exp_avg[o:o + n])
        assert torch.equal(st["exp_avg_sq"].reshape(-1), tr.exp_avg_sq[o:o + n])
    # one torch-Adam step from the loaded state moves the parameters as Adam with t = 42 must
    for p in params2:
        p.grad = torch.ones_like(p)
    before = [p.detach().clone() for p in params2]
    opt.step()
    p0, b0 = params2[0], before[0]
    o = by_id[id(getoptim_order(model)[0])]
    m = 0.9 * tr.exp_avg[o:o + p0.numel()].view(p0.shape) + 0.1
    v = 0.999 * tr.exp_avg_sq[o:o + p0.numel()].view(p0.shape) + 0.001
    want = b0 - 3e-4 * (m / (1 - 0.9 ** 42)) / ((v / (1 - 0.999 ** 42)).sqrt() + 1e-8)
    assert torch.allclose(p0.detach(), want, rtol=1e-5, atol=1e-7)
    # and back: the trainer reads what torch's Adam writes
    tr2 = T.DataParallelTrainer(model_from_fixture(g), lr=1e-4)
    tr2.load_optimizer_state_dict(opt.state_dict())
    assert tr2.step_count == 42


def test_mismatched_optimizer_state_raises():
    g = load("ref_checkpoint_sphere300")
    tr = T.DataParallelTrainer(model_from_fixture(g))
    sd = tr.optimizer_state_dict()
    sd["param_groups"][0]["params"] = sd["param_groups"][0]["params"][:-1]
    with pytest.raises(RuntimeError, match="parameters"):
        tr.load_optimizer_state_dict(sd)
    with pytest.raises(RuntimeError, match="unknown"):
        tr.load_optimizer_state_dict({"foo": 1})
    legacy = {"step": 5, "lr": 1e-3, "exp_avg": torch.ones_like(tr.exp_avg), "exp_avg_sq": torch.ones_like(tr.exp_avg)}
    tr.load_optimizer_state_dict(legacy)                    # round-1 flat layout still loads
    assert tr.step_count == 5 and float(tr.exp_avg.sum()) == tr.exp_avg.numel()
