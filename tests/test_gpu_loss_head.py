"""Fused loss head (vcm_reparam_kl_*, vcm_loss_head_*) against the reference formulas
(graphVAE_train.py:199-225, graphVAESSW.py:416-421, :507-542) evaluated in fp64 with plain torch ops, and
against the unfused composition of the same model."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from vcmeshconv_b200 import functional as F_
from vcmeshconv_b200 import graphVAE_train as T
from test_gpu_train import _small_model

pytestmark = pytest.mark.gpu


def _ref_reparam_kl(raw_mu, raw_ls, eps, s_mu, s_ls):
    mu, ls = raw_mu * s_mu, raw_ls * s_ls                      # graphVAESSW.py:286-288
    z = mu + ls.exp() * eps                                    # graphVAE_train.py:199-203
    kl = torch.mean(-0.5 - ls + 0.5 * mu ** 2 + 0.5 * torch.exp(2 * ls))   # :206
    return z, kl


@pytest.mark.parametrize("n_shape", [(4, 13, 8), (16, 51, 64), (8, 700, 64)])
def test_reparam_kl_matches_fp64_formula(n_shape):
    g = torch.Generator(device="cuda").manual_seed(3)
    raw_mu = torch.randn(n_shape, device="cuda", generator=g).requires_grad_(True)
    raw_ls = (torch.randn(n_shape, device="cuda", generator=g) * 3).requires_grad_(True)
    eps = torch.randn(n_shape, device="cuda", generator=g)
    gz = torch.randn(n_shape, device="cuda", generator=g)
    z, kl = F_.reparam_kl(raw_mu, raw_ls, eps, 0.1, 0.01)
    (z * gz).sum().add(kl * 0.37).backward()
    a, b = raw_mu.detach().double().requires_grad_(True), raw_ls.detach().double().requires_grad_(True)
    z64, kl64 = _ref_reparam_kl(a, b, eps.double(), 0.1, 0.01)
    (z64 * gz.double()).sum().add(kl64 * 0.37).backward()
    assert rel_err(z.detach().cpu(), z64.detach().cpu()) < 1e-6
    assert abs(kl.item() - kl64.item()) < 1e-6 * abs(kl64.item())
    assert rel_err(raw_mu.grad.cpu(), a.grad.cpu()) < 1e-5
    assert rel_err(raw_ls.grad.cpu(), b.grad.cpu()) < 1e-5
    # only one of the two outputs used: the other gradient is a NULL pointer for the kernel
    raw_mu.grad = raw_ls.grad = None
    F_.reparam_kl(raw_mu, raw_ls, eps, 0.1, 0.01)[1].backward()
    a.grad = b.grad = None
    _ref_reparam_kl(a, b, eps.double(), 0.1, 0.01)[1].backward()
    assert rel_err(raw_mu.grad.cpu(), a.grad.cpu()) < 1e-5 and rel_err(raw_ls.grad.cpu(), b.grad.cpu()) < 1e-5


def _ref_losses64(model, out_pc, gt, nor):
    """The reference's three geometric losses, op for op, in fp64."""
    faces = model.t_facedata
    dif = out_pc - gt
    tri = dif[:, faces[:, 0]] + dif[:, faces[:, 1]] + dif[:, faces[:, 2]]                      # :213-215
    loss_normal = (tri / 3.0 * nor).sum(2).pow(2).mean()                                      # :217
    l1 = torch.abs(gt - out_pc).mean()                                                        # graphVAESSW.py:419
    ids = model.net_loss.initial_neighbor_id_lstlst
    cnt = model.net_loss.initial_neighbor_num_lst.double()
    B, P, _ = gt.shape

    def lap(pc):                                                                              # graphVAESSW.py:515-534
        pad = torch.cat((pc, torch.zeros(B, 1, 3, dtype=pc.dtype, device=pc.device)), 1)
        out = pad[:, ids[:, 0]] * cnt.view(1, P, 1)
        for n in range(1, ids.shape[1]):
            out = out - pad[:, ids[:, n]]
        return out
    lap_l2 = (lap(gt) - lap(out_pc)).pow(2).sum(2).mean()                                      # :537-539
    return l1, lap_l2, loss_normal


def test_loss_head_matches_reference_formulas_fp64():
    model, pc, nor = _small_model(0)
    g = torch.Generator(device="cuda").manual_seed(5)
    out_pc = (pc + torch.randn(pc.shape, device="cuda", generator=g) * 0.05).requires_grad_(True)
    dev = pc.device
    w = (1.0, 100.0, 10.0)
    total, parts = F_.loss_head(out_pc, pc, nor, model.net_loss._lap_coeffs, model.net_loss.tables(dev),
                                model._face_tables_on(dev), *w)
    (total * 1.7).backward()
    o64 = out_pc.detach().double().requires_grad_(True)
    l1, lap, lnor = _ref_losses64(model, o64, pc.double(), nor.double())
    tot64 = l1 * w[0] + lap * w[1] + lnor * w[2]
    (tot64 * 1.7).backward()
    for got, ref in zip(parts.tolist() + [total.item()], [l1.item(), lap.item(), lnor.item(), tot64.item()]):
        assert abs(got - ref) <= 2e-6 * abs(ref), (got, ref)
    assert rel_err(out_pc.grad.cpu(), o64.grad.cpu()) < 1e-5
    # bitwise reproducible (fixed-order partial sums, reverse-CSR gathers)
    g1 = out_pc.grad.clone()
    out_pc.grad = None
    total2, _ = F_.loss_head(out_pc, pc, nor, model.net_loss._lap_coeffs, model.net_loss.tables(dev),
                             model._face_tables_on(dev), *w)
    (total2 * 1.7).backward()
    assert torch.equal(total, total2) and torch.equal(g1, out_pc.grad)


def test_fused_and_unfused_losses_agree_through_the_model():
    model, pc, nor = _small_model(1)
    eps = torch.randn(4, model.nrpt_latent, 8, device="cuda")
    params = model.trainable_parameters()
    outs_f = model.losses(pc, nor, eps, fused=True)
    gf = torch.autograd.grad(outs_f[0], params, allow_unused=True)
    outs_u = model.losses(pc, nor, eps, fused=False)
    gu = torch.autograd.grad(outs_u[0], params, allow_unused=True)
    for a, b in zip(outs_f[:5], outs_u[:5]):
        assert abs(a.item() - b.item()) <= 5e-6 * abs(b.item()), (a.item(), b.item())
    assert torch.equal(outs_f[5], outs_u[5]) or rel_err(outs_f[5].detach().cpu(), outs_u[5].detach().cpu()) < 1e-6
    for p, a, b in zip(params, gf, gu):
        assert (a is None) == (b is None)
        if a is not None:
            assert rel_err(a.cpu(), b.cpu()) < 2e-5, tuple(p.shape)


def test_fused_losses_draw_the_same_eps_stream():
    """Without an explicit eps both paths consume the generator identically (one normal_() of the latent shape)."""
    model, pc, nor = _small_model(2)
    torch.manual_seed(11)
    a = model.losses(pc, nor, fused=True)[0]
    torch.manual_seed(11)
    b = model.losses(pc, nor, fused=False)[0]
    assert abs(a.item() - b.item()) <= 5e-6 * abs(b.item())


def test_loss_head_rejects_bad_shapes_and_cpu():
    model, pc, nor = _small_model(0)
    dev = pc.device
    args = (model.net_loss._lap_coeffs, model.net_loss.tables(dev), model._face_tables_on(dev), 1.0, 100.0, 10.0)
    with pytest.raises(RuntimeError):
        F_.loss_head(pc.cpu(), pc.cpu(), nor.cpu(), *args)
    with pytest.raises(RuntimeError):
        F_.loss_head(pc[:, :-1].contiguous(), pc[:, :-1].contiguous(), nor, *args)
    with pytest.raises(RuntimeError):
        F_.reparam_kl(torch.zeros(3, 4), torch.zeros(3, 4), torch.zeros(3, 4))
