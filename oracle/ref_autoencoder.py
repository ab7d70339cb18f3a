"""TEST INFRASTRUCTURE — the reference's own modules as the CPU arm of bench.py.

When oracle/_ref/graphVAESSW.py exists (a build-time copy of the UNMODIFIED reference module made by
oracle/Makefile; git-ignored, travels to the GPU box) the CPU baseline runs the reference itself
(`kind: "reference"`): its LASMConvssw / MCStructure / MCVcoeffs / MCEnc / MCDec / MCLoss classes, under the
trainer glue of graphVAE_train.py:199-225 and the Adam of :252-262 restated here (that script cannot be imported:
it executes at import and needs plyfile / tensorboardX). Otherwise callers fall back to the port
(oracle/vcconv_oracle.py, `kind: "port"`), which tests/test_oracle_golden.py pins to the same module's outputs.
Only bench.py's CPU legs use this file; the product never imports it.
"""
import contextlib
import importlib.util
import io
import itertools
import os
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PY = os.path.join(HERE, "_ref", "graphVAESSW.py")
_mod = None


def available():
    return os.path.exists(REF_PY)


def module():
    global _mod
    if _mod is None:
        spec = importlib.util.spec_from_file_location("graphVAESSW_reference", REF_PY)
        m = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(m)
        _mod = m
    return _mod


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):          # the reference prints from its constructors
        return fn(*a, **k)


class ReferenceLayer:
    """One LASMConvssw of the reference with freshly initialised parameters (BASELINE config C1 / C5)."""

    def __init__(self, table, cin, cout, weight_num, p_in, residual_rate, seed=0):
        ref = module()
        torch.manual_seed(seed)
        self.layer = _quiet(ref.LASMConvssw, cin, cout, weight_num, p_in, np.asarray(table), True, residual_rate)
        n = self.layer.max_neighbor_num
        self.coeffs = torch.nn.Parameter(torch.randn(table.shape[0], n, weight_num) /
                                         (self.layer.avg_neighbor_num * weight_num))

    def step(self, x, gy):
        """forward + backward (BASELINE configs[0]: 'single graphVAESSW vc-conv layer fwd+bwd')."""
        for p in list(self.layer.parameters()) + [self.coeffs]:
            p.grad = None
        x = x.detach().requires_grad_(True)
        y = self.layer(x, self.coeffs)
        y.backward(gy)
        return y


class ReferenceAutoEncoder:
    """Net_autoenc of graphVAE_train.py:132-249 out of the reference's own classes (channels and M as hard-coded
    there, :136, :151, :172), with getoptim's Adam (:252-262)."""

    ENC = [3, 32, 64, 128, 256, 512, 64]
    DEC = [64, 512, 256, 128, 64, 32, 3]

    def __init__(self, enc_tables, dec_tables, point_num, table0, faces, residual_rate=0.9, seed=0, lr=1e-4,
                 weight_num=17):
        ref = module()
        torch.manual_seed(seed)
        self._tmp = tempfile.TemporaryDirectory()            # MCStructure np.load()s file names
        paths = {"enc": [], "dec": []}
        for side, tabs in (("enc", enc_tables), ("dec", dec_tables)):
            for i, t in enumerate(tabs):
                fn = os.path.join(self._tmp.name, "_%s%d.npy" % (side, i))
                np.save(fn, np.asarray(t))
                paths[side].append(fn)
        t0 = np.asarray(table0)
        param = types.SimpleNamespace(residual_rate=residual_rate, conv_max=0,
                                      connection_layer_fn_lst_enc=paths["enc"], connection_layer_fn_lst_dec=paths["dec"],
                                      neighbor_id_lstlst=t0[:, 1:].reshape(t0.shape[0], -1, 2)[:, :, 0],
                                      neighbor_num_lst=t0[:, 0])
        self.s_enc = _quiet(ref.MCStructure, param, point_num, weight_num, bDec=False)
        self.vc_enc = _quiet(ref.MCVcoeffs, self.s_enc, weight_num)
        self.enc = _quiet(ref.MCEnc, self.s_enc, self.ENC, weight_num)
        self.s_dec = _quiet(ref.MCStructure, param, self.enc.out_nrpts, weight_num, bDec=True)
        self.vc_dec = _quiet(ref.MCVcoeffs, self.s_dec, weight_num)
        self.dec = _quiet(ref.MCDec, self.s_dec, self.DEC, weight_num)
        self.loss_mod = ref.MCLoss(param)
        self.faces = torch.as_tensor(np.asarray(faces)).long()
        self.latent_shape = (self.enc.out_nrpts, self.ENC[-1])
        params = [p for p in itertools.chain(self.dec.parameters(), self.enc.parameters(), self.vc_dec.parameters(),
                                             self.vc_enc.parameters()) if p.requires_grad]
        self.opt = torch.optim.Adam(params, lr=lr, betas=(0.9, 0.999))
        self._ref = ref

    def step(self, x, nor, eps):
        """graphVAE_train.py:199-225 + :314-316 (weights :184-187)."""
        ref, f = self._ref, self.faces
        mu, logstd = self.enc(x, self.vc_enc)
        z = mu + logstd.exp() * eps
        kl = torch.mean(-0.5 - logstd + 0.5 * mu ** 2 + 0.5 * torch.exp(2 * logstd))
        out = self.dec(z, self.vc_dec)[:, :, 0:3]
        dif = out - x
        v0, v1, v2 = (ref.index_selection_nd(dif, f[:, c], 1) for c in range(3))
        l_nor = ((v1 + v0 + v2) / 3.0 * nor).sum(2).pow(2).mean()
        loss = (self.loss_mod.compute_geometric_loss_l1(x, out) * 1.0 +
                self.loss_mod.compute_laplace_loss_l2(x, out) * 100.0 + kl * 1e-5 + l_nor * 10.0)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss
