"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED
reference (imported from /root/reference/GraphAutoEncoder) on seeded inputs.

Run in the build container only (the reference does not travel to the GPU box):

    python oracle/gen_golden.py

The vectors pin oracle/vcconv_oracle.py (tests/test_oracle_golden.py) and are
the fixed points the CUDA path is compared with on the GPU
(tests/test_gpu_parity.py). For every case the reference is run

  * in fp64 (torch.set_default_dtype(float64) + .double(): type promotion
    absorbs the reference's hard-coded .float() masks) -> gold outputs and
    autograd gradients the 1e-5 relative tolerance is measured against;
  * in fp32 through LASMConvssw.forward and .forward2 -> recorded so that the
    reference's own fp32 spread between its two formulations is on file.
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/GraphAutoEncoder")

import graphVAESSW as ref  # noqa: E402  (the reference itself)

from oracle import graphsampling_ref  # noqa: E402
from vcmeshconv_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def tables_for(pts, faces, levels, green=()):
    with tempfile.TemporaryDirectory() as d:
        obj = os.path.join(d, "m.obj")
        synthetic.write_obj(obj, pts, faces, green)
        return graphsampling_ref.run(obj, levels)


def randomise(module, gen, scale_bias=0.1):
    """Reference initialisers leave bias at 0; give every parameter a non-trivial
    value so each gradient path is exercised."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=gen) * scale_bias)


def layer_case(name, table, p_in, cin, cout, m, batch, residual, perpoint_bias, final, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    layer = quiet(ref.LASMConvssw, cin, cout, m, p_in, table, perpoint_bias, residual)
    randomise(layer, gen)
    p_out, nmax = layer.out_point_num, layer.max_neighbor_num
    coeffs = torch.randn(p_out, nmax, m, generator=gen) / (layer.avg_neighbor_num * m) * 4.0
    x = torch.randn(batch, p_in, cin, generator=gen) * 0.3
    gy = torch.randn(batch, p_out, cout, generator=gen)

    rec = {"table": table.astype(np.int32), "in_point_num": np.int64(p_in), "in_channel": np.int64(cin),
           "out_channel": np.int64(cout), "weight_num": np.int64(m), "residual_rate": np.float64(residual),
           "is_final_layer": np.bool_(final), "x": x.numpy(), "coeffs": coeffs.numpy(), "gy": gy.numpy()}
    for k, v in layer.state_dict().items():
        rec["param." + k] = v.numpy().copy()

    # fp32 through both reference formulations
    with torch.no_grad():
        rec["y32_forward"] = layer(x, coeffs, is_final_layer=final).numpy()
        rec["y32_forward2"] = layer.forward2(x, coeffs, is_final_layer=final).numpy()

    # fp64 gold + autograd gradients
    torch.set_default_dtype(torch.float64)
    try:
        l64 = quiet(ref.LASMConvssw, cin, cout, m, p_in, table, perpoint_bias, residual).double()
        l64.load_state_dict({k: v.double() for k, v in layer.state_dict().items()})
        x64 = x.double().requires_grad_(True)
        a64 = coeffs.double().requires_grad_(True)
        y = l64(x64, a64, is_final_layer=final)
        y2 = l64.forward2(x64, a64, is_final_layer=final)
        rec["y64"] = y.detach().numpy()
        rec["y64_forward2"] = y2.detach().numpy()
        y.backward(gy.double())
        rec["grad.x"] = x64.grad.numpy()
        rec["grad.coeffs"] = a64.grad.numpy()
        for k, p in l64.named_parameters():
            rec["grad." + k] = p.grad.numpy()
    finally:
        torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    d32 = np.abs(rec["y32_forward"] - rec["y32_forward2"]).max()
    print("%-28s P %4d->%4d N %2d I/O %3d/%3d M %2d r %.1f  |fwd-fwd2|_32 %.2e" %
          (name, p_in, p_out, nmax, cin, cout, m, residual, d32))


def stack_case(name, tabs, n_vert, faces, seed):
    """Small 3+3 autoencoder through the reference's MCStructure / MCVcoeffs /
    MCEnc / MCDec / MCLoss, plus the trainer's loss glue
    (graphVAE_train.py:199-225) evaluated with the reference's own modules."""
    m = 5
    enc_names, dec_names = ["pool1", "pool2", "pool3"], ["unpool3", "unpool2", "unpool1"]
    enc_ch, dec_ch = [3, 8, 16, 4], [4, 16, 8, 3]
    resid = 0.9
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    with tempfile.TemporaryDirectory() as d:
        paths = {}
        for k, v in tabs.items():
            paths[k] = os.path.join(d, "_%s.npy" % k)
            np.save(paths[k], v)
        param = types.SimpleNamespace(residual_rate=resid, conv_max=0,
                                      connection_layer_fn_lst_enc=[paths[k] for k in enc_names],
                                      connection_layer_fn_lst_dec=[paths[k] for k in dec_names])
        t0 = tabs["pool0"]
        param.neighbor_id_lstlst = t0[:, 1:].reshape(t0.shape[0], -1, 2)[:, :, 0]   # graphAE_param_iso.py:88-93
        param.neighbor_num_lst = t0[:, 0]

        torch.set_default_dtype(torch.float64)
        try:
            s_enc = quiet(ref.MCStructure, param, n_vert, m, bDec=False)
            vc_enc = quiet(ref.MCVcoeffs, s_enc, m)
            enc = quiet(ref.MCEnc, s_enc, enc_ch, m)
            s_dec = quiet(ref.MCStructure, param, enc.out_nrpts, m, bDec=True)
            vc_dec = quiet(ref.MCVcoeffs, s_dec, m)
            dec = quiet(ref.MCDec, s_dec, dec_ch, m)
            loss_mod = ref.MCLoss(param).double()
            for mod in (enc, dec):
                randomise(mod.double(), gen)
            with torch.no_grad():
                for p in list(vc_enc.parameters()) + list(vc_dec.parameters()):
                    p.mul_(4.0)

            batch = 3
            pts, _ = synthetic.sphere(n_vert, seed=seed)
            x = (torch.from_numpy(pts).double()[None] * 0.3 +
                 torch.randn(batch, n_vert, 3, generator=gen, dtype=torch.float64) * 0.05)
            f = torch.from_numpy(faces).long()
            nor = torch.from_numpy(synthetic.face_normals(x.numpy(), faces))
            eps = torch.randn(batch, enc.out_nrpts, enc_ch[-1], generator=gen, dtype=torch.float64)

            mu, logstd = enc(x, vc_enc)
            std = logstd.exp()
            z = mu + std * eps
            kl = torch.mean(-0.5 - logstd + 0.5 * mu ** 2 + 0.5 * torch.exp(2 * logstd))
            out = dec(z, vc_dec)[:, :, 0:3]
            dif = out - x
            v0 = ref.index_selection_nd(dif, f[:, 0], 1)
            v1 = ref.index_selection_nd(dif, f[:, 1], 1)
            v2 = ref.index_selection_nd(dif, f[:, 2], 1)
            l_nor = ((v1 + v0 + v2) / 3.0 * nor).sum(2).pow(2).mean()
            l_l1 = loss_mod.compute_geometric_loss_l1(x[:, :, 0:3], out[:, :, 0:3])
            l_lap = loss_mod.compute_laplace_loss_l2(x[:, :, 0:3], out[:, :, 0:3])
            loss = l_l1 * 1.0 + l_lap * 100.0 + kl * 1e-5 + l_nor * 10.0
            loss.backward()

            rec = {"n_vert": np.int64(n_vert), "weight_num": np.int64(m), "residual_rate": np.float64(resid),
                   "enc_channels": np.array(enc_ch), "dec_channels": np.array(dec_ch),
                   "enc_names": np.array(enc_names), "dec_names": np.array(dec_names),
                   "faces": faces.astype(np.int32), "x": x.numpy(), "normals": nor.numpy(), "eps": eps.numpy(),
                   "mu": mu.detach().numpy(), "logstd": logstd.detach().numpy(), "out": out.detach().numpy(),
                   "loss": loss.item(), "loss_l1": l_l1.item(), "loss_laplace": l_lap.item(), "loss_kl": kl.item(),
                   "loss_normal": l_nor.item()}
            for k, v in tabs.items():
                rec["table." + k] = v.astype(np.int32)
            for tag, mod in (("enc", enc), ("dec", dec), ("vce", vc_enc), ("vcd", vc_dec)):
                for k, v in mod.state_dict().items():
                    rec["%s.%s" % (tag, k)] = v.numpy().copy()
                for k, p in mod.named_parameters():
                    rec["grad.%s.%s" % (tag, k)] = p.grad.numpy().copy()
        finally:
            torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print("%-28s loss %.6f (l1 %.4f lap %.6f kl %.4f nor %.6f)" %
          (name, rec["loss"], rec["loss_l1"], rec["loss_laplace"], rec["loss_kl"], rec["loss_normal"]))


def checkpoint_case(name, tabs, n_vert, faces, seed):
    """A checkpoint written the way the reference's trainer writes it (graphVAE_train.py:339-347: four module
    state dicts + `optimizer.state_dict()` of the Adam built by getoptim, :252-262) after two fp32 training steps
    of the reference's own modules, plus the inputs of a third step and the parameters the reference holds after
    it. tests/ load the file into DataParallelTrainer, take the third step on the GPU and must land on the same
    parameters (which only happens if moments, step count and learning rate were all taken over)."""
    import itertools
    m = 5
    enc_names, dec_names = ["pool1", "pool2", "pool3"], ["unpool3", "unpool2", "unpool1"]
    enc_ch, dec_ch = [3, 8, 16, 4], [4, 16, 8, 3]
    resid, lr = 0.9, 2e-3
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    with tempfile.TemporaryDirectory() as d:
        paths = {}
        for k, v in tabs.items():
            paths[k] = os.path.join(d, "_%s.npy" % k)
            np.save(paths[k], v)
        param = types.SimpleNamespace(residual_rate=resid, conv_max=0,
                                      connection_layer_fn_lst_enc=[paths[k] for k in enc_names],
                                      connection_layer_fn_lst_dec=[paths[k] for k in dec_names])
        t0 = tabs["pool0"]
        param.neighbor_id_lstlst = t0[:, 1:].reshape(t0.shape[0], -1, 2)[:, :, 0]
        param.neighbor_num_lst = t0[:, 0]
        s_enc = quiet(ref.MCStructure, param, n_vert, m, bDec=False)
        vc_enc = quiet(ref.MCVcoeffs, s_enc, m)
        enc = quiet(ref.MCEnc, s_enc, enc_ch, m)
        s_dec = quiet(ref.MCStructure, param, enc.out_nrpts, m, bDec=True)
        vc_dec = quiet(ref.MCVcoeffs, s_dec, m)
        dec = quiet(ref.MCDec, s_dec, dec_ch, m)
        loss_mod = ref.MCLoss(param)
    for mod in (enc, dec):
        randomise(mod, gen)
    # getoptim, graphVAE_train.py:252-262
    params = [p for p in itertools.chain(dec.parameters(), enc.parameters(), vc_dec.parameters(), vc_enc.parameters())
              if p.requires_grad]
    opt = torch.optim.Adam(params, lr=lr, betas=(0.9, 0.999))
    batch = 4
    pts, _ = synthetic.sphere(n_vert, seed=seed)
    f = torch.from_numpy(faces).long()
    rec = {"n_vert": np.int64(n_vert), "weight_num": np.int64(m), "residual_rate": np.float64(resid), "lr": np.float64(lr),
           "enc_channels": np.array(enc_ch), "dec_channels": np.array(dec_ch),
           "enc_names": np.array(enc_names), "dec_names": np.array(dec_names), "faces": faces.astype(np.int32)}
    for k, v in tabs.items():
        rec["table." + k] = v.astype(np.int32)

    def step(i):
        x = (torch.from_numpy(pts)[None] * 0.3 + torch.randn(batch, n_vert, 3, generator=gen) * 0.05).float()
        nor = torch.from_numpy(synthetic.face_normals(x.numpy(), faces)).float()
        eps = torch.randn(batch, enc.out_nrpts, enc_ch[-1], generator=gen)
        mu, logstd = enc(x, vc_enc)                                                   # graphVAE_train.py:199-225
        z = mu + logstd.exp() * eps
        kl = torch.mean(-0.5 - logstd + 0.5 * mu ** 2 + 0.5 * torch.exp(2 * logstd))
        out = dec(z, vc_dec)[:, :, 0:3]
        dif = out - x
        v0, v1, v2 = (ref.index_selection_nd(dif, f[:, c], 1) for c in range(3))
        l_nor = ((v1 + v0 + v2) / 3.0 * nor).sum(2).pow(2).mean()
        loss = (loss_mod.compute_geometric_loss_l1(x, out) * 1.0 + loss_mod.compute_laplace_loss_l2(x, out) * 100.0 +
                kl * 1e-5 + l_nor * 10.0)
        opt.zero_grad()
        loss.backward()
        opt.step()
        rec["x%d" % i], rec["normals%d" % i], rec["eps%d" % i], rec["loss%d" % i] = x.numpy(), nor.numpy(), eps.numpy(), loss.item()

    step(0)
    step(1)
    torch.save({"encgeo_state_dict": enc.state_dict(), "decgeo_state_dict": dec.state_dict(),
                "mcvcoeffsdec_dict": vc_dec.state_dict(), "mcvcoeffsenc_dict": vc_enc.state_dict(),
                "optimizer_state_dict": opt.state_dict()}, os.path.join(OUT, name + ".weight"))
    step(2)
    for tag, mod in (("enc", enc), ("dec", dec), ("vce", vc_enc), ("vcd", vc_dec)):
        for k, v in mod.state_dict().items():
            rec["after3.%s.%s" % (tag, k)] = v.numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print("%-28s reference-made checkpoint after 2 Adam steps; losses %.5f %.5f %.5f" %
          (name, rec["loss0"], rec["loss1"], rec["loss2"]))


def main():
    if not graphsampling_ref.build():
        raise SystemExit("reference GraphSampling could not be built")
    os.makedirs(OUT, exist_ok=True)
    pts, faces = synthetic.sphere(300, seed=1)
    tabs = tables_for(pts, faces, synthetic.alternating_levels(5))
    n = [300] + [tabs["pool%d" % l].shape[0] for l in range(5)]
    # name, table, P_in, I, O, M, B, residual, per-point bias, final, seed
    layer_case("layer_conv_3to32", tabs["pool0"], n[0], 3, 32, 17, 2, 0.0, True, False, 10)
    layer_case("layer_conv_res_same", tabs["pool2"], n[2], 16, 16, 5, 3, 0.5, True, False, 11)
    layer_case("layer_conv_res_proj", tabs["pool2"], n[2], 8, 24, 17, 2, 0.9, True, False, 12)
    layer_case("layer_pool_res", tabs["pool1"], n[1], 12, 20, 17, 3, 0.9, True, False, 13)
    layer_case("layer_pool_plain", tabs["pool3"], n[3], 32, 64, 17, 2, 0.0, False, False, 14)
    # shapes the tcgen05 GEMMs and the tensor-pipe K5 take (I, O multiples of 32, <= 32 neighbours)
    layer_case("layer_conv_tc_res", tabs["pool2"], n[2], 32, 64, 17, 4, 0.9, True, False, 21)
    layer_case("layer_unpool_res_same", tabs["unpool1"], n[2], 16, 16, 7, 2, 0.9, True, False, 15)
    layer_case("layer_unpool_final", tabs["unpool3"], n[4], 32, 3, 17, 5, 0.9, True, True, 16)
    layer_case("layer_unpool_sharedbias", tabs["unpool1"], n[2], 6, 10, 3, 1, 0.0, False, True, 17)
    # irregular, non-manifold connectivity (config C5 in miniature)
    tp, tf = synthetic.tets(160, seed=2)
    ttabs = tables_for(tp, tf, synthetic.alternating_levels(2))
    layer_case("layer_tets_conv", ttabs["pool0"], 160, 8, 16, 17, 2, 0.0, True, False, 18)
    layer_case("layer_tets_pool_res", ttabs["pool1"], 160, 8, 16, 17, 2, 0.9, True, False, 19)
    stack_case("stack_sphere300", tabs, 300, faces, 20)
    checkpoint_case("ref_checkpoint_sphere300", tabs, 300, faces, 22)


if __name__ == "__main__":
    main()
