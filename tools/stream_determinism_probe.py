#!/usr/bin/env python
"""Ten captured steps of a C2-like autoencoder under different branch-stream settings: run-to-run reproducibility of
each setting (a race shows up as a difference between two runs of the SAME setting) and the distance between settings
(only the order in which autograd adds the gradients of a shared input may differ)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vcmeshconv_b200 import functional as F_, graphVAE_train as T, synthetic

F_.set_precision(os.environ.get("VCM_PRECISION", "tf32"))


def run(off, steps=10, eager=0):
    model, pts, faces, tabs, nums = T.build_synthetic_autoencoder(3000, seed=0)
    tr = T.DataParallelTrainer(model.cuda(), lr=1e-3)
    for name in off:
        setattr(tr, name, None)
    g = torch.Generator(device="cuda").manual_seed(1)
    pc = torch.from_numpy(pts)[None].repeat(16, 1, 1).cuda() * 0.9 + torch.randn(16, len(pts), 3, device="cuda", generator=g) * 0.02
    nor = torch.from_numpy(synthetic.face_normals(pc.cpu().numpy(), faces)).float().cuda()
    eps = torch.randn(16, tr.model.nrpt_latent, 64, device="cuda", generator=g)
    for _ in range(eager):
        tr.step(pc, nor, eps=eps)
    torch.cuda.synchronize()
    if eager and steps == 0:
        return tr.flat.flat_p.double().cpu().numpy()
    torch.manual_seed(123)
    tr.capture(pc, nor, warmup=1)
    for _ in range(steps):
        tr.step(pc, nor)
    torch.cuda.synchronize()
    return tr.flat.flat_p.double().cpu().numpy()


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


settings = {"none": ("dw_stream", "res_stream", "head_stream", "prep_stream", "early_stream"),
            "dw+res": ("head_stream",), "all": (), "head only": ("dw_stream", "res_stream")}
out = {}
for k, off in settings.items():
    a, b = run(off), run(off)
    out[k] = a
    print("%-10s run-to-run %.3e" % (k, rel(a, b)))
for k in ("dw+res", "all", "head only"):
    pass
for eager, steps in ((3, 0), (3, 10)):
    res = {}
    for k, off in settings.items():
        a, b = run(off, steps, eager), run(off, steps, eager)
        res[k] = a
        print("eager %d + captured %d: %-10s run-to-run %.3e   vs none %.3e" % (eager, steps, k, rel(a, b), rel(a, res["none"])))
for k in ("dw+res", "all", "head only"):
    print("%-10s vs none   %.3e" % (k, rel(out[k], out["none"])))
print("all vs dw+res        %.3e" % rel(out["all"], out["dw+res"]))
