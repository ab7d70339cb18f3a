#!/usr/bin/env python
"""Probe: does running producer -> consumer pairs of the layer in BATCH SLICES keep the big intermediate (z forward,
dz backward) in the 126 MB L2?  K1 -> K2 (z) and K3 -> K5 (dz, one slice-sized workspace reused by every slice) on a
C2 layer shape, whole batch vs 2 / 4 / 8 slices, timed with CUDA events behind a device-side delay.

    python tools/slice_probe.py --layer dec3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vcmeshconv_b200 import _lib, connectivity
from vcmeshconv_b200 import graph_sampling as gs
from vcmeshconv_b200 import synthetic
from kernel_bench import LAYERS

ap = argparse.ArgumentParser()
ap.add_argument("--layer", default="dec3")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
pts, faces = synthetic.sphere(6890, seed=0)
tabs, nums = gs.build_tables(faces, 6890, synthetic.alternating_levels(7))
tname, lvl, I, O = LAYERS[a.layer]
table = tabs[tname]
mx = int(table[:, 1::2].max())
p_in = mx if mx in nums else mx + 1
t = connectivity.DeviceTables(table, p_in)
B, M, P, N = 16, 17, t.p_out, t.n_max
L, prec = _lib.lib(), _lib.PRECISIONS["tf32"]
st = _lib.stream_ptr()
x = torch.randn(B, p_in, I, device="cuda")
A = torch.randn(P, N, M, device="cuda")
z = torch.empty(B, P, M * I, device="cuda")
W = torch.randn(M, O * I, device="cuda")
Wt = torch.empty(M * I * O, device="cuda")
_lib.check(L.vcm_basis_stack(W.data_ptr(), Wt.data_ptr(), M, I, O, st))
ds = torch.randn(B, P, O, device="cuda")
out = torch.empty(B, P, O, device="cuda")
dG = torch.empty(B, P, N, I, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    tot = 0.0
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(1000000)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return 1e3 * tot / a.iters


def fwd(ns):
    bs = B // ns
    ws = torch.empty(max(L.vcm_basis_fwd_workspace(bs * P, M, I, O, prec), 1), device="cuda")

    def f():
        for s in range(ns):
            o = s * bs
            _lib.check(L.vcm_gather_contract_fwd_ex(t.handle, x.data_ptr() + 4 * o * p_in * I, A.data_ptr(),
                                                    z.data_ptr() + 4 * o * P * M * I, bs, I, M, prec, st))
            _lib.check(L.vcm_basis_fwd(z.data_ptr() + 4 * o * P * M * I, W.data_ptr(), None, 0, None, None,
                                       out.data_ptr() + 4 * o * P * O, ws.data_ptr(), bs, P, M, I, O, 1, 1.0, 0.0, prec, st))
    return f


def bwd(ns):
    bs = B // ns
    dzs = torch.empty(bs, P, M * I, device="cuda")
    dA = torch.empty(ns, P, N, M, device="cuda")
    ws = torch.empty(max(L.vcm_coeff_grad_workspace_ex(t.handle, bs, I, M, prec), 1), device="cuda")

    def f():
        for s in range(ns):
            o = s * bs
            _lib.check(L.vcm_basis_bwd_dz_stacked(ds.data_ptr() + 4 * o * P * O, Wt.data_ptr(), dzs.data_ptr(), bs * P, M, I,
                                                  O, prec, st))
            _lib.check(L.vcm_coeff_grad_ex(t.handle, dzs.data_ptr(), x.data_ptr() + 4 * o * p_in * I, A.data_ptr(),
                                           dA[s].data_ptr(), dG.data_ptr() + 4 * o * P * N * I, ws.data_ptr(), bs, I, M,
                                           prec, st))
    return f


Z = 4e-6 * B * P * M * I
print("%s: P %d->%d N %d I %d O %d, Z = %.0f MB" % (a.layer, p_in, P, N, I, O, Z))
for ns in (1, 2, 4, 8):
    f, g = fwd(ns), bwd(ns)
    for _ in range(2):
        f(); g()
    torch.cuda.synchronize()
    print("  %d slice(s) of %.0f MB: K1 -> K2 %.1f us   K3 -> K5 %.1f us" % (ns, Z / ns, timed(f), timed(g)))
