#!/usr/bin/env python
"""Micro-benchmark of the gather-side kernels (K1 / K5 / K6) and the basis GEMMs on one C2 layer shape,
through the C ABI. Used for ncu captures (short, one kernel family per run).

    python tools/kernel_bench.py --layer dec3 --kernel k1 --iters 20
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vcmeshconv_b200 import _lib, connectivity
from vcmeshconv_b200 import graph_sampling as gs
from vcmeshconv_b200 import synthetic

# name: (table, level-in index, I, O)   on the 6890-vertex C2 hierarchy
LAYERS = {"enc1": ("pool2", 2, 32, 64), "enc2": ("pool3", 3, 64, 128), "enc3": ("pool4", 4, 128, 256),
          "enc4": ("pool5", 5, 256, 512), "dec1": ("unpool5", 6, 512, 256), "dec2": ("unpool4", 5, 256, 128),
          "dec3": ("unpool3", 4, 128, 64), "dec4": ("unpool2", 2, 64, 32)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", default="dec3", choices=sorted(LAYERS))
    ap.add_argument("--kernel", default="k1", choices=["k1", "k5", "k6", "fwd", "dz", "dw", "fused", "fused_noz"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--precision", default="tf32")
    a = ap.parse_args()
    pts, faces = synthetic.sphere(6890, seed=0)
    tabs, nums = gs.build_tables(faces, 6890, synthetic.alternating_levels(7))
    tname, lvl, I, O = LAYERS[a.layer]
    table = tabs[tname]
    mx = int(table[:, 1::2].max())                   # the padding sentinel is p_in itself; without padding max id = p_in - 1
    p_in = mx if mx in nums else mx + 1
    t = connectivity.DeviceTables(table, p_in)
    B, M, P, N = a.batch, 17, t.p_out, t.n_max
    L, st = _lib.lib(), _lib.stream_ptr()
    prec = _lib.PRECISIONS[a.precision]
    x = torch.randn(B, p_in, I, device="cuda")
    A = torch.randn(P, N, M, device="cuda")
    z = torch.empty(B, P, M * I, device="cuda")
    dz = torch.randn(B, P, M * I, device="cuda")
    dA, dG, dx = torch.empty_like(A), torch.empty(B, P, N, I, device="cuda"), torch.empty_like(x)
    W = torch.randn(M, O * I, device="cuda")
    ds = torch.randn(B, P, O, device="cuda")
    out = torch.empty(B, P, O, device="cuda")
    dW = torch.empty_like(W)
    ws = torch.empty(max(L.vcm_coeff_grad_workspace_ex(t.handle, B, I, M, prec), L.vcm_basis_bwd_dw_workspace(B * P, M, I, O, prec),
                         L.vcm_basis_bwd_dz_workspace(B * P, M, I, O, prec), L.vcm_basis_fwd_workspace(B * P, M, I, O, prec), 1),
                     device="cuda")
    nnz = t.nnz
    ws_f = torch.empty(max(L.vcm_conv_fwd_fused_workspace(t.handle, B, M, I, O, prec), 1), device="cuda")
    fused = lambda zz: L.vcm_conv_fwd_fused(t.handle, x.data_ptr(), A.data_ptr(), W.data_ptr(), None, 0, None, None,
                                            out.data_ptr(), zz, ws_f.data_ptr(), B, M, I, O, 1, 1.0, 0.0, prec, st)
    X, Z, Ab, Gb, Y, Wb = 4 * x.numel(), 4 * z.numel(), 4 * nnz * M, 4 * B * nnz * I, 4 * out.numel(), 4 * W.numel()
    calls = {
        "k1": (lambda: L.vcm_gather_contract_fwd_ex(t.handle, x.data_ptr(), A.data_ptr(), z.data_ptr(), B, I, M, prec, st),
               X + Ab + 4 * nnz + Z, 2 * B * nnz * M * I),
        "k5": (lambda: L.vcm_coeff_grad_ex(t.handle, dz.data_ptr(), x.data_ptr(), A.data_ptr(), dA.data_ptr(), dG.data_ptr(),
                                           ws.data_ptr(), B, I, M, prec, st), Z + X + 2 * Ab + 4 * nnz + Gb, 4 * B * nnz * M * I),
        "k6": (lambda: L.vcm_scatter_rev(t.handle, dG.data_ptr(), dx.data_ptr(), B, I, 0, st), Gb + 4 * nnz + X, B * nnz * I),
        "fwd": (lambda: L.vcm_basis_fwd(z.data_ptr(), W.data_ptr(), None, 0, None, None, out.data_ptr(), ws.data_ptr(), B, P,
                                        M, I, O, 1, 1.0, 0.0, prec, st), Z + Wb + Y, 2 * B * P * M * I * O),
        "dz": (lambda: L.vcm_basis_bwd_dz(ds.data_ptr(), W.data_ptr(), dz.data_ptr(), ws.data_ptr(), B * P, M, I, O, 0, prec,
                                          st), Z + Wb + Y, 2 * B * P * M * I * O),
        "dw": (lambda: L.vcm_basis_bwd_dw(ds.data_ptr(), z.data_ptr(), dW.data_ptr(), ws.data_ptr(), B * P, M, I, O, prec,
                                          st), Z + Wb + Y, 2 * B * P * M * I * O),
    }
    calls["fused"] = (lambda: fused(z.data_ptr()), X + Ab + 4 * nnz + Z + Wb + Y, 2 * B * nnz * M * I + 2 * B * P * M * I * O)
    calls["fused_noz"] = (lambda: fused(None), X + Ab + 4 * nnz + Wb + Y, 2 * B * nnz * M * I + 2 * B * P * M * I * O)
    fn, nbytes, flops = calls[a.kernel]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        _lib.check(fn())
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(a.iters):
        flush.zero_()                               # evict L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400000)                   # the launch below is queued before the stream reaches e0
        e0.record()
        _lib.check(fn())
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / a.iters
    print("%s %s  P %d->%d N %d nnz %d I %d O %d B %d : %.1f us  %.0f GB/s (algorithmic)  %.2f TFLOP/s  bytes=%d" %
          (a.layer, a.kernel, p_in, P, N, nnz, I, O, B, ms * 1e3, nbytes / ms / 1e6, flops / ms / 1e9, nbytes))


if __name__ == "__main__":
    main()
