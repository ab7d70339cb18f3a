import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import numpy as np, torch
from vcmeshconv_b200 import _lib
from vcmeshconv_b200.connectivity import DeviceTables
from test_gpu_fused import random_table, reference, trunc_tf32
TF32 = _lib.PREC_TF32
B, p_in, p_out, n_max, M, I, O = 16, 40, 23, 7, 17, 32, 64
tab, ids = random_table(p_in, p_out, n_max, seed=1)
t = DeviceTables(tab, p_in); L = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(7)
x = torch.randn(B, p_in, I, device="cuda", generator=g)
A0 = torch.randn(p_out, n_max, M, device="cuda", generator=g)
W = torch.randn(M, O * I, device="cuda", generator=g) / np.sqrt(M * I)
ptr = lambda a: a.data_ptr() if a is not None else None
def run(A, Wm):
    out = torch.full((B, p_out, O), float("nan"), device="cuda"); z = torch.full((B, p_out, M * I), float("nan"), device="cuda")
    ws = torch.empty(max(L.vcm_conv_fwd_fused_workspace(t.handle, B, M, I, O, TF32), 1), device="cuda")
    _lib.check(L.vcm_conv_fwd_fused(t.handle, ptr(x), ptr(A), ptr(Wm), None, 0, None, None, ptr(out), ptr(z), ptr(ws), B, M, I, O, 0, 1.0, 0.0, TF32, None))
    torch.cuda.synchronize()
    z64 = reference(x, A, ids, Wm, M, I, O, p_in)
    want = torch.einsum("bpmi,moi->bpo", z64, Wm.view(M, O, I).double())
    return out, z, z64, want
out, z, z64, want = run(A0, W)
ez = (z.double() - z64.reshape(B, p_out, M * I)).abs()
print("z max err", ez.max().item(), "nan", torch.isnan(z).sum().item())
print("z err by m:", ez.view(B, p_out, M, I).amax((0, 1, 3)).cpu().numpy().round(4))
print("z err by i:", ez.view(B, p_out, M, I).amax((0, 1, 2)).cpu().numpy().round(4))
print("out err", (out.double() - want).abs().max().item(), "max", want.abs().max().item())
for m in range(M):
    A = torch.zeros_like(A0); A[:, :, m] = A0[:, :, m]
    out, z, z64, want = run(A, W)
    print("only m=%d: out err %.4f (max %.3f)" % (m, (out.double() - want).abs().max().item(), want.abs().max().item()))
for i in range(0, I, 4):
    Wm = torch.zeros_like(W).view(M, O, I); Wm[:, :, i:i + 4] = W.view(M, O, I)[:, :, i:i + 4]; Wm = Wm.reshape(M, O * I).contiguous()
    out, z, z64, want = run(A0, Wm)
    print("only i=%d..%d: out err %.4f (max %.3f)" % (i, i + 3, (out.double() - want).abs().max().item(), want.abs().max().item()))
