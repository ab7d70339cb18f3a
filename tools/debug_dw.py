"""Bring-up diagnostics for the MN-major dW tensor-core GEMM (not a test)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcmeshconv_b200 import _lib
L = _lib.lib()
fn = L.vcm_debug_dw_stage
fn.restype = C.c_int
fn.argtypes = [C.c_void_p] * 5 + [C.c_int64] + [C.c_int] * 10 + [C.c_void_p]

R, K, O = 64, 128, 64
torch.manual_seed(0)
z = torch.randn(R, K, device="cuda"); ds = torch.randn(R, O, device="cuda")
tr = lambda t: (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
want = (tr(z)[:32].double().t() @ tr(ds)[:32].double()).float()        # k-block 0: rows 0..31 -> [K, O]
for (a_mn, b_mn, lbo, sbo, js, swz32, lt) in [
        (1, 1, 4096, 512, 1024, 1, 1), (1, 1, 512, 4096, 1024, 1, 1), (1, 1, 4096, 1024, 1024, 1, 1),
        (1, 1, 4096, 512, 1024, 0, 1), (1, 1, 4096, 512, 1024, 1, 2), (1, 1, 4096, 256, 1024, 1, 1),
        (1, 1, 1024, 512, 1024, 1, 1)]:
    da = torch.zeros(128 * 32, device="cuda"); db = torch.zeros(256 * 32, device="cuda"); dd = torch.zeros(128 * 64, device="cuda")
    rc = fn(ds.data_ptr(), z.data_ptr(), da.data_ptr(), db.data_ptr(), dd.data_ptr(), R, K, O, a_mn, b_mn, lbo, sbo, js, 0, swz32, lt, None)
    torch.cuda.synchronize()
    D = dd.view(128, 64)
    err = (D - want).abs().max().item()
    print("cfg mn=%d%d lbo=%d sbo=%d jstride=%d tma32=%d ltype=%d rc=%d |D| max %.3f  err vs want %.5f  nonzero %d" %
          (a_mn, b_mn, lbo, sbo, js, swz32, lt, rc, D.abs().max().item(), err, int((D != 0).sum())))
    if swz32 and lt == 1 and lbo == 4096 and sbo == 512:
        A = da.view(4, 32, 4, 8).cpu()          # [blk][r][32B chunk][8 floats]
        zz = z[:32].cpu()
        for r in range(8):
            src = zz[r, 0:8]
            pos = [c for c in range(4) if torch.equal(A[0, r, c], src)]
            print("  row %d: global chunk 0 landed in smem chunk %s" % (r, pos))
