"""The data-parallel reduction kernel (csrc/collective.cu, vcm_reduce_adam_bcast) on ONE GPU.

  world = 1   degenerates to the fused Adam of the single-GPU step: bit-identical to vcm_adam_apply on ranges that cut
              through the 16 KB ownership chunks;
  world = 2   two "ranks" emulated on one device: two gradient / parameter / flag buffer sets, the two ranks' kernels
              launched on two streams so that they are co-resident and meet at the flag barriers. Both parameter
              buffers must end up identical and bit-equal to Adam on g0 + g1; each rank's moments are live exactly on
              the chunks it owns (chunk c belongs to rank c % world).

Runs in a subprocess: a protocol bug traps the kernel (bounded barrier waits) and must not take the test session's
CUDA context with it. The real multi-GPU path is covered by tests/multi_gpu_equiv.py (tests/test_gpu_multi.py)."""
import os
import subprocess
import sys
import textwrap

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = textwrap.dedent('''
    import ctypes as C, sys, torch
    sys.path.insert(0, %r)
    from vcmeshconv_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    N = 3 * 4096 * 7 + 1280                      # not a multiple of the 4096-element ownership chunk
    chunk = L.vcm_reduce_adam_chunk()
    g = torch.Generator(device=dev).manual_seed(3)
    p0 = torch.randn(N, device=dev, generator=g)
    grads = [torch.randn(N, device=dev, generator=g) * 1e-3 for _ in range(2)]
    state = torch.tensor([0.0, 1e-4, 0.0, 0.0], device=dev)
    st = lambda s: C.c_void_p(s.cuda_stream)
    cur = torch.cuda.current_stream()
    _lib.check(L.vcm_adam_advance(state.data_ptr(), 0.9, 0.999, st(cur)))
    arr = lambda *t: (C.c_void_p * len(t))(*[x.data_ptr() for x in t])

    def reference(gsum, lo, n):
        p, m, v = p0.clone(), torch.zeros(N, device=dev), torch.zeros(N, device=dev)
        _lib.check(L.vcm_adam_apply(p.data_ptr() + 4 * lo, gsum.data_ptr() + 4 * lo, m.data_ptr() + 4 * lo,
                                    v.data_ptr() + 4 * lo, n, state.data_ptr(), 0.9, 0.999, 1e-8, 1.0, st(cur)))
        return p, m, v

    # world = 1, two ranges that cut through chunks
    p, m, v = p0.clone(), torch.zeros(N, device=dev), torch.zeros(N, device=dev)
    flags = torch.zeros(8 * 3 * 128, dtype=torch.int32, device=dev)
    for lo, n in ((0, 5000), (5000, N - 5000)):
        _lib.check(L.vcm_reduce_adam_bcast(arr(grads[0]), arr(p), arr(flags), m.data_ptr(), v.data_ptr(), lo, n,
                                           state.data_ptr(), 0.9, 0.999, 1e-8, 0, 1, 0, st(cur)))
    torch.cuda.synchronize()
    pr, mr, vr = reference(grads[0], 0, N)
    assert torch.equal(p, pr) and torch.equal(m, mr) and torch.equal(v, vr), "world 1 != vcm_adam_apply"
    assert int(flags.abs().sum()) == 0

    # world = 2 on one device: the two ranks' kernels on two streams
    ps = [p0.clone(), p0.clone()]
    ms = [torch.zeros(N, device=dev) for _ in range(2)]
    vs = [torch.zeros(N, device=dev) for _ in range(2)]
    fl = [torch.zeros(8 * 3 * 128, dtype=torch.int32, device=dev) for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for lo, n, slot in ((0, 8192 + 1024, 0), (8192 + 1024, N - 8192 - 1024, 128)):
        for r in range(2):
            _lib.check(L.vcm_reduce_adam_bcast(arr(*grads), arr(*ps), arr(*fl), ms[r].data_ptr(), vs[r].data_ptr(), lo, n,
                                               state.data_ptr(), 0.9, 0.999, 1e-8, r, 2, slot, st(streams[r])))
    torch.cuda.synchronize()
    pr, mr, vr = reference(grads[0] + grads[1], 0, N)
    assert torch.equal(ps[0], ps[1]), "ranks disagree"
    assert torch.equal(ps[0], pr), "2 ranks != Adam on the summed gradient"
    own = [((torch.arange(N, device=dev) // chunk) %% 2) == r for r in range(2)]
    for r in range(2):
        assert torch.equal(ms[r][own[r]], mr[own[r]]) and torch.equal(vs[r][own[r]], vr[own[r]])
        assert float(ms[r][~own[r]].abs().sum()) == 0.0 and float(vs[r][~own[r]].abs().sum()) == 0.0
        assert int(fl[r].abs().sum()) == 0          # every flag was consumed: the barriers are reusable
    print("COLLECTIVE_OK")
''')


def test_reduce_adam_bcast_single_gpu_emulation():
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and "COLLECTIVE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
