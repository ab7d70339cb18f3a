#!/usr/bin/env python
"""Probe (torchrun, 2..8 GPUs): the repo's reduce + Adam + broadcast kernel over NVLink peer memory
(vcm_reduce_adam_bcast, csrc/collective.cu) against torch's two-shot symmetric-memory all-reduce followed by the
replicated fused Adam (vcm_adam_apply) and against NCCL all-reduce + Adam, on a buffer of the C2 (or --n) size:
bit-equality of the resulting parameters and device time per call."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from vcmeshconv_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=13_900_032)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
import torch.distributed._symmetric_memory as sm

gname = dist.group.WORLD.group_name
N = a.n // 64 * 64
L = _lib.lib()


def log(*x):
    if rank == 0:
        print(*x, flush=True)


g = sm.empty(N, dtype=torch.float32, device=dev); hg = sm.rendezvous(g, gname)
p = sm.empty(N, dtype=torch.float32, device=dev); hp = sm.rendezvous(p, gname)
fl = sm.empty(8 * 3 * 128, dtype=torch.int32, device=dev).zero_(); hf = sm.rendezvous(fl, gname)
arr = C.c_void_p * world
peers = lambda h, t: arr(*[int(h.buffer_ptrs[k]) + (t.data_ptr() - int(h.buffer_ptrs[rank])) for k in range(world)])
pg, pp, pf = peers(hg, g), peers(hp, p), peers(hf, fl)
m, v = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
state = torch.tensor([0.0, 1e-4, 0.0, 0.0], device=dev)
torch.manual_seed(1)
p0 = torch.randn(N, device=dev)
torch.manual_seed(100 + rank)
g0 = torch.randn(N, device=dev) * 1e-3
torch.cuda.synchronize(); dist.barrier()
st = lambda: _lib.stream_ptr()


def fused():
    _lib.check(L.vcm_reduce_adam_bcast(pg, pp, pf, m.data_ptr(), v.data_ptr(), 0, N, state.data_ptr(), 0.9, 0.999, 1e-8,
                                       rank, world, 0, st()))


def two_shot():
    torch.ops.symm_mem.two_shot_all_reduce_(g, "sum", gname)
    _lib.check(L.vcm_adam_apply(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), N, state.data_ptr(), 0.9, 0.999,
                                1e-8, 1.0, st()))


def nccl():
    dist.all_reduce(g)
    _lib.check(L.vcm_adam_apply(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), N, state.data_ptr(), 0.9, 0.999,
                                1e-8, 1.0, st()))


res = {}
for name, fn in (("fused", fused), ("two_shot", two_shot), ("nccl", nccl)):
    p.copy_(p0); g.copy_(g0); m.zero_(); v.zero_(); state.copy_(torch.tensor([0.0, 1e-4, 0.0, 0.0]))
    _lib.check(L.vcm_adam_advance(state.data_ptr(), 0.9, 0.999, st()))
    torch.cuda.synchronize(); dist.barrier()
    fn()
    torch.cuda.synchronize(); dist.barrier()
    res[name] = p.clone()
    ref = p.clone(); dist.broadcast(ref, 0)
    same = torch.tensor([float(torch.equal(ref, p))], device=dev); dist.all_reduce(same, op=dist.ReduceOp.MIN)
    log("%-9s identical parameters on all ranks: %s" % (name, bool(same.item())))
log("fused == two_shot + replicated Adam: %s (max abs diff %.3e); == nccl: %s" % (
    torch.equal(res["fused"], res["two_shot"]), (res["fused"] - res["two_shot"]).abs().max().item(),
    torch.equal(res["fused"], res["nccl"])))
for name, fn in (("fused", fused), ("two_shot", two_shot), ("nccl", nccl)):
    for _ in range(3):
        g.copy_(g0); fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    log("%-9s %.1f us per reduce + Adam of %.1f MB on %d GPUs" % (name, 1e3 * t.item(), 4e-6 * N, world))
dist.barrier()
dist.destroy_process_group()
