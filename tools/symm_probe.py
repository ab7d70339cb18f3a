#!/usr/bin/env python
"""Probe (torchrun, N >= 2 GPUs): can the flat gradient buffer be all-reduced by torch's symmetric-memory kernels
(NVLS multimem / two-shot over NVLink peer memory), on a slice, inside a CUDA graph? Prints timings next to NCCL."""
import faulthandler
import os
import sys
import time

import torch
import torch.distributed as dist

faulthandler.dump_traceback_later(90, exit=True)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
import torch.distributed._symmetric_memory as sm

def log(*a):
    if rank == 0:
        print(*a, flush=True)

N = 13_900_032                                  # ~ the C2 flat buffer (55.6 MB), a multiple of 64
gname = dist.group.WORLD.group_name
try:
    g = sm.empty(N, dtype=torch.float32, device=dev)
    hdl = sm.rendezvous(g, gname)
    log("rendezvous ok; multicast ptr:", getattr(hdl, "multicast_ptr", None))
except Exception as e:
    log("symm alloc / rendezvous FAILED:", repr(e)[:300])
    dist.destroy_process_group(); sys.exit(0)

ref = torch.empty(N, device=dev)
def fill():
    torch.manual_seed(rank)
    g.copy_(torch.randn(N, device=dev)); ref.copy_(g)

ops = {}
for name in ("multimem_all_reduce_", "two_shot_all_reduce_"):
    def f(t, name=name):
        return getattr(torch.ops.symm_mem, name)(t, "sum", gname)
    ops[name] = f

split = 7_000_064
for name, f in ops.items():
    try:
        fill()
        dist.all_reduce(ref)
        f(g)
        torch.cuda.synchronize()
        err = (g - ref).abs().max().item() / ref.abs().max().item()
        log("%s full buffer: rel err vs NCCL %.2e" % (name, err))
        fill()
        dist.all_reduce(ref[:split]); dist.all_reduce(ref[split:])
        f(g[:split]); f(g[split:])
        torch.cuda.synchronize()
        err = (g - ref).abs().max().item() / ref.abs().max().item()
        log("%s two slices: rel err %.2e" % (name, err))
        # identical on all ranks?
        chk = g.double().sum().reshape(1).clone(); lst = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        log("%s identical across ranks: %s" % (name, all(torch.equal(lst[0], x) for x in lst)))
        for label, fn in (("eager", lambda: f(g)), ("nccl", lambda: dist.all_reduce(ref))):
            for _ in range(3): fn()
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            log("%s %s: %.1f us per 55.6 MB all-reduce" % (name, label, 1e3 * e0.elapsed_time(e1) / 20))
        # captured, on a side branch next to a compute kernel
        a = torch.randn(4096, 4096, device=dev)
        side = torch.cuda.Stream()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, capture_error_mode="relaxed"):
            cur = torch.cuda.current_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                f(g[:split])
            b = a @ a
            cur.wait_stream(side)
            f(g[split:])
        for _ in range(3): gr.replay()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): gr.replay()
        e1.record(); torch.cuda.synchronize()
        log("%s captured (slice on a branch next to a 4096^3 GEMM + slice): %.1f us per replay" % (name, 1e3 * e0.elapsed_time(e1) / 20))
    except Exception as e:
        log("%s FAILED: %s" % (name, repr(e)[:400]))
dist.barrier()
dist.destroy_process_group()
