#!/usr/bin/env python
"""Kernel-level view of one eager training step with torch.profiler (which kernels are ours, which are torch glue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vcmeshconv_b200 import functional as F_, graphVAE_train as T

F_.set_precision("tf32")
pts, faces, tabs, nums = bench.build_tables(6890)
enc = [tabs["pool%d" % l] for l in range(1, 7)]; dec = [tabs["unpool%d" % l] for l in range(6, 0, -1)]
torch.manual_seed(0)
model = T.Net_autoenc(T.make_param(enc, dec, tabs["pool0"], 6890, 0.9), faces).cuda()
tr = T.DataParallelTrainer(model)
pc, nor = bench.synthetic_batch(pts, faces, 16, 0)
pc, nor = torch.from_numpy(pc).cuda(), torch.from_numpy(nor).cuda()
for _ in range(3): tr.step(pc, nor)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): tr.step(pc, nor)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages()]
tot = sum(r[2] for r in rows)
print("total device time per step: %.3f ms" % (tot / 3e3))
ours = sum(r[2] for r in rows if "vcm" in r[0]); print("ours %.3f ms, torch %.3f ms" % (ours / 3e3, (tot - ours) / 3e3))
for k, c, t in sorted(rows, key=lambda r: -r[2])[:45]:
    print("%8.1f us x%-4d %s" % (t / 3, c // 3, k[:110]))
