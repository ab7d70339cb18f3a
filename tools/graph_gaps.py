#!/usr/bin/env python
"""Idle time between kernels inside the CUDA-graph replay of one training step (torch.profiler / CUPTI timestamps):
span of a replay vs the sum of its kernels' durations, and the gap histogram."""
import os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vcmeshconv_b200 import functional as F_, graphVAE_train as T

F_.set_precision(os.environ.get("VCM_PRECISION", "tf32"))
pts, faces, tabs, nums = bench.build_tables(6890)[:4]
enc = [tabs["pool%d" % l] for l in range(1, 7)]; dec = [tabs["unpool%d" % l] for l in range(6, 0, -1)]
torch.manual_seed(0)
model = T.Net_autoenc(T.make_param(enc, dec, tabs["pool0"], 6890, 0.9), faces).cuda()
tr = T.DataParallelTrainer(model)
pc, nor = bench.synthetic_batch(pts, faces, 16, 0)
pc, nor = torch.from_numpy(pc).cuda(), torch.from_numpy(nor).cuda()
tr.capture(pc, nor)
for _ in range(5): tr.step(pc, nor)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
NREP = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NREP):
        tr.step(pc, nor)
        torch.cuda.synchronize()
fn = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(fn)
ev = [e for e in json.load(open(fn))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
per = len(ev) // NREP
print("kernels per replay:", per)
for r in range(NREP):
    k = ev[r * per:(r + 1) * per]
    span = k[-1]["ts"] + k[-1]["dur"] - k[0]["ts"]
    busy = sum(e["dur"] for e in k)
    gaps = [k[i + 1]["ts"] - (k[i]["ts"] + k[i]["dur"]) for i in range(len(k) - 1)]
    pos = [g for g in gaps if g > 0]
    print("replay %d: span %.1f us, sum of kernel durations %.1f us, idle %.1f us (%d gaps, median %.2f us, max %.1f us), overlapped %d"
          % (r, span, busy, sum(pos), len(pos), sorted(pos)[len(pos) // 2] if pos else 0, max(pos) if pos else 0,
             sum(1 for g in gaps if g <= 0)))
k = ev[(NREP - 1) * per:]
if len(sys.argv) > 1:                      # timeline of the last replay: start (us from the first kernel), duration, stream, name
    with open(sys.argv[1], "w") as f:
        t0 = k[0]["ts"]
        for e in k:
            a = e.get("args", {})
            f.write("%.2f\t%.2f\t%s\t%s\t%s\n" % (e["ts"] - t0, e["dur"], a.get("stream", ""), a.get("grid", ""),
                                                  e["name"].replace("void ", "")[:160]))
    iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in k)
    busy_u, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    for s_, e_ in iv[1:]:
        if s_ > cur_e:
            busy_u += cur_e - cur_s; cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    busy_u += cur_e - cur_s
    print("last replay: union of kernel intervals %.1f us, no kernel running for %.1f us" % (busy_u, iv[-1][1] if False else (max(e for _, e in iv) - iv[0][0] - busy_u)))
small = sum(1 for e in k if e["dur"] < 5); print("kernels < 5 us: %d (sum %.1f us)" % (small, sum(e["dur"] for e in k if e["dur"] < 5)))
agg = {}
for e in k:
    n = e["name"].replace("vcm::<unnamed>::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    n = n.split("(")[0][:110] + " g%s" % (e.get("args", {}).get("grid", ""),)
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += e["dur"]
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print("%8.1f us x%-3d %s" % (t, c, n))

tor = {n: v for n, v in agg.items() if not n.startswith("vcm::")}
print("torch glue: %d kernels, %.1f us; ours: %d kernels, %.1f us" % (sum(v[0] for v in tor.values()), sum(v[1] for v in tor.values()),
      sum(v[0] for n, v in agg.items() if n.startswith("vcm::")), sum(v[1] for n, v in agg.items() if n.startswith("vcm::"))))
for n, (c, t) in sorted(tor.items(), key=lambda kv: -kv[1][1]):
    print("   glue %8.1f us x%-3d %s" % (t, c, n[:150]))
