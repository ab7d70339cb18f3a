#!/usr/bin/env python
"""Per-kernel totals of one training step from an ncu launch list
(ncu --metrics gpu__time_duration.sum --csv ... bench.py --no-graph): the last complete step in the file."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
recs = []
for r in rows[start:]:
    try:
        recs.append((r[ki], float(r[vi].replace(",", ""))))
    except (IndexError, ValueError):
        pass
# a step begins with the adam_advance launch... no: the optimizer scalars advance once per step, right before the first
# Adam range is applied, and the step ENDS with the last adam_dev launch after it. Take the last complete step:
# from the kernel after the previous step's last Adam launch to this step's last Adam launch.
adv = [i for i, (k, _) in enumerate(recs) if "adam_advance" in k]
adam = [i for i, (k, _) in enumerate(recs) if "adam_dev" in k or "reduce_adam_bcast" in k]
ends = []
for j, a in enumerate(adv):
    nxt = adv[j + 1] if j + 1 < len(adv) else len(recs)
    last = [i for i in adam if a < i < nxt]
    if last:
        ends.append(last[-1])
step = recs[ends[-2] + 1:ends[-1] + 1] if len(ends) >= 2 else recs[:ends[-1] + 1]


def key(k):
    k = k.replace("void ", "").replace("vcm::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    base = re.match(r"([A-Za-z_0-9:]+)", k).group(1)
    if base.startswith("at::") or base.startswith("memcpy"):
        return "torch glue (elementwise / reduce / fill)"
    if base == "tc_gemm_kernel":
        mode = k.split("<")[1].split(",")[1].strip()
        return "tc_gemm_kernel mode %s (%s)" % (mode, {"0": "K2 fwd", "1": "K3 dz", "2": "K4 dW"}.get(mode, "?"))
    return base


agg, tot = {}, 0.0
for k, v in step:
    a = agg.setdefault(key(k), [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print("kernels in the step: %d, sum of durations %.3f ms (cold-cache, serialised)" % (len(step), tot / 1e6))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-46s x%-3d %8.1f us %5.1f %%" % (k, c, t / 1e3, 100 * t / tot))
