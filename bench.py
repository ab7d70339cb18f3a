#!/usr/bin/env python
"""Benchmark of the vc-conv hot path: training meshes/sec of the D-FAUST-shaped
mesh autoencoder (BASELINE.json metric), fwd + bwd + gradient all-reduce + Adam.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload at every N (weak scaling): BASELINE config C2 — the 6+6-layer
autoencoder of graphVAE_train.py:151,172 on a synthetic 6890-vertex
SMPL-topology template (convex-hull sphere: 6890 v / 13776 f), per-GPU batch
16, residual_rate 0.9, per-point bias, M = 17, fp32 data, synthetic meshes.

One JSON line on stdout (rank 0). `value` = meshes/s with the batch already
resident in HBM; `e2e` = the same step through the public trainer API with the
batch copied from pinned host memory and the loss read back every step.
`--impl reference` times the CPU restatement of the reference (oracle/) on the
host cores; it is also the `cpu_baseline` of the default run (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_VERT = 6890
BATCH = 16
N_LEVELS = 7
RESIDUAL = 0.9
METRIC = "training meshes/sec (fwd+bwd) D-FAUST AE"


def workload(args):
    tag = {(6890, 16): "C2: D-FAUST-shaped AE", (150000, 8): "C4-sized AE (150k vertices)",
           (5023, 64): "C3-sized AE (FLAME-like)"}.get((args.vertices, args.batch), "AE")
    return "%s, %d-vertex sphere template, 6+6 layers [3,32,64,128,256,512,64], M=17, residual 0.9, batch %d per GPU" % (
        tag, args.vertices, args.batch)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("VCM_PRECISION", "auto"))
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--vertices", type=int, default=N_VERT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-fp32-mode", action="store_true", help="skip the secondary timing of the exact-fp32 mode")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying CUDA graphs")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--breakdown-file", default="", help="write the full per-kernel, per-shape timing table here")
    return ap.parse_args()


# ---------------------------------------------------------------------------------
# synthetic data of the reference's shape (graphVAE_train.py:96-123: vertices in
# metres, unit face normals)
# ---------------------------------------------------------------------------------
def synthetic_batch(pts, faces, batch, seed):
    import numpy as np
    from vcmeshconv_b200 import synthetic
    rng = np.random.default_rng(seed)
    pc = (pts[None] * 0.9 + rng.standard_normal((batch,) + pts.shape).astype(np.float32) * 0.02).astype(np.float32)
    nor = synthetic.face_normals(pc, faces).astype(np.float32)
    return pc, nor


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [c.strip() for c in line.split(",")]
            if len(f) < 7 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1]); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference, on the host cores
# ---------------------------------------------------------------------------------
def cpu_reference_run(args, steps, warmup, tabs, pts, faces, device="cpu"):
    """`device="cuda"` runs the same restatement (the reference's ATen op sequence: index_select, einsum, sum, elu,
    autograd) on the GPU: the reference's own stock GPU path, as a second, stronger reported baseline."""
    import torch
    from oracle import vcconv_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    enc = [tabs["pool%d" % l] for l in range(1, 7)]
    dec = [tabs["unpool%d" % l] for l in range(6, 0, -1)]
    ae = orc.AutoEncoder(enc, dec, len(pts), tabs["pool0"], faces, residual_rate=RESIDUAL, seed=0).requires_grad_(True)
    ae = ae.to(device)
    opt = torch.optim.Adam(ae.parameters(), lr=1e-4, betas=(0.9, 0.999))    # graphVAE_train.py:260
    pc, nor = synthetic_batch(pts, faces, args.batch, 0)
    pc, nor = torch.from_numpy(pc).to(device), torch.from_numpy(nor).to(device)
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)
    times = []
    for it in range(warmup + steps):
        sync()
        t0 = time.perf_counter()
        eps = torch.randn(args.batch, *ae.latent_shape, device=device)
        loss = ae.losses(pc, nor, eps)[0]
        opt.zero_grad()
        loss.backward()
        opt.step()
        sync()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return args.batch / mean, mean * 1e3, torch.get_num_threads()


def build_tables(n_vert):
    from vcmeshconv_b200 import graph_sampling, synthetic
    pts, faces = synthetic.sphere(n_vert, seed=0)
    tabs, nums = graph_sampling.build_tables(faces, n_vert, synthetic.alternating_levels(N_LEVELS))
    return pts, faces, tabs, nums


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pts, faces, tabs, _ = build_tables(args.vertices)
    value, ms, cores = cpu_reference_run(args, args.steps, args.warmup, tabs, pts, faces)
    sample = "%d warm-up + %d timed full steps (fwd+bwd+Adam) of batch %d on the host cores" % (args.warmup, args.steps, args.batch)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "meshes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args), "global_batch": args.batch, "vertices": args.vertices,
                   "note": "CPU restatement of the reference (oracle/vcconv_oracle.py, same ATen op sequence as "
                           "graphVAESSW.py:97-163); the Python reference itself cannot travel to the GPU box"},
        "cpu_baseline": {"value": value, "unit": "meshes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "meshes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from vcmeshconv_b200 import _lib, functional as F_
    from vcmeshconv_b200 import graphVAE_train as T

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # NCCL's version banner must not land in front of the JSON line
    rank, local, world = T.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local)
    # default = the north_star's design: tcgen05 TF32 basis GEMMs (fp32 storage and accumulation, stated
    # tolerance 3e-3); the fp32-grade mode (3-pass TF32 split on the same kernels, <= 1e-5 vs the reference) is
    # timed as well and reported under "fp32_mode" so both arithmetic modes are on the same line.
    precision = args.precision
    if precision == "auto":
        precision = os.environ.get("VCM_DEFAULT_PRECISION", "tf32")
    F_.set_precision(precision)

    pts, faces, tabs, nums = build_tables(args.vertices)
    enc = [tabs["pool%d" % l] for l in range(1, 7)]
    dec = [tabs["unpool%d" % l] for l in range(6, 0, -1)]
    param = T.make_param(enc, dec, tabs["pool0"], len(pts), RESIDUAL)
    torch.manual_seed(0)
    model = T.Net_autoenc(param, faces).to(dev)
    trainer = T.DataParallelTrainer(model, lr=1e-4)
    n_params = sum(p.numel() for p in model.trainable_parameters())

    # a few distinct synthetic batches, rotated; the step's own working set (GBs of intermediates)
    # is far larger than the 126 MB L2, so nothing stays cached from one step to the next
    n_rot = 4
    host = [synthetic_batch(pts, faces, args.batch, 1000 * rank + i) for i in range(n_rot)]
    pin = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(n).pin_memory()) for p, n in host]
    resident = [(p.to(dev), n.to(dev)) for p, n in pin]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.lib().vcm_launch_count()
        t0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        n_launch = _lib.lib().vcm_launch_count() - l0
        if trainer._graph is not None:
            n_launch = trainer.graph_launches * steps              # replayed, not re-issued from the host
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, n_launch, t0, t1

    def step_resident(i):
        pc, nor = resident[i % n_rot]
        trainer.step(pc, nor)

    # End to end through the trainer's host-fed API: every step copies ITS batch from pinned host memory (staged on a
    # copy stream while the previous step runs) and its loss goes back to pinned host memory; the host reads each
    # loss value one step later, so it never runs more than two steps ahead of the device.
    loss_host = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
    loss_ev = [None, None]
    loss_seen = []
    h2d = resident[0][0].numel() * 4 + resident[0][1].numel() * 4

    def step_e2e(i):
        loss, _ = trainer.step_prefetched()                        # consumes the batch staged for step i
        hp, hn = pin[(i + 1) % n_rot]
        trainer.prefetch(hp, hn)                                   # pinned host -> device for step i + 1
        s2 = i & 1
        if loss_ev[s2] is not None:
            loss_ev[s2].synchronize()
            loss_seen.append(float(loss_host[s2][0]))              # the host consumes the loss of step i - 2
        loss_host[s2].copy_(loss.reshape(1), non_blocking=True)    # device -> host read of the step's result
        loss_ev[s2] = torch.cuda.Event()
        loss_ev[s2].record()

    if not args.no_graph:
        trainer.capture(resident[0][0], resident[0][1], warmup=2)
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, t0, t1 = timed(step_resident, args.steps)
    clocks = sampler.stop(t0, t1) if sampler else None
    trainer.prefetch(*pin[0])
    for i in range(2):
        step_e2e(i)
    ms_e2e, _, _, _ = timed(step_e2e, args.steps)

    gbatch = args.batch * world
    value = gbatch * args.steps / (ms / 1e3)
    e2e = gbatch * args.steps / (ms_e2e / 1e3)

    # fp32-grade arithmetic (<= 1e-5 vs the reference) on the same tensor-core kernels: 3-pass TF32 hi/lo split.
    # Re-captured as its own CUDA graph and timed the same way.
    fp32_mode = None
    if precision == "tf32" and not args.no_fp32_mode:
        F_.set_precision("tf32x3")
        g, trainer._graph = trainer._graph, None
        if not args.no_graph:
            trainer.capture(resident[0][0], resident[0][1], warmup=2)
        for i in range(3):
            step_resident(i)
        ms32, _, _, _ = timed(step_resident, args.steps)
        fp32_mode = {"value": gbatch * args.steps / (ms32 / 1e3), "unit": "meshes/s", "ms_per_step": ms32 / args.steps,
                     "steps": args.steps, "gemm_precision": "tf32x3",
                     "gemm": "tcgen05, 3-pass TF32 hi/lo split: fp32-grade (<= 1e-5 vs reference, tests/test_gpu_gemm_tc.py)"}
        trainer._graph = g
        F_.set_precision(precision)

    roof, breakdown = None, None
    if not args.no_profile:
        # per-kernel events need eager launches: same kernels, same stream, launched one by one. Every rank
        # runs the pass (the step contains the gradient all-reduce); rank 0 reports.
        # one stream, so that every call's event pair brackets that call's kernels alone
        g, trainer._graph = trainer._graph, None
        branches = ("dw_stream", "res_stream", "prep_stream", "early_stream", "da_stream")
        saved = {k: getattr(trainer, k) for k in branches}
        for k in branches:
            setattr(trainer, k, None)
        roof, breakdown = roofline(trainer, resident, n_rot, min(args.steps, 5),
                                   args.breakdown_file if rank == 0 else "")
        for k, st in saved.items():
            setattr(trainer, k, st)
        trainer._graph = g
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cms, cores = cpu_reference_run(args, args.cpu_steps, 1, tabs, pts, faces)
        cpu = {"value": v, "unit": "meshes/s", "cores": cores, "kind": "port", "ms_per_step": cms,
               "sample": "1 warm-up + %d timed full steps (fwd+bwd+Adam, batch %d) of oracle/vcconv_oracle.py on "
                         "the host cores" % (args.cpu_steps, args.batch)}
    aten = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, ams, _ = cpu_reference_run(args, 10, 3, tabs, pts, faces, device="cuda")
            aten = {"value": v, "unit": "meshes/s", "ms_per_step": ams, "kind": "port",
                    "what": "oracle/vcconv_oracle.py (the reference's ATen op sequence, fp32, eager autograd, "
                            "torch.optim.Adam) on this GPU: 3 warm-up + 10 timed steps, batch %d" % args.batch}
        except Exception as e:                                   # a reported baseline must not fail the bench
            aten = {"unavailable": repr(e)[:200]}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "meshes/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32" if precision == "fp32" else "tf32 multiply, f32 storage and accumulate", "data": "synthetic",
            "config": {"workload": workload(args), "global_batch": gbatch, "vertices": args.vertices,
                       "parallelism": "dp%d" % world, "gemm_precision": precision,
                       "launch": "eager" if args.no_graph else "cuda-graph replay of the whole step", "trainable_params": n_params,
                       "point_nums": nums, "l2": "inputs rotate over %d batches; each step streams GBs of "
                       "intermediates (>> 126 MB L2), nothing survives in L2 between steps" % n_rot},
            "e2e": {"value": e2e, "unit": "meshes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "fp32_mode": fp32_mode, "aten_gpu_baseline": aten,
            "kernel_breakdown": breakdown}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def roofline(trainer, resident, n_rot, steps, dump=""):
    """Per-entry-point CUDA-event timing of `steps` further training steps
    (events on the launching stream). `achieved` = algorithmic bytes (SURVEY
    §8d formulas, counted per call in functional.py) of the dominant kernel
    (entry point x layer shape) over its summed duration."""
    import torch
    from vcmeshconv_b200 import _lib
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, which = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    with _lib.Profiler() as prof:
        for i in range(steps):
            pc, nor = resident[i % n_rot]
            trainer.step(pc, nor)
    summ = prof.summary()
    total = sum(v["ms"] for v in summ.values())
    if dump:
        with open(dump, "w") as f:
            for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
                f.write("%-26s %8.3f ms/step %5.1f%%\n" % (k, v["ms"] / steps, 100 * v["ms"] / total))
                for t, tv in sorted(v["by_tag"].items(), key=lambda kv: -kv[1]["ms"]):
                    f.write("    %-30s x%-3d %8.1f us %8.0f GB/s %7.2f TFLOP/s\n" %
                            (t, tv["calls"] // steps, 1e3 * tv["ms"] / steps, tv["bytes"] / tv["ms"] / 1e6,
                             tv["flops"] / tv["ms"] / 1e9))
    breakdown = {k: {"ms_per_step": v["ms"] / steps, "share": v["ms"] / total, "calls_per_step": v["calls"] / steps,
                     "GBps": v["bytes"] / v["ms"] / 1e6 if v["ms"] else None,
                     "TFLOPs": v["flops"] / v["ms"] / 1e9 if v["ms"] else None,
                     "top": [{"shape": t, "ms_per_step": tv["ms"] / steps,
                              "GBps": tv["bytes"] / tv["ms"] / 1e6 if tv["ms"] else None,
                              "TFLOPs": tv["flops"] / tv["ms"] / 1e9 if tv["ms"] else None}
                             for t, tv in sorted(v["by_tag"].items(), key=lambda kv: -kv[1]["ms"])[:4]]}
                 for k, v in summ.items()}
    # The dominant KERNEL: the entry point with the largest summed time, and inside it the kernel that carries most of
    # that time (functional.py tags the K5 calls "mma" = tensor-pipe kernel / "fp32" = CUDA-core kernel,
    # include/vcmesh.h vcm_coeff_grad_variant), over ALL the layer shapes that kernel was launched on.
    top = max(summ, key=lambda k: summ[k]["ms"])
    groups = {}
    for tag, tv in summ[top]["by_tag"].items():
        var = tag.split()[-1] if tag.split() and tag.split()[-1] in ("mma", "fp32") else ""
        g = groups.setdefault(var, {"ms": 0.0, "bytes": 0, "calls": 0, "shapes": {}})
        g["ms"] += tv["ms"]; g["bytes"] += tv["bytes"]; g["calls"] += tv["calls"]
        g["shapes"][tag] = tv
    var, g = max(groups.items(), key=lambda kv: kv[1]["ms"])
    achieved = g["bytes"] / g["ms"] / 1e6                      # GB/s
    shapes = [{"shape": t, "share_of_step": tv["ms"] / total, "avg_launch_ms": tv["ms"] / tv["calls"],
               "algorithmic_bytes_per_launch": tv["bytes"] / tv["calls"], "achieved": tv["bytes"] / tv["ms"] / 1e6,
               "frac": tv["bytes"] / tv["ms"] / 1e6 / peak}
              for t, tv in sorted(g["shapes"].items(), key=lambda kv: -kv[1]["ms"])]
    # DRAM traffic: measured / algorithmic bytes of the committed ncu --set full capture of this kernel on one of
    # these shapes (profiles/traffic_r1.json), applied to this run's average algorithmic bytes per launch
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))
        if tr["kernel"] in (top, top[:-3] if top.endswith("_ex") else top) and any(
                t.startswith(tr["shape"]) for t in g["shapes"]):
            traffic = tr["ratio"] * g["bytes"] / g["calls"]
            traffic_src = "%s: %.1f MB measured vs %.1f MB algorithmic on %s" % (
                tr["source"], tr["dram_bytes"] / 1e6, tr["algorithmic_bytes"] / 1e6, tr["layer"])
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": top + (" (%s kernel)" % var if var else ""), "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)",
            "launches": g["calls"], "avg_launch_ms": g["ms"] / g["calls"],
            "algorithmic_bytes_per_launch": g["bytes"] / g["calls"], "share_of_step": g["ms"] / total,
            "shapes": shapes[:4],
            "entry_point_all_kernels": {"achieved": summ[top]["bytes"] / summ[top]["ms"] / 1e6,
                                        "frac": summ[top]["bytes"] / summ[top]["ms"] / 1e6 / peak,
                                        "launches": summ[top]["calls"], "share_of_step": summ[top]["ms"] / total},
            "how": "CUDA events around every call of this entry point on the launching stream, over %d eager steps "
                   "on one stream right after the timed region; all launches of the kernel (%d layer shapes) summed; "
                   "per-call times include launch gaps the graph replay hides" % (steps, len(shapes))}
    return roof, breakdown


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
