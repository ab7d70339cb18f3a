#!/usr/bin/env python
"""Benchmark of the vc-conv hot path (BASELINE.json metric and configs).

    python bench.py [--config c1|c2|c3|c4|c5] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

  c2 (default)  BASELINE configs[1]: the 6+6-layer autoencoder of graphVAE_train.py:151,172 on a synthetic
                6890-vertex SMPL-topology template (convex-hull sphere: 6890 v / 13776 f), batch 16 per GPU,
                residual_rate 0.9, per-point bias, M = 17: training meshes/s (fwd + bwd + all-reduce + Adam)
  c3            the same stack on a 5023-vertex (FLAME-sized) template, batch 64 per GPU
  c4            the same stack on a 150 000-vertex template, batch 8 per GPU
  c1            configs[0]: ONE LASMConvssw layer, 6890 vertices, 3 -> 32 channels, batch 16, forward + backward
  c5            one layer on the tetrahedral, non-manifold template with irregular degree (N_max = 32), 64 -> 64

One JSON line on stdout (rank 0). `value` = meshes/s with the batch already resident in HBM (CUDA-graph replay of
the step); `e2e` = the same step through the public API with the batch copied from pinned host memory and the result
read back every step. `roofline` = the dominant kernel of the step against the measured HBM bandwidth, from CUDA
events around every C-ABI call of an eager, one-stream pass whose launches were queued behind a device-side delay (so
an event pair brackets the kernels alone, not the host's launch gaps). `--impl reference` times the UNMODIFIED
reference modules (oracle/_ref/graphVAESSW.py, shipped there by oracle/Makefile; else the pinned CPU port) on the
host cores; the same thing, on a bounded sample, is the `cpu_baseline` of the default run.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_LEVELS = 7
RESIDUAL = 0.9
M = 17
METRIC_AE = "training meshes/sec (fwd+bwd) D-FAUST AE"
METRIC_LAYER = "vc-conv layer meshes/sec (fwd+bwd)"

CONFIGS = {
    "c1": dict(kind="layer", template="sphere", vertices=6890, batch=16, cin=3, cout=32, residual=0.0,
               name="C1: single graphVAESSW vc-conv layer fwd+bwd, 6890-vertex SMPL-topology template, 3->32 ch"),
    "c2": dict(kind="ae", template="sphere", vertices=6890, batch=16, name="C2: D-FAUST-shaped AE"),
    "c3": dict(kind="ae", template="sphere", vertices=5023, batch=64, name="C3: CoMA FLAME-sized AE"),
    "c4": dict(kind="ae", template="sphere", vertices=150000, batch=8, name="C4: high-resolution AE"),
    "c5": dict(kind="layer", template="tets_capped", vertices=20000, batch=16, cin=64, cout=64, residual=0.0,
               name="C5: single vc-conv layer fwd+bwd, tetrahedral non-manifold template, irregular degree "
                    "(N_max 32), 64->64 ch"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default=os.environ.get("VCM_PRECISION", "auto"))
    ap.add_argument("--batch", type=int, default=0, help="override the config's per-GPU batch")
    ap.add_argument("--vertices", type=int, default=0, help="override the config's template size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-other-modes", "--no-fp32-mode", dest="no_other_modes", action="store_true",
                    help="skip the secondary timings of the fp32-grade arithmetic modes")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying CUDA graphs")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--cpu-batch", type=int, default=0, help="batch of the CPU sample (default: the config's, 2 for c4)")
    ap.add_argument("--breakdown-file", default="", help="write the full per-kernel, per-shape timing table here")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["batch"] = args.batch
    if args.vertices:
        cfg["vertices"] = args.vertices
        cfg["name"] += " (template size overridden)"
    args.cfg = cfg
    return args


def workload(cfg):
    if cfg["kind"] == "ae":
        return "%s, %d-vertex sphere template, 6+6 layers [3,32,64,128,256,512,64], M=17, residual 0.9, batch %d per GPU" % (
            cfg["name"], cfg["vertices"], cfg["batch"])
    return "%s, %d vertices, M=17, per-point bias, residual %.1f, batch %d" % (cfg["name"], cfg["vertices"],
                                                                                cfg["residual"], cfg["batch"])


def config_keys(cfg, gbatch, world, precision, launch, trainable_params, point_nums, l2, tables, arm):
    """The `config` object: both arms print exactly the same keys (the driver compares them)."""
    return {"workload": workload(cfg), "global_batch": gbatch, "vertices": cfg["vertices"], "parallelism": "dp%d" % world,
            "gemm_precision": precision, "launch": launch, "trainable_params": trainable_params,
            "point_nums": point_nums, "l2": l2, "tables": tables, "arm": arm}


# ---------------------------------------------------------------------------------
# synthetic data of the reference's shape (graphVAE_train.py:96-123: vertices in metres, unit face normals)
# ---------------------------------------------------------------------------------
def synthetic_batch(pts, faces, batch, seed):
    import numpy as np
    from vcmeshconv_b200 import synthetic
    rng = np.random.default_rng(seed)
    pc = (pts[None] * 0.9 + rng.standard_normal((batch,) + pts.shape).astype(np.float32) * 0.02).astype(np.float32)
    nor = synthetic.face_normals(pc, faces).astype(np.float32)
    return pc, nor


def template(cfg):
    from vcmeshconv_b200 import synthetic
    return getattr(synthetic, cfg["template"])(cfg["vertices"], seed=0)


def levels(cfg):
    from vcmeshconv_b200 import synthetic
    return synthetic.alternating_levels(N_LEVELS) if cfg["kind"] == "ae" else [(1, 1, 1)]


def build_tables(n_vert_or_cfg):
    """Product arm: connectivity tables from the repo's own sampler (libvcmesh_sampling.so)."""
    from vcmeshconv_b200 import graph_sampling
    cfg = n_vert_or_cfg if isinstance(n_vert_or_cfg, dict) else dict(CONFIGS["c2"], vertices=n_vert_or_cfg)
    pts, faces = template(cfg)
    tabs, nums = graph_sampling.build_tables(faces, len(pts), levels(cfg))
    return pts, faces, tabs, nums


def build_tables_reference(cfg):
    """Reference arm: tables from the reference's own GraphSampling (oracle/_ref/graphsampling_ref, compiled by
    oracle/Makefile from the reference's sources); none of this repo's libraries are involved. Without the binary
    (a tree where /root/reference was never present) a pure-numpy breadth-first restatement is not available, so the
    tables then come from the repo's sampler, and the line says so."""
    from oracle import graphsampling_ref
    from vcmeshconv_b200 import synthetic
    pts, faces = template(cfg)
    if graphsampling_ref.available():
        with tempfile.TemporaryDirectory() as d:
            obj = os.path.join(d, "template.obj")
            synthetic.write_obj(obj, pts, faces)
            tabs = graphsampling_ref.run(obj, levels(cfg))
        return pts, faces, tabs, "reference GraphSampling (oracle/_ref/graphsampling_ref)"
    from vcmeshconv_b200 import graph_sampling
    tabs, _ = graph_sampling.build_tables(faces, len(pts), levels(cfg))
    return pts, faces, tabs, "vcmeshconv_b200.graph_sampling (oracle/_ref/graphsampling_ref absent)"


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
# CPU arm: the reference's own modules (or the pinned port) on the host cores
# ---------------------------------------------------------------------------------
class CpuArm:
    """One step of the config's workload with the reference's arithmetic on `device`: the UNMODIFIED reference
    modules out of oracle/_ref/graphVAESSW.py when that build-time copy exists (`kind` "reference"), else the
    line-cited port oracle/vcconv_oracle.py (`kind` "port", pinned to the reference by tests/golden)."""

    def __init__(self, cfg, tabs, pts, faces, batch, device="cpu", force_port=False):
        import numpy as np
        import torch
        from oracle import ref_autoencoder as ref, vcconv_oracle as orc
        self.torch, self.cfg, self.batch, self.device = torch, cfg, batch, device
        torch.set_num_threads(os.cpu_count() or 1)
        self.kind = "reference" if ref.available() and not force_port and device == "cpu" else "port"
        if cfg["kind"] == "ae":
            enc = [tabs["pool%d" % l] for l in range(1, 7)]
            dec = [tabs["unpool%d" % l] for l in range(6, 0, -1)]
            pc, nor = synthetic_batch(pts, faces, batch, 0)
            self.pc, self.nor = torch.from_numpy(pc).to(device), torch.from_numpy(nor).to(device)
            if self.kind == "reference":
                self.ae = ref.ReferenceAutoEncoder(enc, dec, len(pts), tabs["pool0"], faces, residual_rate=RESIDUAL, seed=0)
            else:
                self.ae = orc.AutoEncoder(enc, dec, len(pts), tabs["pool0"], faces, residual_rate=RESIDUAL,
                                          seed=0).requires_grad_(True).to(device)
                self.opt = torch.optim.Adam(self.ae.parameters(), lr=1e-4, betas=(0.9, 0.999))   # graphVAE_train.py:260
        else:
            table = tabs["pool0"]
            g = torch.Generator().manual_seed(0)
            self.x = (torch.randn(batch, len(pts), cfg["cin"], generator=g) * 0.3).to(device)
            self.gy = torch.randn(batch, table.shape[0], cfg["cout"], generator=g).to(device)
            if self.kind == "reference":
                self.layer = ref.ReferenceLayer(table, cfg["cin"], cfg["cout"], M, len(pts), cfg["residual"], seed=0)
            else:
                self.layer = orc.Layer(table, cfg["cin"], cfg["cout"], len(pts), cfg["residual"])
                self.layer.counts, self.layer.ids = self.layer.counts.to(device), self.layer.ids.to(device)
                self.params = {k: v.to(device).requires_grad_(True) for k, v in orc.init_layer_params(
                    table, cfg["cin"], cfg["cout"], M, len(pts), True, cfg["residual"], g).items()}
                self.coeffs = orc.init_coeffs(table, M, g).to(device).requires_grad_(True)

    def step(self):
        torch = self.torch
        if self.cfg["kind"] == "ae":
            eps = torch.randn(self.batch, *self.ae.latent_shape, device=self.device)
            if self.kind == "reference":
                return self.ae.step(self.pc, self.nor, eps)
            loss = self.ae.losses(self.pc, self.nor, eps)[0]
            self.opt.zero_grad()
            loss.backward()
            self.opt.step()
            return loss
        if self.kind == "reference":
            return self.layer.step(self.x, self.gy)
        for p in list(self.params.values()) + [self.coeffs]:
            p.grad = None
        x = self.x.detach().requires_grad_(True)
        y = self.layer(x, self.coeffs, self.params)
        y.backward(self.gy)
        return y

    def run(self, steps, warmup):
        sync = self.torch.cuda.synchronize if self.device != "cpu" else (lambda: None)
        times = []
        for it in range(warmup + steps):
            sync()
            t0 = time.perf_counter()
            self.step()
            sync()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        mean = sum(times) / len(times)
        return self.batch / mean, mean * 1e3, self.torch.get_num_threads()

    def n_params(self):
        if self.cfg["kind"] == "ae":
            ps = self.ae.opt.param_groups[0]["params"] if self.kind == "reference" else self.ae.parameters()
        elif self.kind == "reference":
            ps = list(self.layer.layer.parameters()) + [self.layer.coeffs]
        else:
            ps = list(self.params.values()) + [self.coeffs]
        return int(sum(p.numel() for p in ps))

    def what(self):
        if self.kind == "reference":
            return "the unmodified reference modules (oracle/_ref/graphVAESSW.py: LASMConvssw / MCEnc / MCDec / MCLoss" \
                   + (" under the trainer glue of graphVAE_train.py:199-225 + torch.optim.Adam)" if self.cfg["kind"] == "ae" else ")")
        return "oracle/vcconv_oracle.py, the line-cited port of graphVAESSW.py:97-163 (same ATen op sequence)"


def cpu_sample_batch(args):
    if args.cpu_batch:
        return args.cpu_batch
    return 2 if args.cfg["vertices"] >= 100000 else args.cfg["batch"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.cfg
    pts, faces, tabs, tables_from = build_tables_reference(cfg)
    batch = cpu_sample_batch(args)
    arm = CpuArm(cfg, tabs, pts, faces, batch)
    value, ms, cores = arm.run(args.steps, args.warmup)
    what = "fwd+bwd+Adam" if cfg["kind"] == "ae" else "fwd+bwd"
    sample = "%d warm-up + %d timed full steps (%s) of batch %d%s on the host cores" % (
        args.warmup, args.steps, what, batch, "" if batch == cfg["batch"] else " (of %d)" % cfg["batch"])
    world = max(int(os.environ.get("WORLD_SIZE", args.gpus)), 1)
    point_nums = [len(pts)] + ([int(tabs["pool%d" % l].shape[0]) for l in range(N_LEVELS)] if cfg["kind"] == "ae"
                               else [int(tabs["pool0"].shape[0])])
    print(json.dumps({
        "impl": "reference", "metric": METRIC_AE if cfg["kind"] == "ae" else METRIC_LAYER, "value": value,
        "unit": "meshes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_keys(cfg, cfg["batch"] * world, world, "fp32 (ATen on the host cores)",
                              "eager PyTorch on the CPU", arm.n_params(), point_nums, "n/a (host memory)", tables_from,
                              arm.what()),
        "cpu_baseline": {"value": value, "unit": "meshes/s", "cores": cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": "meshes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------
# our arm: runners (autoencoder training step / single layer forward + backward)
# ---------------------------------------------------------------------------------
class AERunner:
    def __init__(self, args, dev, rank, world):
        import torch
        from vcmeshconv_b200 import graphVAE_train as T
        self.torch, self.args, self.dev = torch, args, dev
        cfg = args.cfg
        self.pts, self.faces, self.tabs, self.nums = build_tables(cfg)
        enc = [self.tabs["pool%d" % l] for l in range(1, 7)]
        dec = [self.tabs["unpool%d" % l] for l in range(6, 0, -1)]
        param = T.make_param(enc, dec, self.tabs["pool0"], len(self.pts), RESIDUAL)
        torch.manual_seed(0)
        self.model = T.Net_autoenc(param, self.faces).to(dev)
        self.trainer = T.DataParallelTrainer(self.model, lr=1e-4)
        self.n_params = sum(p.numel() for p in self.model.trainable_parameters())
        # a few distinct synthetic batches, rotated; the step's own working set (GBs of intermediates) is far larger
        # than the 126 MB L2, so nothing stays cached from one step to the next
        self.n_rot = 4
        host = [synthetic_batch(self.pts, self.faces, cfg["batch"], 1000 * rank + i) for i in range(self.n_rot)]
        self.pin = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(n).pin_memory()) for p, n in host]
        self.resident = [(p.to(dev), n.to(dev)) for p, n in self.pin]
        self.h2d = self.resident[0][0].numel() * 4 + self.resident[0][1].numel() * 4
        self.d2h = 4
        self.loss_host = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
        self.loss_ev = [None, None]
        self.loss_seen = []
        self.extra = {"trainable_params": self.n_params, "point_nums": self.nums,
                      "l2": "inputs rotate over %d batches; each step streams GBs of intermediates (>> 126 MB L2), "
                            "nothing survives in L2 between steps" % self.n_rot}

    def capture(self):
        self.trainer._graph = None
        self.trainer.capture(self.resident[0][0], self.resident[0][1], warmup=2)

    def launches_per_step(self):
        return self.trainer.graph_launches if self.trainer._graph is not None else None

    def step(self, i):
        pc, nor = self.resident[i % self.n_rot]
        self.trainer.step(pc, nor)

    # End to end through the trainer's host-fed API: every step copies ITS batch from pinned host memory (staged on a
    # copy stream while the previous step runs) and its loss goes back to pinned host memory; the host reads each
    # loss value one step later, so it never runs more than two steps ahead of the device.
    def e2e_begin(self):
        self.trainer.prefetch(*self.pin[0])

    def step_e2e(self, i):
        torch = self.torch
        loss, _ = self.trainer.step_prefetched()                       # consumes the batch staged for step i
        hp, hn = self.pin[(i + 1) % self.n_rot]
        self.trainer.prefetch(hp, hn)                                  # pinned host -> device for step i + 1
        s2 = i & 1
        if self.loss_ev[s2] is not None:
            self.loss_ev[s2].synchronize()
            self.loss_seen.append(float(self.loss_host[s2][0]))        # the host consumes the loss of step i - 2
        self.loss_host[s2].copy_(loss.reshape(1), non_blocking=True)   # device -> host read of the step's result
        self.loss_ev[s2] = torch.cuda.Event()
        self.loss_ev[s2].record()

    def eager_one_stream(self):
        """Context for the per-kernel pass: no graph, branch streams off (every call's kernels alone on one stream)."""
        tr = self.trainer
        branches = ("dw_stream", "res_stream", "prep_stream", "early_stream", "da_stream")
        saved = {k: getattr(tr, k) for k in branches}
        g = tr._graph

        class Ctx:
            def __enter__(s):
                tr._graph = None
                for k in branches:
                    setattr(tr, k, None)

            def __exit__(s, *exc):
                for k, st in saved.items():
                    setattr(tr, k, st)
                tr._graph = g
        return Ctx()


class LayerRunner:
    """BASELINE configs[0] / [4]: ONE LASMConvssw layer (reference module API) with its MCVcoeffs-style coefficient
    parameter, forward + backward (gradients of x, the coefficients, the bases and the bias)."""

    def __init__(self, args, dev, rank, world):
        import torch
        from vcmeshconv_b200 import graphVAESSW as vae
        self.torch, self.args, self.dev = torch, args, dev
        cfg = args.cfg
        self.pts, self.faces, self.tabs, self.nums = build_tables(cfg)
        table = self.tabs["pool0"]
        torch.manual_seed(0)
        self.layer = vae.LASMConvssw(cfg["cin"], cfg["cout"], M, len(self.pts), table, True, cfg["residual"]).to(dev)
        n = self.layer.max_neighbor_num
        self.A = torch.nn.Parameter((torch.randn(table.shape[0], n, M) / (self.layer.avg_neighbor_num * M)).to(dev))
        self.params = [self.A] + [p for p in self.layer.parameters() if p.requires_grad]
        self.n_rot = 4
        B, P, Po = cfg["batch"], len(self.pts), table.shape[0]
        g = torch.Generator().manual_seed(1 + rank)
        self.pin = [((torch.randn(B, P, cfg["cin"], generator=g) * 0.3).pin_memory(),
                     torch.randn(B, Po, cfg["cout"], generator=g).pin_memory()) for _ in range(self.n_rot)]
        self.resident = [(x.to(dev), gy.to(dev)) for x, gy in self.pin]
        self.sx, self.sgy = self.resident[0][0].clone().requires_grad_(True), self.resident[0][1].clone()
        self.graph, self.outs = None, None
        self.h2d = 4 * (self.sx.numel() + self.sgy.numel())
        self.d2h = 4 * (B * Po * cfg["cout"] + self.sx.numel())
        self.host_y = torch.empty(B, Po, cfg["cout"]).pin_memory()
        self.host_dx = torch.empty(B, P, cfg["cin"]).pin_memory()
        self.n_launch = None
        self.extra = {"trainable_params": sum(p.numel() for p in self.params), "n_max": int(n),
                      "nnz": int(table[:, 0].sum()), "point_nums": [P, Po],
                      "l2": "inputs rotate over %d batches; one step streams %.0f MB" % (
                          self.n_rot, 4e-6 * B * Po * M * cfg["cin"] * 5)}

    def _fb(self):
        torch = self.torch
        y = self.layer(self.sx, self.A)
        grads = torch.autograd.grad(y, [self.sx] + self.params, self.sgy)
        return y, grads

    def capture(self):
        torch = self.torch
        from vcmeshconv_b200 import _lib
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._fb()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        l0 = _lib.lib().vcm_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outs = self._fb()
        self.n_launch = int(_lib.lib().vcm_launch_count() - l0)

    def launches_per_step(self):
        return self.n_launch if self.graph is not None else None

    def step(self, i):
        x, gy = self.resident[i % self.n_rot]
        with self.torch.no_grad():
            self.sx.copy_(x, non_blocking=True)
            self.sgy.copy_(gy, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.outs = self._fb()

    def e2e_begin(self):
        pass

    def step_e2e(self, i):
        """Host buffers in, host buffers out: x and the upstream gradient come from pinned host memory, y and dx go
        back to pinned host memory, all inside the timed region."""
        x, gy = self.pin[i % self.n_rot]
        with self.torch.no_grad():
            self.sx.copy_(x, non_blocking=True)
            self.sgy.copy_(gy, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.outs = self._fb()
        self.host_y.copy_(self.outs[0], non_blocking=True)
        self.host_dx.copy_(self.outs[1][0], non_blocking=True)

    def eager_one_stream(self):
        me, g = self, self.graph

        class Ctx:
            def __enter__(s):
                me.graph = None

            def __exit__(s, *exc):
                me.graph = g
        return Ctx()


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
    cfg = args.cfg
    # default = the north_star's design: tcgen05 TF32 basis GEMMs (fp32 storage and accumulation, stated tolerance,
    # tests/test_gpu_fullsize.py); the fp32-grade modes (3-pass TF32 split on the same kernels; CUDA-core fp32) are
    # timed as well and reported under "modes" / "e2e.fp32_grade" so all arithmetic modes are on the same line.
    precision = args.precision
    if precision == "auto":
        precision = os.environ.get("VCM_DEFAULT_PRECISION", "tf32")
    F_.set_precision(precision)
    runner = (AERunner if cfg["kind"] == "ae" else LayerRunner)(args, dev, rank, world)
    lib = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.vcm_launch_count()
        t0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        n_launch = lib.vcm_launch_count() - l0
        if runner.launches_per_step() is not None:
            n_launch = runner.launches_per_step() * steps          # replayed, not re-issued from the host
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, n_launch, t0, t1

    warm = max(args.warmup, 3)
    if not args.no_graph:
        runner.capture()
    for i in range(warm):
        runner.step(i)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, t0, t1 = timed(runner.step, args.steps)
    clocks = sampler.stop(t0, t1) if sampler else None
    runner.e2e_begin()
    for i in range(2):
        runner.step_e2e(i)
    ms_e2e, _, _, _ = timed(runner.step_e2e, args.steps)

    gbatch = cfg["batch"] * world
    value = gbatch * args.steps / (ms / 1e3)
    e2e = gbatch * args.steps / (ms_e2e / 1e3)

    # The other arithmetic modes on the same workload, each re-captured as its own CUDA graph and timed the same
    # way (device-resident and end to end).
    modes = {precision: {"value": value, "ms_per_step": ms / args.steps, "e2e": e2e}}
    if not args.no_other_modes:
        for mode in ("tf32", "tf32x3", "fp32"):
            if mode in modes or (mode == "fp32" and cfg["vertices"] > 20000):
                continue
            F_.set_precision(mode)
            if not args.no_graph:
                runner.capture()
            for i in range(3):
                runner.step(i)
            m1, _, _, _ = timed(runner.step, args.steps)
            runner.e2e_begin()
            for i in range(2):
                runner.step_e2e(i)
            m2, _, _, _ = timed(runner.step_e2e, args.steps)
            modes[mode] = {"value": gbatch * args.steps / (m1 / 1e3), "ms_per_step": m1 / args.steps,
                           "e2e": gbatch * args.steps / (m2 / 1e3)}
        F_.set_precision(precision)
        if not args.no_graph:
            runner.capture()
    notes = {"tf32": "tcgen05 kind::tf32, single pass; stated tolerance (C2 step vs fp64 oracle: out 1e-3, gradients 5e-2)",
             "tf32x3": "tcgen05, 3-pass TF32 hi/lo split: fp32-grade (C2 step vs fp64 oracle <= 5e-5)",
             "fp32": "CUDA-core FMA everywhere, the reference's own arithmetic (C2 step vs fp64 oracle <= 1e-5)"}
    for k in modes:
        modes[k]["arithmetic"] = notes[k]

    roof, breakdown = None, None
    if not args.no_profile:
        # every rank runs the pass (the step contains the gradient all-reduce); rank 0 reports
        with runner.eager_one_stream():
            roof, breakdown = roofline(runner, min(args.steps, 5), args.breakdown_file if rank == 0 else "")
    cpu, aten = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_sample_batch(args)
        arm = CpuArm(cfg, runner.tabs, runner.pts, runner.faces, cb)
        v, cms, cores = arm.run(args.cpu_steps, 1)
        cpu = {"value": v, "unit": "meshes/s", "cores": cores, "kind": arm.kind, "ms_per_step": cms,
               "sample": "1 warm-up + %d timed full steps (batch %d%s) of %s on the host cores" % (
                   args.cpu_steps, cb, "" if cb == cfg["batch"] else " of %d" % cfg["batch"], arm.what())}
        del arm
        try:
            # the reference's ATen op sequence (eager autograd, torch.optim.Adam) on this GPU: its stock GPU path, a
            # second, stronger reported baseline
            arm = CpuArm(cfg, runner.tabs, runner.pts, runner.faces, cb, device="cuda")
            v, ams, _ = arm.run(10, 3)
            aten = {"value": v, "unit": "meshes/s", "ms_per_step": ams, "kind": "port",
                    "what": "oracle/vcconv_oracle.py (the reference's ATen op sequence, fp32, eager autograd) on this "
                            "GPU: 3 warm-up + 10 timed steps, batch %d" % cb}
        except Exception as e:                                   # a reported baseline must not fail the bench
            aten = {"unavailable": repr(e)[:200]}
    if rank == 0:
        grade = modes.get("tf32x3") or modes.get("fp32")
        print(json.dumps({
            "metric": METRIC_AE if cfg["kind"] == "ae" else METRIC_LAYER, "value": value, "unit": "meshes/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if precision == "fp32" else ("tf32 multiply, f32 storage and accumulate" if precision == "tf32"
                                                       else "f32-grade (3-pass tf32 split), f32 storage and accumulate"),
            "data": "synthetic",
            "config": config_keys(cfg, gbatch, world, precision,
                                  "eager" if args.no_graph else "cuda-graph replay of the whole step",
                                  runner.extra["trainable_params"], runner.extra["point_nums"], runner.extra["l2"],
                                  "vcmeshconv_b200.graph_sampling (bit-identical to the reference sampler: "
                                  "tests/test_graph_sampling.py)", "vcmeshconv_b200 (libvcmesh.so through the module API)"),
            "e2e": {"value": e2e, "unit": "meshes/s", "h2d_bytes_per_step": runner.h2d, "d2h_bytes_per_step": runner.d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "fp32_grade": None if grade is None or precision != "tf32" else
                    {"value": grade["e2e"], "unit": "meshes/s", "gemm_precision": "tf32x3" if "tf32x3" in modes else "fp32"}},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "modes": modes, "fp32_mode": modes.get("tf32x3"), "aten_gpu_baseline": aten,
            "kernel_breakdown": breakdown}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def roofline(runner, steps, dump=""):
    """Per-entry-point CUDA-event timing of `steps` further steps, launched eagerly on ONE stream behind a device-side
    delay: the host enqueues the whole step while the stream is still busy with the delay kernel, so every event pair
    (recorded on the launching stream right before and after a C-ABI call) brackets that call's kernels alone -
    no host launch gap is inside any pair. `achieved` = algorithmic bytes (SURVEY 8d formulas, counted per call in
    functional.py) of the dominant kernel = the (entry point, kernel variant) pair with the largest summed device
    time, over ALL its launches, / that summed time."""
    import torch
    from vcmeshconv_b200 import _lib
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, which = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    delay = int(os.environ.get("VCM_BENCH_DELAY_CYCLES", "40000000"))      # ~20 ms: longer than the host needs per step
    runner.step(0)                                                # eager warm-up of this mode (allocator, streams)
    torch.cuda.synchronize()
    with _lib.Profiler() as prof:
        for i in range(steps):
            torch.cuda._sleep(delay)
            runner.step(i)
            torch.cuda.synchronize()
    # what an event pair around nothing measures behind the same delay (the two record commands): subtracted per call
    torch.cuda._sleep(delay)
    pairs = []
    for _ in range(200):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); b.record()
        pairs.append((a, b))
    torch.cuda.synchronize()
    empty = sorted(a.elapsed_time(b) for a, b in pairs)
    pair_ms = empty[len(empty) // 2]
    summ = prof.summary(pair_ms)
    total = sum(v["ms"] for v in summ.values())
    if dump:
        with open(dump, "w") as f:
            f.write("# sum of the calls' device times: %.3f ms/step (one stream, no overlap, no launch gaps)\n" % (total / steps))
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
    # (entry point, variant) groups: functional.py tags the K5 calls "mma" = tensor-pipe kernel / "fp32" = CUDA-core
    # kernel (include/vcmesh.h vcm_coeff_grad_variant); other entry points have one kernel family
    # The three basis-GEMM entry points (K2 forward, K3 dz, K4 dW) launch ONE kernel, tc_gemm_kernel<BN, MODE, ...>
    # (csrc/gemm_tc.cu), in its three operand modes: they form one group, so that the dominant kernel does not flip
    # between three near-equal thirds of the same kernel from run to run.
    family = {"vcm_basis_fwd": "basis_gemm", "vcm_basis_bwd_dz": "basis_gemm", "vcm_basis_bwd_dz_stacked": "basis_gemm",
              "vcm_basis_bwd_dw": "basis_gemm"}
    label = {"basis_gemm": "tc_gemm_kernel (tcgen05 basis GEMMs: vcm_basis_fwd + vcm_basis_bwd_dz + vcm_basis_bwd_dw)"}
    groups = {}
    for name, v in summ.items():
        for tag, tv in v["by_tag"].items():
            var = tag.split()[-1] if tag.split() and tag.split()[-1] in ("mma", "fp32", "tma") else ""
            fam = family.get(name, name)
            g = groups.setdefault((fam, var), {"ms": 0.0, "bytes": 0, "calls": 0, "shapes": {}})
            g["ms"] += tv["ms"]; g["bytes"] += tv["bytes"]; g["calls"] += tv["calls"]
            key = tag if fam == name else "%s %s" % (name.replace("vcm_basis_", ""), tag)
            sh = g["shapes"].setdefault(key, {"ms": 0.0, "bytes": 0, "calls": 0, "flops": 0})
            for kk in ("ms", "bytes", "calls", "flops"):
                sh[kk] += tv[kk]
    (top, var), g = max(groups.items(), key=lambda kv: kv[1]["ms"])
    achieved = g["bytes"] / g["ms"] / 1e6                      # GB/s
    shapes = [{"shape": t, "share_of_step": tv["ms"] / total, "avg_launch_ms": tv["ms"] / tv["calls"],
               "algorithmic_bytes_per_launch": tv["bytes"] / tv["calls"], "achieved": tv["bytes"] / tv["ms"] / 1e6,
               "frac": tv["bytes"] / tv["ms"] / 1e6 / peak}
              for t, tv in sorted(g["shapes"].items(), key=lambda kv: -kv[1]["ms"])]
    # DRAM traffic: measured / algorithmic bytes of the committed `ncu --set full` capture of this kernel on one of
    # its shapes (profiles/traffic_r2.json), applied to this run's average algorithmic bytes per launch
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r2.json")))["kernels"].get((top + " " + var).strip())
        if tr:
            traffic = tr["ratio"] * g["bytes"] / g["calls"]
            traffic_src = "%s: %.1f MB measured vs %.1f MB algorithmic on %s" % (
                tr["source"], tr["dram_bytes"] / 1e6, tr["algorithmic_bytes"] / 1e6, tr["layer"])
    except Exception:
        pass
    ranking = [{"kernel": (n + " " + v).strip(), "ms_per_step": gg["ms"] / steps, "share_of_step": gg["ms"] / total,
                "achieved": gg["bytes"] / gg["ms"] / 1e6 if gg["ms"] else None,
                "frac": gg["bytes"] / gg["ms"] / 1e6 / peak if gg["ms"] else None}
               for (n, v), gg in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])[:8]]
    roof = {"bound": "hbm", "kernel": label.get(top, top) + (" (%s kernel)" % var if var else ""), "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)",
            "launches": g["calls"], "avg_launch_ms": g["ms"] / g["calls"],
            "algorithmic_bytes_per_launch": g["bytes"] / g["calls"], "share_of_step": g["ms"] / total,
            "shapes": shapes[:4], "ranking": ranking,
            "whole_step": {"algorithmic_bytes": sum(v["bytes"] for v in summ.values()) / steps,
                           "sum_of_kernel_ms": total / steps, "calls_per_step": sum(v["calls"] for v in summ.values()) / steps,
                           "event_pair_overhead_us_subtracted_per_call": 1e3 * pair_ms},
            "how": "CUDA events on the launching stream around every C-ABI call of %d eager one-stream steps run right "
                   "after the timed region, each step queued behind a ~20 ms device-side delay so that no host launch "
                   "gap falls inside an event pair (the median time of an empty event pair is subtracted per call); all launches "
                   "of the kernel (%d layer shapes) summed; the dominant "
                   "kernel is the (entry point, kernel variant) with the largest summed device time" % (steps, len(shapes))}
    return roof, breakdown


def main():
    args = parse()
    # Only the JSON line may reach stdout: libraries that print from C (NCCL's version banner, ...) write to file
    # descriptor 1, so for the duration of the run descriptor 1 is pointed at stderr and the line goes to the saved one.
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    global print
    out = os.fdopen(real, "w")

    def print(*a, **k):                                          # noqa: A001 (the two arms print exactly one line)
        k.setdefault("file", out)
        import builtins
        builtins.print(*a, **k)
        out.flush()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
