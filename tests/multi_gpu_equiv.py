#!/usr/bin/env python
"""Multi-rank equivalence on real GPUs (SURVEY §4 d; reference semantics graphVAE_train.py:249, 277, 315):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_equiv.py [--steps 3] [--vertices 6890] [--batch 16] [--precision tf32]

N ranks train the C2 autoencoder for `steps` steps on the path that ships (flat gradient buffer in symmetric memory,
segment-wise reduction + Adam inside ONE captured step graph per rank), each rank on its own shard with a fixed
eps. Rank 0 then replays the same steps in ONE process: every shard's forward/backward in turn, the shard
gradients SUMMED (the reference sums the per-replica mean losses: `loss.backward(torch.ones(ndev))`), one Adam
update. The two parameter vectors must agree to <= 1e-6 (2 ranks: a + b has one order, so they are bit-equal), and
all ranks must hold identical parameters. Prints one JSON line; non-zero exit on failure. Test infrastructure."""
import argparse
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from vcmeshconv_b200 import functional as F_
from vcmeshconv_b200 import graphVAE_train as T
from vcmeshconv_b200 import synthetic


def shard(pts, faces, batch, rank, step, latent_shape, dev):
    rng = np.random.default_rng(100 * step + rank)
    pc = (pts[None] * 0.9 + rng.standard_normal((batch,) + pts.shape).astype(np.float32) * 0.02).astype(np.float32)
    nor = synthetic.face_normals(pc, faces).astype(np.float32)
    g = torch.Generator().manual_seed(7000 + 100 * step + rank)
    eps = torch.randn((batch,) + latent_shape, generator=g)
    return torch.from_numpy(pc).to(dev), torch.from_numpy(nor).to(dev), eps.to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--vertices", type=int, default=6890)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "multi_gpu_equiv.json"))
    args = ap.parse_args()
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank, local, world = T.init_distributed()
    assert world > 1, "launch with torchrun --nproc-per-node >= 2"
    dev = torch.device("cuda", local)
    F_.set_precision(args.precision)
    solo = dist.new_group([0])                                   # a world of one, for rank 0's single-process replay

    model, pts, faces, tabs, nums = T.build_synthetic_autoencoder(args.vertices)
    model = model.to(dev)
    initial = copy.deepcopy(model.state_dict())
    latent = (model.nrpt_latent, model.net_geoenc.out_nrchs)
    tr = T.DataParallelTrainer(model, lr=1e-4)
    path = "symmetric-memory %s in the step graph" % tr._symm_op if tr._symm_op else "NCCL between two graphs"
    pc, nor, eps = shard(pts, faces, args.batch, rank, 0, latent, dev)
    tr.capture(pc, nor, warmup=0)                                # warm-up restores the state: training starts at step 0
    losses = []
    for s in range(args.steps):
        pc, nor, eps = shard(pts, faces, args.batch, rank, s, latent, dev)
        loss, _ = tr.step(pc, nor, eps)
        losses.append(float(loss))
    torch.cuda.synchronize()
    # all ranks hold the same parameters
    ref = tr.flat.flat_p.clone()
    dist.broadcast(ref, 0)
    spread = torch.tensor([(ref - tr.flat.flat_p).abs().max().item()], device=dev)
    dist.all_reduce(spread, op=dist.ReduceOp.MAX)

    result = None
    if rank == 0:
        model1, _, _, _, _ = T.build_synthetic_autoencoder(args.vertices)
        model1 = model1.to(dev)
        model1.load_state_dict(initial)
        one = T.DataParallelTrainer(model1, lr=1e-4, process_group=solo)
        assert one.world == 1
        assert [tuple(p.shape) for p in one.flat.params] == [tuple(p.shape) for p in tr.flat.params]
        for s in range(args.steps):
            acc = torch.zeros_like(one.flat.flat_g)
            for r in range(world):
                pc, nor, eps = shard(pts, faces, args.batch, r, s, latent, dev)
                one.forward_backward(pc, nor, eps)
                acc += one.flat.flat_g                           # SUM over replicas of per-replica mean-loss gradients
            one.flat.flat_g.copy_(acc)
            one.optimizer_step()
        torch.cuda.synchronize()
        d = (one.flat.flat_p - tr.flat.flat_p).abs()
        result = {"what": "N-rank captured step (%s) vs one process fed all shards with summed gradients" % path,
                  "ranks": world, "steps": args.steps, "vertices": args.vertices, "batch_per_rank": args.batch,
                  "precision": args.precision, "n_params": int(one.flat.total),
                  "max_abs_param_diff": d.max().item(), "n_differing": int((d > 0).sum().item()),
                  "max_abs_param_spread_over_ranks": spread.item(), "losses_rank0": losses,
                  "adam_steps": [tr.step_count, one.step_count], "tolerance": 1e-6}
        ok = d.max().item() <= 1e-6 and spread.item() == 0.0 and tr.step_count == one.step_count == args.steps
        result["ok"] = bool(ok)
        print(json.dumps(result))
        if os.path.isdir(os.path.dirname(args.out)):
            json.dump(result, open(args.out, "w"), indent=1)
    flag = torch.tensor([1 if (result is None or result["ok"]) else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
