"""`python -m vcmeshconv_b200.train config_train.config` — the reference's training entry point
(`python graphVAE_train.py config`, GraphAutoEncoder/graphVAE_train.py:36-58, 265-353) on this package:

  config (INI, graphAE_param_iso.Parameters)  ->  template faces (:265-268), frame list (:54), MeshDataset (:76-123)
  Net_autoenc (:132-249, channels hard-coded there as here)  ->  DataParallelTrainer (one process per GPU; launch
  with torchrun for several)  ->  fit(): the loop of :303-353 with host-fed batches.

The global batch is `param.batch` meshes per GPU (:294); every rank decodes the same shuffled order and takes its
own slice, so the job sees each frame once per epoch exactly like the reference's DataLoader. Printing follows the
reference (every 20 iterations, :318-327); tensorboard is not written (out of scope).
"""
import sys

import numpy as np
import torch

from . import ply_dataset
from .graphAE_param_iso import Parameters


def read_frame_list(fn):
    """graphVAE_train.py:54: one frame id per line (np.genfromtxt(..., dtype=str))."""
    with open(fn) as f:
        return [tok for line in f for tok in line.split()]


def build_from_config(config_fn, make_dirs=True, enc_channels=None, dec_channels=None, weight_num=None):
    """(param, model on the CPU, dataset) from a reference-format config; nothing here needs a GPU."""
    from . import graphVAE_train as T
    param = Parameters().read_config(config_fn, make_dirs=make_dirs)
    faces = ply_dataset.load_template_faces(param.template_ply_fn)
    dataset = ply_dataset.MeshDataset(param.mesh_train, read_frame_list(param.framelist), faces)
    kw = {} if weight_num is None else {"weight_num": weight_num}
    model = T.Net_autoenc(param, faces.numpy(), enc_channels=enc_channels, dec_channels=dec_channels, **kw)
    return param, model, dataset


def rank_slice(pc, nor, rank, world):
    """This rank's contiguous share of a global batch (graphVAE_train.py:294: batch * number of devices)."""
    per = pc.shape[0] // world
    return pc[rank * per:(rank + 1) * per], nor[rank * per:(rank + 1) * per]


def epochs(dataset, trainer, per_gpu_batch, rank, world, n_epochs=10000, seed=0):
    """Iterable of epochs of device batches for DataParallelTrainer.fit: every batch is decoded on the host, pinned
    and copied asynchronously (the copy is queued behind the running step; DataParallelTrainer.prefetch /
    step_prefetched is the variant that overlaps it with the step on a copy stream)."""
    dev = trainer.flat.flat_p.device
    for epoch in range(n_epochs):
        def one_epoch(epoch=epoch):
            for pc, nor in ply_dataset.host_batches(dataset, per_gpu_batch * world, shuffle=True, seed=seed + epoch):
                pc, nor = rank_slice(pc, nor, rank, world)
                yield pc.to(dev, non_blocking=True), nor.to(dev, non_blocking=True)
        yield one_epoch()


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        raise SystemExit("usage: python -m vcmeshconv_b200.train <config_train.config>")
    from . import graphVAE_train as T
    if not torch.cuda.is_available():
        raise SystemExit("training needs a CUDA device: vcmeshconv_b200 has no CPU path")
    rank, local, world = T.init_distributed()
    param, model, dataset = build_from_config(argv[0], make_dirs=rank == 0)
    model = model.to(torch.device("cuda", local))
    trainer = T.DataParallelTrainer(model, lr=param.lr)
    if getattr(param, "read_weight_path", ""):
        trainer.load(param.read_weight_path)
        if world > 1:
            trainer.broadcast_parameters()

    def log(iteration, loss, parts):
        if iteration % 20 == 0 and rank == 0:                # graphVAE_train.py:318-327
            l1, lap, kl, nor = (float(p) for p in parts)
            print("###Iteration", iteration)
            print("loss_pose_l1:", l1)
            print("loss_laplace_l1:", lap)
            print("loss_kl:", kl)
            print("normal_loss:", nor)
            print("loss:", float(loss))
            print(trainer.lr)

    batches = epochs(dataset, trainer, param.batch, rank, world)
    if world == 1 and len(dataset) >= param.batch:
        # single GPU: the step is replayed from one CUDA graph (captured on the shapes of a first batch, without
        # warm-up steps, so no parameter is touched before iteration start_iter); several GPUs run the eager step
        # with the bucketed all-reduce overlapping the backward pass
        pc0, nor0 = next(iter(next(iter(epochs(dataset, trainer, param.batch, rank, world, n_epochs=1)))))
        trainer.capture(pc0, nor0, warmup=0)
    trainer.fit(param, batches, log=log)


if __name__ == "__main__":
    main()
