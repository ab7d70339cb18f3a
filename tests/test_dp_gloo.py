"""Host-side logic of the data-parallel trainer (flat gradient buffer, buckets,
hook-driven all-reduce with SUM semantics) on CPU with gloo, world_size 2.
The model is a stand-in with the same interface as Net_autoenc (the real layers
are CUDA-only); what is under test is the reduction machinery."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from vcmeshconv_b200.graphVAE_train import DataParallelTrainer, FlatBuffers


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = nn.Parameter(torch.randn(37, 5))
        self.b = nn.Parameter(torch.randn(5))
        self.c = nn.Parameter(torch.randn(5, 3))
        self.unused = nn.Parameter(torch.randn(7))       # never receives a gradient

    def trainable_parameters(self):
        return [self.c, self.b, self.a, self.unused]

    def losses(self, x, nor, eps=None):
        loss = ((x @ self.a + self.b).tanh() @ self.c).pow(2).mean()     # per-replica MEAN loss
        z = loss.detach()
        return loss, z, z, z, z, None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, overlap, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = Toy()
        if rank == 1:                                     # ranks start different: broadcast must fix it
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        tr = DataParallelTrainer(model, bucket_bytes=256, overlap=overlap is True)
        torch.manual_seed(100 + rank)
        x = torch.randn(8, 37)
        tr.forward_backward(x, None)
        if overlap == "split":                            # a range already reduced "early": only the rest is left
            tr._allreduce(0, 64)
            tr._early_upto = 64
        tr._finish_reduction()
        out[rank] = (tr.flat.flat_g.clone(), tr.flat.flat_p.clone(), x)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False, "split"])
def test_sum_allreduce_matches_single_process(overlap):
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), overlap, out), nprocs=world, join=True)
        g0, p0, x0 = out[0]
        g1, p1, x1 = out[1]
    assert torch.equal(p0, p1)                            # parameters synchronised from rank 0
    assert torch.equal(g0, g1)                            # every rank holds the same reduced gradient
    # single-process reference: SUM over replicas of per-replica mean losses (graphVAE_train.py:249,315)
    model = Toy()
    flat = FlatBuffers(model.trainable_parameters(), 256)
    (model.losses(x0, None)[0] + model.losses(x1, None)[0]).backward()
    assert torch.allclose(flat.flat_g, g0, rtol=1e-6, atol=1e-7)
    assert len(flat.buckets) > 1                          # the bucket path was exercised


def test_flat_buffers_views():
    m = Toy()
    before = [p.detach().clone() for p in m.trainable_parameters()]
    flat = FlatBuffers(m.trainable_parameters(), 128)
    for p, b, o in zip(m.trainable_parameters(), before, flat.offsets):
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() == flat.flat_p.data_ptr() + 4 * o and p.grad.data_ptr() == flat.flat_g.data_ptr() + 4 * o
        assert o % 64 == 0
    covered = sorted(i for _, _, idx in flat.buckets for i in idx)
    assert covered == list(range(len(before)))
    assert flat.buckets[0][0] == 0 and flat.buckets[-1][1] == flat.total
