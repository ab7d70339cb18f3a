"""Training step of the mesh autoencoder, batch data-parallel over B200s.

Mirrors the reference's GraphAutoEncoder/graphVAE_train.py for the part that
is on the hot path (e):

  Net_autoenc           graphVAE_train.py:132-249   same constructor, same
                        sub-module names (checkpoint keys), same forward
                        signature and 5-tuple of [1]-shaped losses
  DataParallelTrainer   replaces nn.DataParallel (:277) + Adam (:252-262, :316):
                        one process per GPU, parameters resident on every GPU,
                        gradients produced directly inside ONE flat fp32 buffer,
                        laid out in the order in which they become final during
                        the backward pass and summed over the ranks (SUM - the
                        reference sums per-replica mean losses, :249/:315)
                        segment by segment while the backward pass is still
                        running: in graph mode by symmetric-memory all-reduce
                        kernels over NVLink peer memory captured in the step
                        graph, eagerly by bucketed NCCL all-reduces launched from
                        autograd hooks; a fused Adam launch per segment on every
                        rank. The whole step replays as one CUDA graph.

The INI reader is graphAE_param_iso.py, the PLY data path ply_dataset.py, the
entry point train.py; tensorboard is out of scope. `make_param`
below builds the fields the model needs from in-memory tables.
"""
import os
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from . import functional as F_
from . import graphVAESSW as vae_model

ENC_CHANNELS = [3, 32, 64, 128, 256, 512, 64]      # graphVAE_train.py:151 (hard-coded there)
DEC_CHANNELS = [64, 512, 256, 128, 64, 32, 3]      # graphVAE_train.py:172
WEIGHT_NUM = 17                                    # graphVAE_train.py:136


def make_param(enc_tables, dec_tables, table0, point_num, residual_rate=0.9, w_pose=1.0, w_laplace=100.0, lr=1e-4,
               batch=16, conv_max=0, write_tmp_folder=""):
    """The fields of the reference's `Parameters` object (graphAE_param_iso.py:21-134)
    that the model reads; tables may be arrays or .npy paths."""
    t0 = np.load(table0) if isinstance(table0, str) else np.asarray(table0)
    return types.SimpleNamespace(
        point_num=point_num, residual_rate=residual_rate, conv_max=conv_max, w_pose=w_pose, w_laplace=w_laplace,
        lr=lr, batch=batch, write_tmp_folder=write_tmp_folder,
        connection_layer_fn_lst_enc=list(enc_tables), connection_layer_fn_lst_dec=list(dec_tables),
        neighbor_id_lstlst=t0[:, 1:].reshape(t0.shape[0], -1, 2)[:, :, 0], neighbor_num_lst=t0[:, 0])


class Net_autoenc(nn.Module):
    def __init__(self, param, facedata, enc_channels=None, dec_channels=None, weight_num=WEIGHT_NUM):
        super().__init__()
        self.weight_num = weight_num
        self.motdim = 94
        enc_channels = ENC_CHANNELS if enc_channels is None else enc_channels
        dec_channels = DEC_CHANNELS if dec_channels is None else dec_channels
        self.mcstructureenc = vae_model.MCStructure(param, param.point_num, weight_num, bDec=False)
        self.mcvcoeffsenc = vae_model.MCVcoeffs(self.mcstructureenc, weight_num)
        self.net_geoenc = vae_model.MCEnc(self.mcstructureenc, enc_channels, weight_num)
        self.nrpt_latent = self.net_geoenc.out_nrpts
        self.mcstructuredec = vae_model.MCStructure(param, self.nrpt_latent, weight_num, bDec=True)
        self.mcvcoeffsdec = vae_model.MCVcoeffs(self.mcstructuredec, weight_num)
        self.net_geodec = vae_model.MCDec(self.mcstructuredec, dec_channels, weight_num)
        self.net_loss = vae_model.MCLoss(param)
        faces = np.asarray(facedata).astype(np.int64)
        self.register_buffer("t_facedata", torch.from_numpy(faces).long())
        # the three face-corner gathers of the normal loss (graphVAE_train.py:213-217) as ONE gather/contract
        # launch: a table with one row per face listing its 3 vertices, M = 1, coefficients 1
        ftab = np.zeros((faces.shape[0], 7), np.int32)
        ftab[:, 0] = 3
        ftab[:, 1::2] = faces
        self._face_table_host = ftab
        self._face_tables = {}
        self.register_buffer("_face_coeffs", torch.ones(faces.shape[0], 3, 1), persistent=False)
        self.w_pose = param.w_pose
        self.w_laplace = param.w_laplace
        self.klweight = 1e-5
        self.w_nor = 10.0
        self.write_tmp_folder = getattr(param, "write_tmp_folder", "")
        # called with d loss / d latent when the decoder's backward pass is complete (set by the trainer: from
        # that moment the decoder-side gradients are final and can be reduced / applied while the encoder's
        # backward pass still runs)
        self.latent_grad_hook = None

    def losses(self, in_pc_batch, t_nor, t_eps=None, fused=None):
        """graphVAE_train.py:199-225. `t_eps` lets a caller fix the N(0,1) draw. `fused` (default: on, env
        VCM_FUSED_LOSS=0 turns it off) runs the reparameterisation + KL and the three geometric losses as the
        library's loss-head kernels (4 launches forward, 2 backward) instead of ~110 ATen elementwise / reduction
        kernels around the M = 1 gathers; both paths compute the same quantities."""
        if fused is None:
            fused = os.environ.get("VCM_FUSED_LOSS", "1") == "1"
        if fused and in_pc_batch.shape[2] == 3:
            enc = self.net_geoenc
            raw_mu, raw_logstd = enc.forward_raw(in_pc_batch, self.mcvcoeffsenc)
            if t_eps is None:
                t_eps = torch.empty_like(raw_mu).normal_()
            t_z, klloss = F_.reparam_kl(raw_mu, raw_logstd, t_eps, enc.MU_SCALE, enc.LOGSTD_SCALE)
            if self.latent_grad_hook is not None and t_z.requires_grad:
                t_z.register_hook(self.latent_grad_hook)
            out_pc_batch = self.net_geodec(t_z, self.mcvcoeffsdec)[:, :, 0:3]
            dev = in_pc_batch.device
            geo, parts = F_.loss_head(out_pc_batch, in_pc_batch, t_nor, self.net_loss._lap_coeffs,
                                      self.net_loss.tables(dev), self._face_tables_on(dev), self.w_pose,
                                      self.w_laplace, self.w_nor)
            loss = geo + klloss * self.klweight
            return loss, parts[0], parts[1], klloss, parts[2], out_pc_batch
        t_mu, t_logstd = self.net_geoenc(in_pc_batch, self.mcvcoeffsenc)
        t_std = t_logstd.exp()
        if t_eps is None:
            t_eps = torch.ones_like(t_std).normal_()
        t_z = t_mu + t_std * t_eps
        if self.latent_grad_hook is not None and t_z.requires_grad:
            t_z.register_hook(self.latent_grad_hook)
        klloss = torch.mean(-0.5 - t_logstd + 0.5 * t_mu ** 2 + 0.5 * torch.exp(2 * t_logstd))
        out_pc_batch = self.net_geodec(t_z, self.mcvcoeffsdec)[:, :, 0:3]
        dif_pos = out_pc_batch - in_pc_batch
        tri = F_.gather_contract(dif_pos.contiguous(), self._face_coeffs, self._face_tables_on(dif_pos.device))
        loss_normal = (tri / 3.0 * t_nor).sum(2).pow(2).mean()
        loss_pose_l1 = self.net_loss.compute_geometric_loss_l1(in_pc_batch[:, :, 0:3], out_pc_batch)
        loss_laplace = self.net_loss.compute_laplace_loss_l2(in_pc_batch[:, :, 0:3], out_pc_batch)
        loss = (loss_pose_l1 * self.w_pose + loss_laplace * self.w_laplace + klloss * self.klweight +
                loss_normal * self.w_nor)
        return loss, loss_pose_l1, loss_laplace, klloss, loss_normal, out_pc_batch

    def _face_tables_on(self, device):
        key = device.index if device.index is not None else torch.cuda.current_device()
        t = self._face_tables.get(key)
        if t is None:
            from .connectivity import DeviceTables
            t = self._face_tables[key] = DeviceTables(self._face_table_host, self.mcstructureenc.point_num,
                                                      torch.device("cuda", key))
        return t

    def forward(self, in_pc_batch, iteration=0, frame=None, t_idx=None, t_nor=None, bDebug=False):
        loss, l1, lap, kl, nor, _ = self.losses(in_pc_batch, t_nor)
        return loss[None], l1[None], lap[None], kl[None], nor[None]

    def trainable_parameters(self):
        """Same parameter set and order as the reference's getoptim (graphVAE_train.py:252-262)."""
        import itertools
        return [p for p in itertools.chain(self.net_geodec.parameters(), self.net_geoenc.parameters(),
                                           self.mcvcoeffsdec.parameters(), self.mcvcoeffsenc.parameters())
                if p.requires_grad]


class FlatBuffers:
    """All trainable parameters (and their gradients) as views into two flat
    fp32 buffers. Layout follows backward order (decoder last layer first) so a
    bucket is a contiguous slice that becomes ready as a whole."""

    def __init__(self, params, bucket_bytes=32 << 20, grad_alloc=None, param_alloc=None):
        """`grad_alloc(n)` / `param_alloc(n)` may supply the gradient / parameter buffer (e.g. symmetric-memory
        allocations that peer GPUs can address); default torch.zeros."""
        params = list(params)
        dev = params[0].device
        # autograd reaches parameters roughly in reverse creation order of use; bucket order only
        # affects overlap, never results
        self.params = params
        sizes = [p.numel() for p in params]
        pad = lambda n: (n + 63) // 64 * 64               # 256-byte aligned views (vectorised kernels)
        offs, total = [], 0
        for n in sizes:
            offs.append(total)
            total += pad(n)
        self.total = total
        self.flat_p = (param_alloc(total).zero_() if param_alloc is not None else
                       torch.zeros(total, device=dev, dtype=torch.float32))
        self.flat_g = grad_alloc(total).zero_() if grad_alloc is not None else torch.zeros(total, device=dev,
                                                                                            dtype=torch.float32)
        self.offsets = offs
        for p, o in zip(params, offs):
            n = p.numel()
            self.flat_p[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + n].view(p.shape)
            p.grad = self.flat_g[o:o + n].view(p.shape)
        # buckets: consecutive parameters up to bucket_bytes
        self.buckets, cur, cur_bytes, start = [], [], 0, 0
        for i, n in enumerate(sizes):
            cur.append(i)
            cur_bytes += pad(n) * 4
            if cur_bytes >= bucket_bytes or i == len(sizes) - 1:
                end = offs[i] + pad(n)
                self.buckets.append((start, end, list(cur)))
                start, cur, cur_bytes = end, [], 0

    def zero_grad(self):
        self.flat_g.zero_()


class DataParallelTrainer:
    """One rank of the batch-data-parallel training job.

    step(pc, nor) = forward + backward + gradient all-reduce (SUM) + Adam.
    world_size 1 needs no process group. Works with any torch.distributed
    backend (NCCL on the GPUs; gloo in the CPU tests of the bucket logic)."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, bucket_bytes=32 << 20, process_group=None,
                 overlap=True):
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        # flat layout: decoder side (layers, then coefficients) first, encoder side after it, so that "everything
        # whose gradient is final when the decoder's backward pass ends" is one contiguous range [0, early_split)
        # (the flat buffers are laid out in the order in which gradients become final: decoder side, then the late
        # encoder layers and heads, then the first encoder layers, whose backward pass ends the step)
        segs, self._enc_hook_layer = self._split_params(model)
        ordered = [p for seg in segs for p in seg]
        self._symm_op, grad_alloc = self._symmetric_allreduce(ordered)
        try:
            self.flat = FlatBuffers(ordered, bucket_bytes, grad_alloc, grad_alloc if self._symm_op == "fused" else None)
            if self._symm_op is not None:
                hg = self._symm_mem.rendezvous(self.flat.flat_g, self._symm_group)
                if self._symm_op == "fused":
                    self._setup_fused(hg)
        except Exception as e:                   # no symmetric memory on this system: NCCL between two graphs
            if grad_alloc is None:
                raise
            import sys
            sys.stderr.write("vcmeshconv_b200: symmetric-memory reduction unavailable (%r), using NCCL\n" % (e,))
            self._symm_op = None
            self.flat = FlatBuffers(ordered, bucket_bytes)
        # early_bounds[i] = end of segment i in the flat buffers; the last segment is never "early"
        ends, n = [], 0
        for seg in segs:
            n += len(seg)
            ends.append(self.flat.offsets[n] if n < len(ordered) else self.flat.total)
        self.early_bounds = ends[:-1]
        self.early_split = self.early_bounds[0] if self.early_bounds else 0
        self.exp_avg = torch.zeros_like(self.flat.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat.flat_p)
        # optimizer scalars live on the device so a captured step can be replayed: {step, lr, 1-b1^t, sqrt(1-b2^t)}
        self.opt_state = torch.tensor([0.0, lr, 0.0, 0.0], device=self.flat.flat_p.device, dtype=torch.float32)
        self._graph = None
        self._copy_stream = None
        # dW GEMMs on a second stream (parallel graph branch); VCM_DW_STREAM=0 turns it off
        self.dw_stream = (torch.cuda.Stream(self.flat.flat_p.device)
                          if self.flat.flat_p.is_cuda and os.environ.get("VCM_DW_STREAM", "1") == "1" else None)
        # parameter-only preparation (stacked bases of the dz GEMMs) as a branch from the start of the step
        self.prep_stream = (torch.cuda.Stream(self.flat.flat_p.device)
                            if self.flat.flat_p.is_cuda and os.environ.get("VCM_PREP_STREAM", "1") == "1" else None)
        # opt-in (VCM_SPLIT_K5=1): the dA half of K5 as a parallel branch. Measured slower on C2 (2.75 vs 2.63 ms):
        # dz is read twice and the extra kernels slow the chain kernels they run next to.
        self.da_stream = (torch.cuda.Stream(self.flat.flat_p.device)
                          if self.flat.flat_p.is_cuda and os.environ.get("VCM_SPLIT_K5", "0") == "1" else None)
        # residual branch of every layer on its own stream (forward, and through autograd its backward)
        # Stream priorities (VCM_STREAM_PRIO=0 turns them off): the backward chain and the residual branches that rejoin
        # it run at high priority, the branches nothing waits for until the optimizer (dW GEMMs, dA halves, stacked
        # bases, early Adam) at the default one, so that a wide off-chain kernel does not keep the chain's next kernel
        # out of the SMs. Captured kernel nodes inherit the priority of the stream they were captured on.
        self._prio = int(os.environ.get("VCM_PRIO_MAIN", "-1")) if os.environ.get("VCM_STREAM_PRIO", "1") == "1" else 0
        # the residual branch rejoins the chain in every layer's epilogue and is the shorter one: one level above the
        # chain (measured on C2: 2.257 -> 2.244 ms; the other way round 2.260)
        self._prio_res = int(os.environ.get("VCM_PRIO_RES", str(self._prio - 1))) if self._prio else 0
        self.res_stream = (torch.cuda.Stream(self.flat.flat_p.device, priority=self._prio_res)
                           if self.flat.flat_p.is_cuda and os.environ.get("VCM_RES_STREAM", "1") == "1" else None)
        # the std head of the encoder next to the mu head (forward and backward); VCM_HEAD_STREAM=0 turns it off
        self.head_stream = (torch.cuda.Stream(self.flat.flat_p.device, priority=self._prio)
                            if self.flat.flat_p.is_cuda and os.environ.get("VCM_HEAD_STREAM", "1") == "1" else None)
        # decoder-side all-reduce + Adam next to the encoder's backward pass (VCM_EARLY_UPDATE=0 turns it off)
        self.early_stream = (torch.cuda.Stream(self.flat.flat_p.device,
                                               priority=int(os.environ.get("VCM_PRIO_EARLY", "0")))
                             if self.flat.flat_p.is_cuda and os.environ.get("VCM_EARLY_UPDATE", "1") == "1" else None)
        self._advanced = self._early_done = False
        self._early_upto = 0
        self.overlap = overlap and self.world > 1 and self._symm_op != "fused"
        self._seg = 0                                # which range of the step the next fused reduction covers
        self._pending = []
        self._handles = []
        if self.overlap:
            self._install_hooks()
        self._direct = not self.overlap              # see forward_backward
        if self.world > 1:
            self.broadcast_parameters()

    def _symmetric_allreduce(self, params):
        """The flat gradient buffer of a multi-GPU job lives in a symmetric-memory allocation (every GPU of the node
        maps every peer's buffer over NVLink; `multimem` additionally through the NVSwitch multicast object) and is
        reduced by torch's symmetric-memory kernels (VCM_SYMM_AR=two_shot, the default | multimem | 0 = NCCL). They
        are plain kernels with device-side barriers, so - unlike c10d's NCCL calls, whose capture hung here - they
        are captured inside the step graph: the gradient segments are reduced on the early branch next to the rest
        of the backward pass and the whole multi-GPU step is ONE graph launch (8 x B200: 2.62 vs 2.80 ms/step with
        NCCL between two graphs). Returns (op name or None, allocator or None); the caller falls back to NCCL when
        the allocation is not available."""
        mode = os.environ.get("VCM_SYMM_AR", "fused")
        if mode in ("0", "") or self.world == 1 or not params or not params[0].is_cuda:
            return None, None
        if self.world > 8 and mode == "fused":
            mode = "two_shot"
        import torch.distributed._symmetric_memory as symm_mem
        op = {"fused": "fused", "two_shot": "two_shot_all_reduce_", "multimem": "multimem_all_reduce_"}[mode]
        self._symm_mem = symm_mem
        self._symm_group = (self.group if self.group is not None else dist.group.WORLD).group_name
        dev = params[0].device
        return op, (lambda n: symm_mem.empty(n, dtype=torch.float32, device=dev))

    def _setup_fused(self, hg):
        """VCM_SYMM_AR=fused (the multi-GPU default): the repo's own reduce + Adam + broadcast kernel
        (csrc/collective.cu, vcm_reduce_adam_bcast) over the peers' gradient, parameter and flag buffers, all three
        in symmetric memory. Collects every rank's base pointers as this process maps them."""
        import ctypes as C
        dev = self.flat.flat_p.device
        hp = self._symm_mem.rendezvous(self.flat.flat_p, self._symm_group)
        self._flags = self._symm_mem.empty(8 * 3 * 128, dtype=torch.int32, device=dev).zero_()
        hf = self._symm_mem.rendezvous(self._flags, self._symm_group)
        arr = C.c_void_p * self.world

        def peers(h, tensor):
            ptrs = [int(h.buffer_ptrs[k]) + (int(tensor.data_ptr()) - int(h.buffer_ptrs[self.rank])) for k in range(self.world)]
            return arr(*ptrs)
        self._peer_g, self._peer_p, self._peer_f = peers(hg, self.flat.flat_g), peers(hp, self.flat.flat_p), peers(hf, self._flags)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)               # every rank's flags are zero before anyone signals

    def _reduce_update(self, lo, hi):
        """Fused path: SUM over ranks of flat_g[lo:hi] -> Adam on the chunks this rank owns -> new parameters into
        every rank's flat_p. One launch; every rank issues the same ranges in the same order."""
        if hi <= lo:
            return
        f = self.flat
        _lib.call("vcm_reduce_adam_bcast", self._peer_g, self._peer_p, self._peer_f, self.exp_avg.data_ptr(),
                  self.exp_avg_sq.data_ptr(), lo, hi - lo, self.opt_state.data_ptr(), self.betas[0], self.betas[1],
                  self.eps, self.rank, self.world, 128 * (self._seg % 3), _lib.stream_ptr(),
                  nbytes=4 * (hi - lo) * (2 * self.world + 5) // self.world, flops=12 * (hi - lo) // self.world)
        self._seg += 1

    def _allreduce(self, lo, hi):
        """SUM over ranks of flat_g[lo:hi], in place, on the current stream."""
        if hi <= lo:
            return
        if self._symm_op == "fused":
            raise RuntimeError("the fused path reduces inside _reduce_update")
        if self._symm_op is not None:
            getattr(torch.ops.symm_mem, self._symm_op)(self.flat.flat_g[lo:hi], "sum", self._symm_group)
        else:
            dist.all_reduce(self.flat.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.group)

    @staticmethod
    def _split_params(model):
        """Trainable parameters as segments in the order in which their gradients become final during the backward
        pass: [decoder layers + decoder coefficients], [encoder layers >= k, both heads, their coefficients],
        [encoder layers < k and their coefficients]. Returns (segments, k); models without that structure are one
        segment."""
        params = model.trainable_parameters()
        need = ("net_geodec", "mcvcoeffsdec", "net_geoenc", "mcvcoeffsenc")
        if not all(hasattr(model, a) for a in need):
            return [params], None
        enc = model.net_geoenc
        k = min(2, enc.layer_num - 1)
        ident = lambda ps: {id(p) for p in ps}
        dec_ids = ident(model.net_geodec.parameters()) | ident(model.mcvcoeffsdec.parameters())
        late_ids = ident(enc.stdlayer.parameters())
        for i in range(k, enc.layer_num):
            late_ids |= ident(enc.layer_lst[i].parameters()) | {id(model.mcvcoeffsenc.vcoeffs_list[i])}
        segs = [[p for p in params if id(p) in dec_ids],
                [p for p in params if id(p) in late_ids and id(p) not in dec_ids],
                [p for p in params if id(p) not in dec_ids and id(p) not in late_ids]]
        segs = [seg for seg in segs if seg]
        return segs, (k if len(segs) == 3 else None)

    # -- communication ------------------------------------------------------------
    def broadcast_parameters(self):
        """Rank 0's initial parameters everywhere (the reference replicates from GPU 0 every
        forward; here they are resident and only synchronised once)."""
        dist.broadcast(self.flat.flat_p, 0, group=self.group)

    def _install_hooks(self):
        owner = {}
        for b, (_, _, idxs) in enumerate(self.flat.buckets):
            for i in idxs:
                owner[i] = b
        self._remaining = [len(idxs) for _, _, idxs in self.flat.buckets]
        self._count = list(self._remaining)
        for i, p in enumerate(self.flat.params):
            def hook(param, i=i):
                if not self.overlap:
                    return
                b = owner[i]
                self._count[b] -= 1
                if self._count[b] == 0:
                    self._launch_bucket(b)
            p.register_post_accumulate_grad_hook(hook)

    def _launch_bucket(self, b):
        s, e, _ = self.flat.buckets[b]
        self._handles.append(dist.all_reduce(self.flat.flat_g[s:e], op=dist.ReduceOp.SUM, group=self.group,
                                             async_op=True))

    def _finish_reduction(self):
        if self.world == 1:
            return
        if self.overlap:
            for b, c in enumerate(self._count):          # parameters that received no gradient this step
                if c != 0:
                    self._launch_bucket(b)
            self._count = list(self._remaining)
            for h in self._handles:
                h.wait()
            self._handles = []
        elif self._symm_op != "fused":                              # fused: optimizer_step reduces what is left
            self._allreduce(self._early_upto, self.flat.total)      # [0, _early_upto) was reduced next to the backward pass

    # -- one training step ---------------------------------------------------------
    def _set_direct_grads(self, on):
        """Backward kernels write parameter gradients straight into the flat buffer (no autograd
        accumulate kernels). Needs no per-parameter hooks, so it is off while bucketed overlap is on."""
        for p in self.flat.params:
            p._vcm_grad_buf = p.grad if on else None

    def forward_backward(self, pc, nor, eps=None, early=False):
        """zero_grad + forward + backward. With `early` (step() and the captured step use it) the decoder-side range
        of the flat buffers is all-reduced and its Adam update applied on `early_stream` as soon as the decoder's
        backward pass is complete, next to the encoder's backward pass; the caller finishes the step with
        _finish_reduction() + optimizer_step(), which then only touch the rest."""
        cur = torch.cuda.current_stream() if self.flat.flat_p.is_cuda else None
        if self.prep_stream is not None:
            # nothing reads or writes the gradient buffer before the backward pass: clear it next to the forward pass
            # (the preparation stream is joined before backward starts)
            self.prep_stream.wait_stream(cur)
            with torch.cuda.stream(self.prep_stream):
                self.flat.zero_grad()
        else:
            self.flat.zero_grad()
        # direct gradient writes (backward kernels fill the flat buffer in place, OVERWRITING it) are enabled for the
        # duration of this call only: a backward() the user runs outside goes through ordinary autograd accumulation
        self._set_direct_grads(self._direct)
        F_.begin_step(self.flat.params)
        self._early_done = False
        cur = torch.cuda.current_stream() if self.flat.flat_p.is_cuda else None
        self._early_upto = 0
        if early and self.early_stream is not None and self.early_bounds and not self.overlap:
            self._advance()                              # this step's optimizer scalars, before anything applies them
            self.early_stream.wait_stream(cur)
            self.model.latent_grad_hook = lambda grad: self._segment_done(0)
            if self._enc_hook_layer is not None:
                self.model.net_geoenc.activation_hooks = {self._enc_hook_layer: lambda grad: self._segment_done(1)}
        if self.res_stream is not None and not self.overlap:
            F_.set_residual_stream(self.res_stream)
        if self.head_stream is not None and not self.overlap:
            F_.set_head_stream(self.head_stream)
        if self.prep_stream is not None:
            self.prep_stream.wait_stream(cur)
            F_.set_prep_stream(self.prep_stream)
        try:
            loss, l1, lap, kl, lnor, _ = self.model.losses(pc, nor, eps)
        finally:
            F_.set_residual_stream(None)
            F_.set_head_stream(None)
            F_.set_prep_stream(None)
            if self.prep_stream is not None:
                cur.wait_stream(self.prep_stream)
        if not self.overlap:
            F_.set_side_stream(self.dw_stream)
            F_.set_da_stream(self.da_stream)
        try:
            loss.backward()
        finally:
            F_.join_side_stream()
            F_.set_side_stream(None)
            F_.set_da_stream(None)
            if self.res_stream is not None and cur is not None:
                cur.wait_stream(self.res_stream)     # residual-branch backward kernels (p_neighbors / weight_res grads)
            if self.head_stream is not None and cur is not None:
                cur.wait_stream(self.head_stream)    # the std head's backward kernels
            self._set_direct_grads(False)
            if getattr(self.model, "latent_grad_hook", None) is not None:
                self.model.latent_grad_hook = None
                self.model.net_geoenc.activation_hooks = {}
                cur.wait_stream(self.early_stream)
        return loss.detach(), (l1.detach(), lap.detach(), kl.detach(), lnor.detach())

    def _segment_done(self, i):
        """Tensor hook (on the latent code for segment 0, on the input of encoder layer k for segment 1): autograd
        runs nodes in reverse creation order, so when the gradient of that activation is complete every backward
        kernel of the layers behind it has been enqueued (on the chain or on a branch stream) and the gradients of
        segment i are final: reduce them over the ranks and apply Adam on the early stream."""
        lo, hi = self._early_upto, self.early_bounds[i]
        s = self.early_stream
        s.wait_stream(torch.cuda.current_stream())
        capturing = torch.cuda.is_current_stream_capturing()
        for st in (F_._side_stream, F_._da_stream, self.res_stream, self.head_stream):
            if st is None:
                continue
            if capturing:
                with torch.cuda.stream(st):              # a branch stream that got no work inside this capture has
                    if not torch.cuda.is_current_stream_capturing():    # nothing to wait for (and may not be waited on)
                        continue
            s.wait_stream(st)
        with torch.cuda.stream(s):
            if self.world > 1 and self._symm_op == "fused":
                self._reduce_update(lo, hi)
            else:
                if self.world > 1:
                    self._allreduce(lo, hi)
                self._apply(lo, hi)
        self._early_upto = hi
        self._early_done = True

    @property
    def step_count(self):
        return int(self.opt_state[0].item())

    def set_lr(self, lr):
        """ExponentialLR of the reference (graphVAE_train.py:281,353) = one scalar write per epoch."""
        self.lr = lr
        self.opt_state[1] = lr

    def _advance(self):
        _lib.call("vcm_adam_advance", self.opt_state.data_ptr(), self.betas[0], self.betas[1], _lib.stream_ptr())
        self._advanced = True

    def _apply(self, lo, hi):
        f = self.flat
        if hi > lo:
            o = 4 * lo
            _lib.call("vcm_adam_apply", f.flat_p.data_ptr() + o, f.flat_g.data_ptr() + o, self.exp_avg.data_ptr() + o,
                      self.exp_avg_sq.data_ptr() + o, hi - lo, self.opt_state.data_ptr(), self.betas[0], self.betas[1],
                      self.eps, 1.0, _lib.stream_ptr(), nbytes=28 * (hi - lo), flops=12 * (hi - lo))

    def optimizer_step(self):
        """Adam over the flat buffers (the part an early update has not already covered in this step)."""
        f = self.flat
        if not f.flat_p.is_cuda:
            raise RuntimeError("the optimizer step runs on CUDA only (no CPU fallback)")
        if not self._advanced:
            self._advance()
        if self.world > 1 and self._symm_op == "fused":
            self._reduce_update(self._early_upto, f.total)
        else:
            self._apply(self._early_upto, f.total)
        self._advanced = self._early_done = False
        self._early_upto = 0
        self._seg = 0

    def step(self, pc, nor, eps=None):
        if self._graph is not None and (eps is None or self.static_eps is not None):
            return self._graph_step(pc, nor, eps)
        loss, parts = self.forward_backward(pc, nor, eps, early=True)
        self._finish_reduction()
        self.optimizer_step()
        return loss, parts

    # -- host-fed steps: the next batch is staged while the current step runs ---------------
    def prefetch(self, host_pc, host_nor):
        """Starts the host -> device copy of the NEXT batch (pinned host tensors) on a copy stream; the batch is
        consumed by the following `step_prefetched()`. Two staging slots alternate, so the copy overlaps the step
        that is running."""
        dev = self.flat.flat_p.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
            self._stage = [(torch.empty(host_pc.shape, device=dev), torch.empty(host_nor.shape, device=dev))
                           for _ in range(2)]
            self._stage_free = [None, None]
            self._stage_i = 0
        slot = self._stage_i
        self._stage_i ^= 1
        buf = self._stage[slot]
        with torch.cuda.stream(self._copy_stream):
            if self._stage_free[slot] is not None:
                self._copy_stream.wait_event(self._stage_free[slot])     # the step that read this slot is done
            buf[0].copy_(host_pc, non_blocking=True)
            buf[1].copy_(host_nor, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        self._staged = (slot, ready)

    def step_prefetched(self):
        slot, ready = self._staged
        cur = torch.cuda.current_stream()
        cur.wait_event(ready)
        out = self.step(*self._stage[slot])
        done = torch.cuda.Event()
        done.record(cur)
        self._stage_free[slot] = done
        return out

    # -- CUDA-graph capture of the step -----------------------------------------------
    def capture(self, pc, nor, warmup=3):
        """Captures the step into CUDA graphs; tables and shapes are static, so every later step() is one graph
        launch (world_size 1: zero_grad, forward, backward, Adam) or two with the NCCL all-reduce between them,
        instead of ~250 kernel launches from Python. The captured graph has parallel branches: residual branches,
        dW GEMMs, the stacked-bases preparation and the early decoder-side Adam."""
        assert pc.is_cuda and self._graph is None
        self.overlap = False                         # no NCCL launches inside the capture
        self._direct = True
        self.static_pc, self.static_nor = pc.clone(), nor.clone()
        # the N(0,1) draw of the reparameterisation (graphVAE_train.py:201) stays outside the graph: one normal_()
        # launch in front of every replay, from torch's generator, instead of a captured generator (whose replay
        # costs two extra fill kernels for seed and offset)
        self.static_eps = None
        if os.environ.get("VCM_EPS_OUTSIDE", "1") == "1":
            m = self.model
            self.static_eps = torch.empty(pc.shape[0], m.nrpt_latent, m.net_geoenc.out_nrchs, device=pc.device)
        # Capture needs a warmed-up process (library loaded, device tables created, kernel attributes set, every
        # kernel's module resident). warmup=0 means "do not advance training": one warm-up step still runs, on a
        # side stream, and parameters, moments and optimizer scalars are restored afterwards.
        snapshot = None
        if warmup <= 0:
            snapshot = [t.clone() for t in (self.flat.flat_p, self.exp_avg, self.exp_avg_sq, self.opt_state)]
            warmup = 1
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                if self.static_eps is not None:
                    self.static_eps.normal_()
                self.forward_backward(self.static_pc, self.static_nor, self.static_eps,
                                      early=self.world == 1 or self._symm_op is not None)
                self._finish_reduction()
                self.optimizer_step()
        torch.cuda.current_stream().wait_stream(side)
        if snapshot is not None:
            with torch.no_grad():
                for t, v in zip((self.flat.flat_p, self.exp_avg, self.exp_avg_sq, self.opt_state), snapshot):
                    t.copy_(v)
        torch.cuda.synchronize()
        l0 = _lib.lib().vcm_launch_count()
        # world_size 1: one graph for the whole step. world_size > 1: graph(fwd + bwd) -> NCCL all-reduce of the flat
        # gradient buffer -> graph(Adam). VCM_GRAPH_NCCL=1 captures the all-reduces as well (decoder-side range on
        # the early stream, the rest after the backward pass; one graph per step) -- opt-in: with c10d's NCCL
        # process group that capture hung on 2 x B200 in this image (torch 2.11 / NCCL 2.28), so it is not the default.
        nccl_in_graph = self.world == 1 or os.environ.get("VCM_GRAPH_NCCL", "0") == "1" or self._symm_op is not None
        g_fb = torch.cuda.CUDAGraph()
        cap_stream = torch.cuda.Stream(self.flat.flat_p.device, priority=self._prio)
        with torch.cuda.graph(g_fb, stream=cap_stream, capture_error_mode="relaxed"):
            self.static_loss, self.static_parts = self.forward_backward(self.static_pc, self.static_nor,
                                                                        self.static_eps, early=nccl_in_graph)
            if nccl_in_graph:
                self._finish_reduction()
                self.optimizer_step()
        g_opt = None
        if not nccl_in_graph:
            g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_opt, capture_error_mode="relaxed"):
                self.optimizer_step()
        self._graph = (g_fb, g_opt)
        self.graph_launches = int(_lib.lib().vcm_launch_count() - l0)     # our kernels inside one replayed step
        return self

    def _graph_step(self, pc, nor, eps=None):
        """One replay of the captured step. `eps` fixes the N(0,1) draw of the reparameterisation (tests and the
        multi-GPU equivalence check compare runs on the same sample); default: a fresh draw from torch's generator."""
        if pc.data_ptr() != self.static_pc.data_ptr():
            self.static_pc.copy_(pc, non_blocking=True)
        if nor.data_ptr() != self.static_nor.data_ptr():
            self.static_nor.copy_(nor, non_blocking=True)
        g_fb, g_opt = self._graph
        if self.static_eps is not None:
            if eps is not None:
                self.static_eps.copy_(eps, non_blocking=True)
            else:
                self.static_eps.normal_()
        g_fb.replay()
        if g_opt is not None:
            dist.all_reduce(self.flat.flat_g, op=dist.ReduceOp.SUM, group=self.group)
            g_opt.replay()
        return self.static_loss, self.static_parts

    # -- checkpoint layout of the reference (graphVAE_train.py:339-347) ----------
    def _sharded_moments(self):
        return self.world > 1 and self._symm_op == "fused"

    def _full_moments(self):
        """The Adam moments over the whole flat buffer. On the fused multi-GPU path each rank keeps them for the
        chunks it owns only (zero elsewhere): summing over the ranks assembles them. Collective in that case."""
        if not self._sharded_moments():
            return self.exp_avg, self.exp_avg_sq
        m, v = self.exp_avg.clone(), self.exp_avg_sq.clone()
        dist.all_reduce(m, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.group)
        return m, v

    def _keep_owned_moments(self):
        """After loading full moments on the fused path: zero what this rank does not own (chunk c of the flat buffer
        belongs to rank c % world, include/vcmesh.h vcm_reduce_adam_bcast), so that _full_moments stays a plain sum."""
        if not self._sharded_moments():
            return
        chunk = _lib.lib().vcm_reduce_adam_chunk()
        own = (torch.arange(self.flat.total, device=self.exp_avg.device) // chunk) % self.world == self.rank
        self.exp_avg.mul_(own)
        self.exp_avg_sq.mul_(own)

    def optimizer_state_dict(self):
        """The Adam state exactly as `torch.optim.Adam.state_dict()` lays it out for the reference's optimizer
        (graphVAE_train.py:252-262, 280, 345): `state[i]` = {step, exp_avg, exp_avg_sq} of the i-th parameter in
        getoptim order (decoder, encoder, decoder coefficients, encoder coefficients), one param group. The flat
        moment buffers are sliced per parameter, so the reference's `optimizer.load_state_dict(...)` (:291) accepts
        the file and resumes with the right bias correction."""
        step = float(self.step_count)
        exp_avg, exp_avg_sq = self._full_moments()
        by_id = {id(p): (o, p) for p, o in zip(self.flat.params, self.flat.offsets)}
        state, order = {}, self.model.trainable_parameters()
        for i, p in enumerate(order):
            o, _ = by_id[id(p)]
            n = p.numel()
            state[i] = {"step": torch.tensor(step, dtype=torch.float32),
                        "exp_avg": exp_avg[o:o + n].view(p.shape).clone(),
                        "exp_avg_sq": exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {"lr": float(self.lr), "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                 "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(len(order)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, opt):
        """Accepts a `torch.optim.Adam` state_dict over the getoptim parameter order (what the reference saves and
        what `optimizer_state_dict` writes) or the round-1 flat layout {step, lr, exp_avg, exp_avg_sq}. Raises when
        the state does not fit the model instead of silently restarting the moments."""
        if not opt:
            return
        with torch.no_grad():
            if "param_groups" in opt:
                order = self.model.trainable_parameters()
                ids = [i for g in opt["param_groups"] for i in g["params"]]
                if len(ids) != len(order):
                    raise RuntimeError("optimizer state covers %d parameters, the model has %d" % (len(ids), len(order)))
                by_id = {id(p): o for p, o in zip(self.flat.params, self.flat.offsets)}
                steps = set()
                for i, p in zip(ids, order):
                    st = opt["state"].get(i)
                    if st is None:                       # parameter that never received a gradient
                        continue
                    if tuple(st["exp_avg"].shape) != tuple(p.shape):
                        raise RuntimeError("optimizer state %d has shape %s, parameter has %s" %
                                           (i, tuple(st["exp_avg"].shape), tuple(p.shape)))
                    o, n = by_id[id(p)], p.numel()
                    self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
                    self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.add(float(st["step"]))
                if len(steps) > 1:
                    raise RuntimeError("per-parameter Adam step counts differ (%s): not a getoptim checkpoint" % sorted(steps))
                self.opt_state[0] = steps.pop() if steps else 0.0
                self.set_lr(float(opt["param_groups"][0]["lr"]))
            elif "exp_avg" in opt:
                if opt["exp_avg"].numel() != self.exp_avg.numel():
                    raise RuntimeError("flat optimizer state has %d elements, expected %d" %
                                       (opt["exp_avg"].numel(), self.exp_avg.numel()))
                self.exp_avg.copy_(opt["exp_avg"])
                self.exp_avg_sq.copy_(opt["exp_avg_sq"])
                self.opt_state[0] = float(opt.get("step", 0))
                self.set_lr(float(opt.get("lr", self.lr)))
            else:
                raise RuntimeError("unknown optimizer_state_dict layout: keys %s" % sorted(opt))
            self._keep_owned_moments()

    def state_dict(self):
        m = self.model
        return {"encgeo_state_dict": m.net_geoenc.state_dict(), "decgeo_state_dict": m.net_geodec.state_dict(),
                "mcvcoeffsdec_dict": m.mcvcoeffsdec.state_dict(), "mcvcoeffsenc_dict": m.mcvcoeffsenc.state_dict(),
                "optimizer_state_dict": self.optimizer_state_dict()}

    def save(self, folder, iteration):
        """Writes `model_%07d_best.weight` exactly as the reference does (graphVAE_train.py:341-347), optimizer
        state included in torch.optim.Adam's own layout."""
        path = os.path.join(folder, "model_%07d" % iteration + "_best.weight")
        sd = self.state_dict()                       # every rank takes part (the fused path shards the Adam moments)
        if self.rank == 0:
            os.makedirs(folder, exist_ok=True)
            torch.save(sd, path)
        return path

    def load(self, path):
        """Loads a checkpoint written by this trainer or by the reference (graphVAE_train.py:284-291): the four
        module state dicts and the Adam state (moments, step count, learning rate)."""
        ckpt = torch.load(path, map_location=self.flat.flat_p.device, weights_only=False)
        self.load_model_state(ckpt)
        self.load_optimizer_state_dict(ckpt.get("optimizer_state_dict"))
        return ckpt

    def fit(self, param, epochs_of_batches, log=None, gamma=0.99):
        """The reference's outer loop (graphVAE_train.py:303-353) over an iterable of epochs, each an
        iterable of (pc, nor) device batches: iteration counter from `start_iter` to `end_iter`,
        checkpoint every `evaluate_iter` iterations into `write_weight_folder` (rank 0 only), and the
        ExponentialLR(gamma) decay once per epoch. `log(iteration, loss, parts)` replaces its prints."""
        iteration = int(getattr(param, "start_iter", 0))
        end_iter = int(getattr(param, "end_iter", 1 << 62))
        every = int(getattr(param, "evaluate_iter", 0))
        for batches in epochs_of_batches:
            for pc, nor in batches:
                loss, parts = self.step(pc, nor)
                if log is not None:
                    log(iteration, loss, parts)
                if every and iteration % every == 0:
                    self.save(param.write_weight_folder, iteration)      # collective; rank 0 writes the file
                iteration += 1
                if iteration > end_iter:
                    break
            self.set_lr(self.lr * gamma)
            if iteration > end_iter:
                break
        return iteration

    def load_model_state(self, ckpt):
        m = self.model
        with torch.no_grad():
            for mod, key in ((m.net_geoenc, "encgeo_state_dict"), (m.net_geodec, "decgeo_state_dict"),
                             (m.mcvcoeffsdec, "mcvcoeffsdec_dict"), (m.mcvcoeffsenc, "mcvcoeffsenc_dict")):
                own = mod.state_dict()
                for k, v in ckpt[key].items():
                    own[k].copy_(v)              # in place: parameters stay views of the flat buffer


def init_distributed():
    """torchrun environment -> (rank, local_rank, world). NCCL over NVLink/NVSwitch."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local) if torch.cuda.is_available() else None)
    return rank, local, world


def build_synthetic_autoencoder(n_vert, n_levels=7, seed=0, residual_rate=0.9, template="sphere"):
    """Synthetic template + tables + model as BASELINE config C2/C3 describes:
    encoder tables pool1..pool6, decoder unpool6..unpool1 of the alternating
    stride-1/2 hierarchy, 6+6 layers with the trainer's hard-coded channels."""
    from . import graph_sampling, synthetic
    pts, faces = getattr(synthetic, template)(n_vert, seed=seed) if template != "torus" else synthetic.torus()
    tabs, nums = graph_sampling.build_tables(faces, len(pts), synthetic.alternating_levels(n_levels))
    enc = [tabs["pool%d" % l] for l in range(1, 7)]
    dec = [tabs["unpool%d" % l] for l in range(6, 0, -1)]
    param = make_param(enc, dec, tabs["pool0"], len(pts), residual_rate)
    torch.manual_seed(seed)
    model = Net_autoenc(param, faces)
    return model, pts, faces, tabs, nums
