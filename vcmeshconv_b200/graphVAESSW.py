"""Drop-in module layer: same classes, constructor signatures, attribute names
and state_dict layout as the reference's GraphAutoEncoder/graphVAESSW.py, with
every forward/backward running on libvcmesh.so (sm_100a CUDA kernels).

    import vcmeshconv_b200.graphVAESSW as vae_model      # instead of `import graphVAESSW`

Checkpoint compatibility (reference graphVAE_train.py:339-347): parameters
`weights [M, O*I]`, `bias [P_out, O] | [O]`, `p_neighbors [P_out, N]`,
`weight_res [1, O*I]` and the persistent buffers `neighbor_num_lst` (float32
[P_out]) and `neighbor_id_lstlst` (int64 [P_out, N]) keep their names, shapes
and dtypes, so reference checkpoints load with load_state_dict and vice versa.
The device-side tables (int32 ids, degree buckets, reverse CSR) are owned by a
vcm_tables handle and are not part of the state_dict.

Parameters are drawn from the CPU default generator in the same order and
shapes as the reference's initialisers, so torch.manual_seed(s) gives the same
initial weights as the reference under the same seed.
"""
import numpy as np
import torch
import torch.nn as nn

from . import functional as F_
from .connectivity import DeviceTables, avg_neighbor_num, split_table


class LASMConvssw(nn.Module):
    """vc-conv (square table), vd-pool (P_out < P_in) or vd-unpool (P_out > P_in)
    with optional residual branch. Mirrors graphVAESSW.py:29-163."""

    def __init__(self, in_channel, out_channel, weight_num, in_point_num, connection_info, b_Perpt_bias=True,
                 residual_rate=0.0):
        super().__init__()
        self.relu = nn.ELU()
        self.in_channel, self.out_channel, self.weight_num = in_channel, out_channel, weight_num
        self.in_point_num = in_point_num
        table = np.asarray(connection_info)
        counts, ids = split_table(table)
        self.out_point_num = table.shape[0]
        self.max_neighbor_num = ids.shape[1]
        self.avg_neighbor_num = avg_neighbor_num(table)
        self.residual_rate = residual_rate
        self.register_buffer("neighbor_num_lst", torch.from_numpy(counts))
        self.register_buffer("neighbor_id_lstlst", torch.from_numpy(ids))

        self.weights = nn.Parameter(torch.randn(weight_num, out_channel * in_channel))
        self.bias = nn.Parameter(torch.zeros(self.out_point_num, out_channel) if b_Perpt_bias
                                 else torch.zeros(out_channel))
        if residual_rate > 0:
            if self.out_point_num != in_point_num:
                self.p_neighbors = nn.Parameter(torch.randn(self.out_point_num, self.max_neighbor_num) /
                                                self.avg_neighbor_num)
            if out_channel != in_channel:
                self.weight_res = nn.Parameter(torch.randn(1, out_channel * in_channel) / out_channel)

        self._table_host = np.ascontiguousarray(table, dtype=np.int32)
        self._tables = {}          # device index -> DeviceTables (not in the state_dict)
        self._tables_x = {}        # same, for the basis-expanded table of the project-first order
        # The op is bilinear, so the basis projection may run before the gather:
        #   y[b,p,:] = sum_{n,m} A[p,n,m] (W_m x[b,idx[p,n],:])
        # which moves rows of M*O floats instead of materialising z (M*I per output vertex). It wins when
        # the projected rows are narrow (last decoder layer: 32 -> 3 channels on the full-resolution mesh).
        self.project_first = weight_num * out_channel <= 2 * in_channel

    # -- device tables ---------------------------------------------------------
    def tables(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("LASMConvssw runs on CUDA only (no CPU fallback); got tensor on %s" % device)
        key = device.index if device.index is not None else torch.cuda.current_device()
        t = self._tables.get(key)
        if t is None:
            t = self._tables[key] = DeviceTables(self._table_host, self.in_point_num, torch.device("cuda", key))
        return t

    def tables_expanded(self, device):
        """Device tables of the basis-expanded graph: virtual neighbour (n, m) of p is row idx[p,n]*M + m of
        the projected input viewed as [B, P_in*M, O]; the coefficient tensor A[P_out, N, M] viewed as
        [P_out, N*M, 1] lines up with it slot for slot."""
        key = device.index if device.index is not None else torch.cuda.current_device()
        t = self._tables_x.get(key)
        if t is None:
            M, p_in = self.weight_num, self.in_point_num
            ids = self._table_host[:, 1::2].astype(np.int64)                     # [P_out, N]
            real = ids != p_in
            ex = np.where(real[:, :, None], ids[:, :, None] * M + np.arange(M)[None, None, :], p_in * M)
            tab = np.full((ids.shape[0], 1 + 2 * ids.shape[1] * M), -1, np.int32)
            tab[:, 0] = real.sum(1) * M
            tab[:, 1::2] = ex.reshape(ids.shape[0], -1)
            t = self._tables_x[key] = DeviceTables(tab, p_in * M, torch.device("cuda", key))
        return t

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        ids = self.neighbor_id_lstlst.detach().cpu().numpy()
        if not np.array_equal(ids, self._table_host[:, 1::2]):
            # a checkpoint may carry different connectivity: rebuild the host table from its buffers
            cnt = self.neighbor_num_lst.detach().cpu().numpy().astype(np.int32)
            t = np.full((ids.shape[0], 1 + 2 * ids.shape[1]), -1, np.int32)
            t[:, 0] = cnt
            t[:, 1::2] = ids
            self._table_host = t
            self._tables, self._tables_x = {}, {}

    # -- forward ------------------------------------------------------------------
    def forward(self, in_pc, raw_w_weights, is_final_layer=False, b_max_pool=False):
        if b_max_pool:
            # the reference's max-pool branch raises TypeError (tuple + tensor, graphVAESSW.py:130-135)
            raise NotImplementedError("b_max_pool is not supported (it is broken in the reference as well)")
        t = self.tables(in_pc.device)
        if self.project_first:
            return self._forward_project_first(in_pc, raw_w_weights, is_final_layer)
        M, I, O = self.weight_num, self.in_channel, self.out_channel
        kind = F_.fused_conv_kind(t, in_pc.shape[0], M, I, O) if in_pc.is_cuda and in_pc.dim() == 3 else 0
        if kind == 2 or (kind == 1 and F_.fused_conv_enabled):
            # K1 + K2 in one launch: z stays on chip (it is stored once only when a backward pass follows)
            res, sy, sr = self._residual_join(self._residual_fork(in_pc, t))
            return F_.fused_conv(in_pc, raw_w_weights, self.weights, self.bias, res, t, M, I, O, not is_final_layer,
                                 sy, sr)
        branch = self._residual_fork(in_pc, t)
        z = F_.gather_contract(in_pc, raw_w_weights, t)
        return self.forward_from_z(z, in_pc, is_final_layer, branch)

    def _residual(self, in_pc, t):
        """graphVAESSW.py:144-159 -> (res or None, scale_y, scale_res)."""
        if self.residual_rate == 0:
            return None, 1.0, 0.0
        sy, sr = F_.blend_scales(self.residual_rate)
        I, O = self.in_channel, self.out_channel
        project = (lambda v: F_.basis_gemm(v, self.weight_res, None, None, 1, I, O, False)) if I != O else (lambda v: v)
        if self.in_point_num == self.out_point_num:
            return project(in_pc), sy, sr
        pn = F_.edge_weights(self.p_neighbors, t).unsqueeze(-1)
        # projection and weighted pooling are both linear and commute: gather on the narrower side
        # (the reference projects first, graphVAESSW.py:144-159; same result up to fp32 rounding)
        if I < O:
            return project(F_.gather_contract(in_pc, pn, t)), sy, sr
        return F_.gather_contract(project(in_pc), pn, t), sy, sr

    def _residual_fork(self, in_pc, t):
        """Runs the residual branch on the trainer's second stream when there is one: it only meets the main chain
        again in the epilogue (join in `_residual_join`). Returns ((res, sy, sr), stream or None)."""
        side = F_.residual_stream()
        if side is None or self.residual_rate == 0 or not in_pc.is_cuda:
            return self._residual(in_pc, t), None
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            r = self._residual(in_pc, t)
        return r, side

    @staticmethod
    def _residual_join(branch):
        r, side = branch
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        return r

    def _forward_project_first(self, in_pc, raw_w_weights, is_final_layer):
        B = in_pc.shape[0]
        M, I, O = self.weight_num, self.in_channel, self.out_channel
        branch = self._residual_fork(in_pc, self.tables(in_pc.device))
        # u[b,q,m,:] = W_m x[b,q,:]: one GEMM with the bases stacked along the output axis ([1][M*O][I] view)
        u = F_.basis_gemm(in_pc, self.weights.view(1, M * O * I), None, None, 1, I, M * O, False)
        t = self.tables(in_pc.device)
        if M * O <= 64:
            s = F_.project_gather(u, raw_w_weights, t, M, O)
        else:                                     # generic: an M = 1 gather over the basis-expanded table
            tx = self.tables_expanded(in_pc.device)
            s = F_.gather_contract(u.view(B, self.in_point_num * M, O),
                                   raw_w_weights.reshape(self.out_point_num, self.max_neighbor_num * M, 1), tx)
        res, sy, sr = self._residual_join(branch)
        return F_.bias_act(s, self.bias, res, not is_final_layer, sy, sr)

    def forward_from_z(self, z, in_pc, is_final_layer=False, branch=None):
        """Stage 2 on an already contracted z (lets two layers that share input,
        table and coefficients — the encoder's mu / std heads — share stage 1)."""
        t = self.tables(in_pc.device)
        M, I, O = self.weight_num, self.in_channel, self.out_channel
        act = not is_final_layer
        res, sy, sr = self._residual_join(branch if branch is not None else self._residual_fork(in_pc, t))
        return F_.basis_gemm(z, self.weights, self.bias, res, M, I, O, act, sy, sr)

    # the reference keeps its original, memory-hungry formulation as `forward2`
    # (graphVAESSW.py:167-233); per-edge kernels are never materialised here, both names run the same kernels.
    forward2 = forward


def _build_layers(structure, channel_lst, weight_num):
    layers = nn.ModuleList()
    for l in range(len(structure.connection_info_lsts)):
        layers.append(LASMConvssw(channel_lst[l], channel_lst[l + 1], weight_num, structure.ptnum_list[l],
                                  structure.connection_info_lsts[l], structure.perpoint_bias,
                                  structure.residual_rate))
    return layers


class MCEnc(nn.Module):
    """graphVAESSW.py:237-288. forward -> (mu, log-std)."""

    def __init__(self, structure, channel_lst, weight_num):
        super().__init__()
        self.point_num = structure.point_num
        self.residual_rate = structure.residual_rate
        self.b_max_pool = structure.b_max_pool
        self.perpoint_bias = structure.perpoint_bias
        self.channel_lst = channel_lst
        self.layer_num = len(structure.connection_info_lsts)
        self.layer_lst = _build_layers(structure, channel_lst, weight_num)
        last = self.layer_num - 1
        # second head on the last table (same draw order as the reference: created right after the last layer)
        self.stdlayer = LASMConvssw(channel_lst[last], channel_lst[last + 1], weight_num, structure.ptnum_list[last],
                                    structure.connection_info_lsts[last], structure.perpoint_bias,
                                    structure.residual_rate)
        self.out_nrpts = structure.ptnum_list[self.layer_num]
        self.out_nrchs = channel_lst[self.layer_num]
        # {layer index: tensor hook on that layer's input}: set by the trainer, which starts reducing / applying the
        # gradients of the layers behind a point of the backward pass as soon as it has passed that point
        self.activation_hooks = {}

    def forward_till_layer_n(self, in_pc, vcoeffs, layer_n):
        out_pc = in_pc
        for i in range(layer_n):
            hook = self.activation_hooks.get(i)
            if hook is not None and out_pc.requires_grad:
                out_pc.register_hook(hook)       # fires when d loss / d (input of layer i) is complete
            out_pc = self.layer_lst[i](out_pc, vcoeffs.vcoeffs_list[i], is_final_layer=False,
                                       b_max_pool=self.b_max_pool)
        return out_pc

    MU_SCALE, LOGSTD_SCALE = 0.1, 0.01           # graphVAESSW.py:286-288

    def forward_raw(self, in_pc, vcoeffs):
        """The two heads before their scaling (the fused reparameterisation kernel applies MU_SCALE / LOGSTD_SCALE)."""
        last = self.layer_num - 1
        h = self.forward_till_layer_n(in_pc, vcoeffs, last)
        if self.b_max_pool:
            raise NotImplementedError("b_max_pool is not supported")
        hook = self.activation_hooks.get(last)
        if hook is not None and h.requires_grad:
            h.register_hook(hook)
        head = self.layer_lst[last]
        # mu and std heads share input, table and coefficients: one gather/contract, two basis GEMMs
        z = F_.gather_contract(h, vcoeffs.vcoeffs_list[last], head.tables(h.device))
        side = F_.head_stream() if h.is_cuda else None
        if side is None:
            return (head.forward_from_z(z, h, is_final_layer=True),
                    self.stdlayer.forward_from_z(z, h, is_final_layer=True))
        # the two heads are independent from here on (51 vertices: a few CTAs per kernel): run them side by side;
        # autograd replays each head's backward on the stream of its forward
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        # The std head keeps its residual projection on its own stream: a residual tensor made on the shared residual
        # stream dies as soon as this head's GEMM is enqueued, and the mu head's residual - ordered after `cur` only -
        # would be handed the same memory while that GEMM may still be reading it.
        res_stream = F_.residual_stream()
        F_.set_residual_stream(None)
        try:
            with torch.cuda.stream(side):
                std = self.stdlayer.forward_from_z(z, h, is_final_layer=True)
        finally:
            F_.set_residual_stream(res_stream)
        mu = head.forward_from_z(z, h, is_final_layer=True)
        cur.wait_stream(side)
        return mu, std

    def forward(self, in_pc, vcoeffs):
        mu, std = self.forward_raw(in_pc, vcoeffs)
        return mu * self.MU_SCALE, std * self.LOGSTD_SCALE


class MCDec(nn.Module):
    """graphVAESSW.py:296-343. Last layer is linear."""

    def __init__(self, structure, channel_lst, weight_num):
        super().__init__()
        self.point_num = structure.point_num
        self.residual_rate = structure.residual_rate
        self.b_max_pool = structure.b_max_pool
        self.perpoint_bias = structure.perpoint_bias
        self.channel_lst = channel_lst
        self.layer_num = len(structure.connection_info_lsts)
        self.layer_lst = _build_layers(structure, channel_lst, weight_num)

    def forward(self, latent, vcoeffs):
        return self.forward_from_layer_n(latent, vcoeffs, 0)

    forward_dec = forward

    def forward_from_layer_n(self, in_pc, vcoeffs, layer_n):
        out_pc = in_pc
        for i in range(layer_n, self.layer_num):
            out_pc = self.layer_lst[i](out_pc, vcoeffs.vcoeffs_list[i], is_final_layer=(i == self.layer_num - 1),
                                       b_max_pool=self.b_max_pool)
        return out_pc


class MCStructure(nn.Module):
    """Loads the per-layer connectivity tables (graphVAESSW.py:346-374). Entries
    of `connection_layer_fn_lst_*` may be paths of `.npy` files (as in the
    reference) or the int tables themselves."""

    def __init__(self, param, inptnr, weight_num, bDec=True, b_perpoint_bias=True):
        super().__init__()
        self.point_num = inptnr
        self.residual_rate = param.residual_rate
        self.b_max_pool = param.conv_max
        self.perpoint_bias = b_perpoint_bias
        self.connection_layer_fn_lst = param.connection_layer_fn_lst_dec if bDec else param.connection_layer_fn_lst_enc
        self.layer_num = len(self.connection_layer_fn_lst)
        self.ptnum_list = [inptnr]
        self.connection_info_lsts = []
        for fn in self.connection_layer_fn_lst:
            info = np.load(fn) if isinstance(fn, (str, bytes)) or hasattr(fn, "__fspath__") else np.asarray(fn)
            if info.ndim != 2 or (info.shape[1] - 1) % 2:
                raise ValueError("connection table %r must be [P_out, 1+2N]" % (fn,))
            self.connection_info_lsts.append(info)
            self.ptnum_list.append(info.shape[0])

    def forward(self):
        return


class MCVcoeffs(nn.Module):
    """Per-vertex, per-neighbour coefficients A_l [P_out, N, M], one per layer
    (graphVAESSW.py:377-400); init randn / (avg_neighbours * M)."""

    def __init__(self, structure, weight_num):
        super().__init__()
        self.layer_num = len(structure.connection_layer_fn_lst)
        self.vcoeffs_list = nn.ParameterList()
        for info in structure.connection_info_lsts:
            n_max = (info.shape[1] - 1) // 2
            self.vcoeffs_list.append(nn.Parameter(torch.randn(info.shape[0], n_max, weight_num) /
                                                  (avg_neighbor_num(info) * weight_num)))


class MCLoss(nn.Module):
    """Geometric and graph-Laplacian losses over the layer-0 table
    (graphVAESSW.py:405-575).

    The reference builds the Laplacian with a Python loop of N advanced-index
    gathers (whose autograd backward is a sort-based index_put). Here it is ONE
    launch of the gather/contract kernel (K1) with M = 1 and constant
    coefficients  L[p,0] = cnt_p, L[p,n>0] = -1  over the same layer-0 table;
    its backward is the deterministic reverse-CSR scatter (K6). The Laplacian is
    linear, so the losses apply it to (gt - prediction) once instead of twice."""

    def __init__(self, param):
        super().__init__()
        ids = np.asarray(param.neighbor_id_lstlst)
        cnt = np.asarray(param.neighbor_num_lst)
        self.register_buffer("initial_neighbor_id_lstlst", torch.as_tensor(ids).long())
        self.register_buffer("initial_neighbor_num_lst", torch.as_tensor(cnt).float())
        self.initial_max_neighbor_num = self.initial_neighbor_id_lstlst.shape[1]
        n_pt, n_max = ids.shape
        coef = -np.ones((n_pt, n_max, 1), np.float32)
        coef[:, 0, 0] = cnt.astype(np.float32)
        self.register_buffer("_lap_coeffs", torch.from_numpy(coef), persistent=False)
        table = np.full((n_pt, 1 + 2 * n_max), -1, np.int32)
        table[:, 0] = cnt
        table[:, 1::2] = ids
        self._table_host = table
        self._tables = {}

    def forward(self):
        return

    def tables(self, dev):
        """Device tables of the layer-0 (Laplacian) connectivity."""
        if dev.type != "cuda":
            raise RuntimeError("MCLoss runs on CUDA only (no CPU fallback)")
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        t = self._tables.get(key)
        if t is None:
            t = self._tables[key] = DeviceTables(self._table_host, self._table_host.shape[0], torch.device("cuda", key))
        return t

    def _laplacian(self, pc):
        return F_.gather_contract(pc, self._lap_coeffs, self.tables(pc.device))          # [B, P, 3]

    def compute_geometric_loss_l1(self, gt_pc, predict_pc, weights=[]):
        if len(weights) == 0:
            return torch.abs(gt_pc - predict_pc).mean()
        return ((gt_pc - predict_pc).abs() * weights).sum() / (weights.sum() + 1e-6)

    def compute_geometric_loss_l2(self, gt_pc, predict_pc, weights=[]):
        if len(weights) == 0:
            return (gt_pc - predict_pc).pow(2).sum(2).mean()
        return ((gt_pc - predict_pc).pow(2).sum(2, keepdim=True) * weights).sum() / (weights.sum() + 1e-6)

    def compute_geometric_mean_euclidean_dist_error(self, gt_pc, predict_pc, weights=[]):
        d = (gt_pc - predict_pc).pow(2).sum(2).pow(0.5)
        return d.mean() if len(weights) == 0 else (d * weights).sum()

    def compute_laplace_loss_l1(self, gt_pc_raw, predict_pc_raw, weights=[]):
        d = self._laplacian(gt_pc_raw - predict_pc_raw).abs()
        return d.mean() if len(weights) == 0 else (d * weights).sum() / (weights.sum() + 1e-6)

    def compute_laplace_loss_l2(self, gt_pc_raw, predict_pc_raw, weights=[]):
        d = self._laplacian(gt_pc_raw - predict_pc_raw).pow(2)
        return d.sum(2).mean() if len(weights) == 0 else (d * weights).sum() / (weights.sum() + 1e-6)

    def compute_laplace_Mean_Euclidean_Error(self, gt_pc_raw, predict_pc_raw):
        return self._laplacian(gt_pc_raw - predict_pc_raw).pow(2).sum(2).pow(0.5).mean()
