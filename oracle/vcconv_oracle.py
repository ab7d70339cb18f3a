"""TEST INFRASTRUCTURE — CPU oracle for the vc-conv / pool / unpool hot path.

This is a restatement, in plain functional PyTorch on the CPU, of the
reference's operator and of the thin glue around it. It is NOT on the product
path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg import it. The product (vcmeshconv_b200/) never does and
raises if its CUDA library is missing.

Pinned: tests/golden/*.npz were produced by oracle/gen_golden.py, which imports
the UNMODIFIED reference (/root/reference/GraphAutoEncoder/graphVAESSW.py) in
the build container and records its outputs and autograd gradients in fp64 and
fp32 for `LASMConvssw.forward`, `.forward2`, `MCEnc`, `MCDec` and `MCLoss`.
tests/test_oracle_golden.py checks every function below against those vectors
(the reference itself ships no tests, fixtures or golden vectors — SURVEY §4).

The arithmetic lives in PyTorch ATen (index_select / einsum->bmm / sum / elu);
the reference does not pin a torch version (requirements.txt:1-4). The vectors
were generated with torch 2.11.0+cu128 (CPU), numpy 2.3.5.

Every function works in whatever dtype its inputs have (fp32 as the reference
runs, fp64 for the high-precision gold the 1e-5 tolerance is measured against).
The op sequence deliberately follows the reference's (same gathers, the same
two einsum contractions, same order of bias / ELU / residual blend) so that
timing this file on host cores is a fair stand-in for timing the reference.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# connectivity tables
# --------------------------------------------------------------------------
def parse_table(table):
    """[P_out, 1+2N] int table -> (counts float32 [P_out], ids int64 [P_out, N]).

    Follows LASMConvssw.__init__ (graphVAESSW.py:44-51): column 0 is the
    neighbour count, the rest are (id, hop) pairs of which only the id is used.
    """
    table = np.asarray(table)
    p_out = table.shape[0]
    counts = torch.from_numpy(table[:, 0].astype(np.float32))
    ids = torch.from_numpy(np.ascontiguousarray(table[:, 1:].reshape(p_out, -1, 2)[:, :, 0])).long()
    return counts, ids


def avg_neighbor_num(table):
    """round(mean count), graphVAESSW.py:56 / :396 (Python round on a float)."""
    return round(torch.from_numpy(np.asarray(table)[:, 0].astype(np.float32)).mean().item())


def _gather_rows(x, ids, dim):
    """x.index_select(dim, ids.flat) reshaped so `dim` is replaced by ids.shape
    (the reference's index_selection_nd, graphVAESSW.py:23-27)."""
    shape = list(x.shape)
    shape[dim:dim + 1] = list(ids.shape)
    return x.index_select(dim, ids.reshape(-1)).reshape(shape)


# --------------------------------------------------------------------------
# the operator
# --------------------------------------------------------------------------
def vcconv_forward(x, coeffs, weights, bias, ids, in_point_num, *, is_final_layer=False,
                   residual_rate=0.0, p_neighbors=None, weight_res=None):
    """One vc-conv / vd-pool / vd-unpool layer; LASMConvssw.forward,
    graphVAESSW.py:97-163.

    x        [B, P_in, I]        coeffs  [P_out, N, M]   (MCVcoeffs entry)
    weights  [M, O*I]            bias    [P_out, O] or [O]
    ids      [P_out, N] int64, padding entries == in_point_num
    """
    batch, p_in, cin = x.shape
    assert p_in == in_point_num
    p_out, nmax = ids.shape
    m = weights.shape[0]
    cout = weights.shape[1] // cin

    # :111-113  sentinel mask obtained by gathering a ones-vector whose slot P_in is 0
    ones = torch.ones(p_in + 1, dtype=x.dtype, device=x.device)
    ones[p_in] = 0
    mask = _gather_rows(ones, ids, 0).contiguous()                       # [P_out, N]
    # :118
    a_eff = coeffs * mask.view(p_out, nmax, 1)
    # :121-122  zero row appended, then gathered
    xpad = torch.cat((x, torch.zeros(batch, 1, cin, dtype=x.dtype, device=x.device)), 1)  # [B, P_in+1, I]
    g = _gather_rows(xpad, ids, 1)                                       # [B, P_out, N, I]
    # :123  stage 1: neighbours -> M basis slots
    z = torch.einsum('pnm,bpni->bpmi', a_eff, g)                         # [B, P_out, M, I]
    # :125-133  stage 2: per-basis channel mixing, then sum over bases
    w = weights.view(m, cout, cin)
    u = torch.einsum('moi,bpmi->bpmo', w, z)                             # [B, P_out, M, O]
    y = u.sum(2)
    # :135-138
    y = y + bias
    if not is_final_layer:
        y = F.elu(y)
    if residual_rate == 0:
        return y

    # :144-145  1x1 projection of the (padded) input when channel counts differ
    if cin != cout:
        xpad = torch.einsum('oi,bpi->bpo', weight_res.view(cout, cin), xpad)
    if p_in == p_out:
        # :148-149 identity skip
        res = xpad[:, 0:p_in].clone()
    else:
        # :151-159 density-weighted pool/unpool with scalar per-edge weights
        gr = _gather_rows(xpad, ids, 1)                                  # [B, P_out, N, O]
        pn = torch.abs(p_neighbors) * mask
        pn_sum = pn.sum(1) + 1e-8
        pn = pn / pn_sum.view(p_out, 1).repeat(1, nmax)
        res = torch.einsum('pn,bpno->bpo', pn, gr)
    # :161  (np.sqrt of a python float: a float64 scalar, weak-typed against the tensor)
    return y * np.sqrt(1 - residual_rate) + res * np.sqrt(residual_rate)


def vcconv_forward_edge_kernels(x, coeffs, weights, bias, ids, in_point_num, *, is_final_layer=False):
    """Second, independent formulation (LASMConvssw.forward2, graphVAESSW.py:
    167-211, conv part only): materialise the per-edge kernels
    K[p,n] = sum_m A[p,n,m] W_m, then apply them. Used as a cross-check."""
    batch, p_in, cin = x.shape
    p_out, nmax = ids.shape
    cout = weights.shape[1] // cin
    ones = torch.ones(p_in + 1, dtype=x.dtype, device=x.device)
    ones[p_in] = 0
    mask = ones[ids].contiguous()
    a_eff = coeffs * mask.view(p_out, nmax, 1)
    k = torch.einsum('pnm,mc->pnc', a_eff, weights).view(p_out, nmax, cout, cin)   # :192-193
    xpad = torch.cat((x, torch.zeros(batch, 1, cin, dtype=x.dtype, device=x.device)), 1)
    g = xpad[:, ids]                                                               # :197
    y = torch.einsum('pnoi,bpni->bpno', k, g).sum(2) + bias                        # :199-207
    return y if is_final_layer else F.elu(y)


# --------------------------------------------------------------------------
# parameter containers (plain dicts of tensors keyed like the reference's
# state_dict entries, graphVAESSW.py:61-85 and :398)
# --------------------------------------------------------------------------
def init_layer_params(table, in_channel, out_channel, weight_num, in_point_num, perpoint_bias=True,
                      residual_rate=0.0, generator=None, dtype=torch.float32):
    table = np.asarray(table)
    p_out = table.shape[0]
    nmax = (table.shape[1] - 1) // 2
    avg = avg_neighbor_num(table)
    rn = lambda *s: torch.randn(*s, generator=generator, dtype=dtype)
    p = {"weights": rn(weight_num, out_channel * in_channel),
         "bias": torch.zeros(p_out, out_channel, dtype=dtype) if perpoint_bias else torch.zeros(out_channel, dtype=dtype)}
    if residual_rate > 0:
        if p_out != in_point_num:
            p["p_neighbors"] = rn(p_out, nmax) / avg
        if out_channel != in_channel:
            p["weight_res"] = rn(1, out_channel * in_channel) / out_channel
    return p


def init_coeffs(table, weight_num, generator=None, dtype=torch.float32):
    table = np.asarray(table)
    p_out = table.shape[0]
    nmax = (table.shape[1] - 1) // 2
    return torch.randn(p_out, nmax, weight_num, generator=generator, dtype=dtype) / (avg_neighbor_num(table) * weight_num)


class Layer:
    """Table + hyper-parameters of one layer (what LASMConvssw.__init__ keeps)."""

    def __init__(self, table, in_channel, out_channel, in_point_num, residual_rate=0.0):
        self.counts, self.ids = parse_table(table)
        self.in_channel, self.out_channel = in_channel, out_channel
        self.in_point_num, self.out_point_num = in_point_num, self.ids.shape[0]
        self.residual_rate = residual_rate

    def __call__(self, x, coeffs, params, is_final_layer=False):
        return vcconv_forward(x, coeffs, params["weights"], params["bias"], self.ids, self.in_point_num,
                              is_final_layer=is_final_layer, residual_rate=self.residual_rate,
                              p_neighbors=params.get("p_neighbors"), weight_res=params.get("weight_res"))


# --------------------------------------------------------------------------
# layer stacks
# --------------------------------------------------------------------------
def encoder_forward(x, layers, coeffs, params, std_params):
    """MCEnc.forward, graphVAESSW.py:276-288: L-1 activated layers, then two
    un-activated heads on the same input and the same table/coefficients,
    scaled 0.1 (mu) and 0.01 (log-std)."""
    h = x.clone()
    last = len(layers) - 1
    for l in range(last):
        h = layers[l](h, coeffs[l], params[l])
    mu = layers[last](h, coeffs[last], params[last], is_final_layer=True) * 0.1
    logstd = layers[last](h, coeffs[last], std_params, is_final_layer=True) * 0.01
    return mu, logstd


def decoder_forward(z, layers, coeffs, params):
    """MCDec.forward_from_layer_n, graphVAESSW.py:335-343: last layer is linear."""
    h = z.clone()
    for l in range(len(layers)):
        h = layers[l](h, coeffs[l], params[l], is_final_layer=(l == len(layers) - 1))
    return h


# --------------------------------------------------------------------------
# losses (MCLoss + Net_autoenc glue)
# --------------------------------------------------------------------------
def geometric_loss_l1(gt, pred):
    """graphVAESSW.py:416-421 (unweighted branch)."""
    return torch.abs(gt - pred).mean()


def _laplacian(pc, ids0, counts0):
    """cnt_p * x_self - sum of the other neighbours, over the layer-0 table
    whose first entry is the vertex itself (graphVAESSW.py:515-525)."""
    batch, n, _ = pc.shape
    pad = torch.cat((pc, torch.zeros(batch, 1, 3, dtype=pc.dtype, device=pc.device)), 1)
    lap = pad[:, ids0[:, 0]] * counts0.to(pc.dtype).view(1, n, 1).repeat(batch, 1, 3)
    for k in range(1, ids0.shape[1]):
        lap = lap - pad[:, ids0[:, k]]
    return lap


def laplace_loss_l2(gt, pred, ids0, counts0):
    """graphVAESSW.py:507-542 (unweighted branch)."""
    return (_laplacian(gt, ids0, counts0) - _laplacian(pred, ids0, counts0)).pow(2).sum(2).mean()


def autoencoder_losses(x, normals, faces, eps, enc, dec, ids0, counts0, w_pose=1.0, w_laplace=100.0,
                       klweight=1e-5, w_nor=10.0):
    """Net_autoenc.forward, graphVAE_train.py:199-225, with the N(0,1) draw
    passed in (`eps`) so two implementations can be compared on the same sample.

    enc = (layers, coeffs, params, std_params), dec = (layers, coeffs, params),
    faces int64 [F,3], normals [B,F,3]. Returns (loss, l1, laplace, kl, normal).
    """
    mu, logstd = encoder_forward(x, *enc)
    std = logstd.exp()
    z = mu + std * eps
    kl = torch.mean(-0.5 - logstd + 0.5 * mu ** 2 + 0.5 * torch.exp(2 * logstd))
    out = decoder_forward(z, *dec)[:, :, 0:3]
    dif = out - x
    v0 = _gather_rows(dif, faces[:, 0], 1)
    v1 = _gather_rows(dif, faces[:, 1], 1)
    v2 = _gather_rows(dif, faces[:, 2], 1)
    l_nor = ((v1 + v0 + v2) / 3.0 * normals).sum(2).pow(2).mean()
    l_l1 = geometric_loss_l1(x[:, :, 0:3], out)
    l_lap = laplace_loss_l2(x[:, :, 0:3], out, ids0, counts0)
    loss = l_l1 * w_pose + l_lap * w_laplace + kl * klweight + l_nor * w_nor
    return loss, l_l1, l_lap, kl, l_nor, out


# --------------------------------------------------------------------------
# whole autoencoder as the trainer builds it (graphVAE_train.py:136-178)
# --------------------------------------------------------------------------
ENC_CHANNELS = [3, 32, 64, 128, 256, 512, 64]     # graphVAE_train.py:151
DEC_CHANNELS = [64, 512, 256, 128, 64, 32, 3]     # graphVAE_train.py:172
WEIGHT_NUM = 17                                   # graphVAE_train.py:136


class AutoEncoder:
    """Tables + freshly initialised parameters for the 6+6 stack; `step` runs
    forward + backward (+ optional Adam) on the CPU."""

    def __init__(self, enc_tables, dec_tables, point_num, table0, faces, residual_rate=0.9, seed=0,
                 enc_channels=ENC_CHANNELS, dec_channels=DEC_CHANNELS, weight_num=WEIGHT_NUM, dtype=torch.float32):
        g = torch.Generator().manual_seed(seed)
        self.dtype = dtype

        def stack(tables, channels, n_in):
            layers, coeffs, params = [], [], []
            for l, t in enumerate(tables):
                layers.append(Layer(t, channels[l], channels[l + 1], n_in, residual_rate))
                coeffs.append(init_coeffs(t, weight_num, g, dtype))
                params.append(init_layer_params(t, channels[l], channels[l + 1], weight_num, n_in, True,
                                                residual_rate, g, dtype))
                n_in = layers[-1].out_point_num
            return layers, coeffs, params, n_in

        el, ec, ep, n_lat = stack(enc_tables, enc_channels, point_num)
        sp = init_layer_params(enc_tables[-1], enc_channels[-2], enc_channels[-1], weight_num,
                               el[-1].in_point_num, True, residual_rate, g, dtype)
        dl, dc, dp, _ = stack(dec_tables, dec_channels, n_lat)
        self.enc, self.dec = (el, ec, ep, sp), (dl, dc, dp)
        self.counts0, self.ids0 = parse_table(table0)
        self.faces = torch.as_tensor(np.asarray(faces)).long()
        self.latent_shape = (n_lat, enc_channels[-1])

    def parameters(self):
        el, ec, ep, sp = self.enc
        dl, dc, dp = self.dec
        out = list(ec) + list(dc) + list(sp.values())
        for d in ep + dp:
            out += list(d.values())
        return out

    def requires_grad_(self, flag=True):
        for p in self.parameters():
            p.requires_grad_(flag)
        return self

    def to(self, device):
        """Moves parameters and tables (bench.py's ATen-on-GPU baseline; everything else runs on the CPU)."""
        for p in self.parameters():
            p.data = p.data.to(device)
        for lay in self.enc[0] + self.dec[0]:
            lay.counts, lay.ids = lay.counts.to(device), lay.ids.to(device)
        self.counts0, self.ids0, self.faces = self.counts0.to(device), self.ids0.to(device), self.faces.to(device)
        return self

    def losses(self, x, normals, eps, **kw):
        return autoencoder_losses(x, normals, self.faces, eps, self.enc, self.dec, self.ids0, self.counts0, **kw)
