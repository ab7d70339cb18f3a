"""Shared helpers for the test-suite (tests may import oracle/, the product may not)."""
import glob
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def layer_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "layer_*.npz")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def rel_err(a, b):
    """max |a-b| / max |b|  (norm-wise relative error, the tolerance the
    north_star states: fp32 fwd/bwd <= 1e-5 relative to the fp64 gold)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = np.abs(b).max() if b.size else 0.0
    num = np.abs(a - b).max() if b.size else 0.0
    return num / den if den > 0 else num


def oracle_layer_from_golden(g, dtype=torch.float64):
    """(Layer, coeffs, params dict, x, gy) in `dtype` from a golden layer record."""
    from oracle import vcconv_oracle as orc
    layer = orc.Layer(g["table"], int(g["in_channel"]), int(g["out_channel"]), int(g["in_point_num"]),
                      float(g["residual_rate"]))
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)
    params = {k[len("param."):]: t(v) for k, v in g.items()
              if k.startswith("param.") and k not in ("param.neighbor_num_lst", "param.neighbor_id_lstlst")}
    return layer, t(g["coeffs"]), params, t(g["x"]), t(g["gy"])
