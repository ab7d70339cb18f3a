"""TEST INFRASTRUCTURE — runs the reference's own GraphSampling (oracle/_ref/
graphsampling_ref, built by oracle/Makefile from /root/reference) and loads the
.npy tables it writes. Only tests/ may use this; the product's tables come from
vcmeshconv_b200.graph_sampling."""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "_ref", "graphsampling_ref")


def available():
    return os.path.exists(BIN)


def build():
    if os.path.isdir("/root/reference/GraphSampling"):
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
    return available()


def run(obj_path, levels):
    """levels: [(stride, r_pool, r_unpool)] -> dict name -> int32 table."""
    with tempfile.TemporaryDirectory() as d:
        prefix = os.path.join(d, "t")
        cmd = [BIN, obj_path, prefix] + ["%d,%d,%d" % tuple(l) for l in levels]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference GraphSampling failed (%d): %s" % (r.returncode, r.stderr[-500:]))
        out = {}
        for l in range(len(levels)):
            out["pool%d" % l] = np.load(prefix + "_pool%d.npy" % l)
            out["unpool%d" % l] = np.load(prefix + "_unpool%d.npy" % l)
        return out
