"""Runs tests/multi_gpu_equiv.py (N-rank captured step over symmetric memory vs one process fed all shards) when the
box has at least two GPUs; on a one-GPU box it is skipped (the driver's `gpurun --gpus 2` run keeps its output in
profiles/multi_gpu_equiv_r2.json)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_two_ranks_equal_one_process_fed_both_shards():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_equiv.py"), "--steps", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["max_abs_param_diff"] <= 1e-6 and res["max_abs_param_spread_over_ranks"] == 0.0
