"""bench.py contract on a machine without a GPU: the reference arm (`--impl reference`) prints ONE JSON line with
the keys the driver reads, and the product arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--vertices", "600", "--batch", "2", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                  # one JSON line, nothing else on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("training meshes/sec") and d["unit"] == "meshes/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and d["dtype"] == "f32" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    from oracle import ref_autoencoder
    assert cb["kind"] == ("reference" if ref_autoencoder.available() else "port") and cb["cores"] >= 1 and cb["unit"] == "meshes/s" and cb["value"] == d["value"]
    assert "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a GPU")
def test_product_arm_refuses_to_run_without_cuda():
    r = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
