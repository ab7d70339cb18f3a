#!/usr/bin/env python
"""profiles/traffic_r2.json: measured DRAM bytes (ncu --set full, dram__bytes_read/write.sum) against the algorithmic
bytes of the same launch (tools/kernel_bench.py prints them) for the hot kernels on the dec3 layer shape
(unpool3: 344 -> 1811 vertices, I = 128, O = 64, M = 17, B = 16). bench.py multiplies the ratio of the dominant
kernel by its average algorithmic bytes per launch to fill `roofline.traffic`.

    python tools/make_traffic.py gpurun_out profiles/ncu_r2_plain_runs.txt > profiles/traffic_r2.json
"""
import json
import re
import subprocess
import sys
import csv
import io

d, plain = sys.argv[1], sys.argv[2]
alg = {}
for line in open(plain):
    m = re.match(r"(\w+) (\w+)\s+P.*bytes=(\d+)", line)
    if m:
        alg[m.group(2)] = int(m.group(3))
KEYS = {"k5": ("vcm_coeff_grad_ex tma", "k5"), "k1": ("vcm_gather_contract_fwd_ex", "k1"), "gemm_fwd": ("vcm_basis_fwd", "fwd"),
        "gemm_dz": ("vcm_basis_bwd_dz", "dz"), "gemm_dw": ("vcm_basis_bwd_dw", "dw"), "k6": ("vcm_scatter_rev tma", "k6")}
out = {"layer": "unpool3 (344->1811, I=128, O=64, M=17, B=16)", "kernels": {}}
for tag, (key, kb) in KEYS.items():
    rep = "%s/r2_ncu_%s.ncu-rep" % (d, tag)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, vals = rows[0], rows[2]
    g = lambda n: float(vals[hdr.index(n)].replace(",", ""))
    unit = lambda n: rows[1][hdr.index(n)]
    scale = lambda n: {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit(n)]
    rd = g("dram__bytes_read.sum") * scale("dram__bytes_read.sum")
    wr = g("dram__bytes_write.sum") * scale("dram__bytes_write.sum")
    out["kernels"][key] = {"kernel": vals[hdr.index("Kernel Name")], "dram_bytes_read": rd, "dram_bytes_write": wr,
                           "dram_bytes": rd + wr, "algorithmic_bytes": alg[kb], "ratio": (rd + wr) / alg[kb],
                           "duration_us": g("gpu__time_duration.sum"), "layer": out["layer"],
                           "source": "ncu --set full --clock-control none, profiles/ncu_r2_%s.txt" % tag}
json.dump(out, sys.stdout, indent=1)
