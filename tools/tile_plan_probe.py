#!/usr/bin/env python
"""Feasibility probe for the tcgen05 gather/contract kernel of DESIGN.md section 7 (CPU only): group the output
vertices of every C2 layer into tiles of at most T vertices (T*17 accumulator rows <= 128) whose UNION of neighbour
ids has at most KMAX entries (the K extent of the dense local coefficient matrix), greedily by shared neighbours, and
report tile fill and K. In table / processing order consecutive vertices share almost no neighbours (7 vertices ->
7 x the per-vertex count), so the kernel needs such a spatial tile plan next to the degree-sorted order.

    python tools/tile_plan_probe.py [T] [KMAX]
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench


def greedy(table, nums, T, KMAX):
    P, ids = table.shape[0], table[:, 1::2]
    mx = int(ids.max())
    p_in = mx if mx in nums else mx + 1
    nb = [[int(q) for q in ids[p] if q != p_in] for p in range(P)]
    rev = collections.defaultdict(list)
    for p in range(P):
        for q in nb[p]:
            rev[q].append(p)
    done, tiles = np.zeros(P, bool), []
    for p0 in range(P):
        if done[p0]:
            continue
        tile, u = [p0], set(nb[p0])
        done[p0] = True
        while len(tile) < T:
            cand = collections.Counter()
            for q in u:
                for p in rev[q]:
                    if not done[p]:
                        cand[p] += 1
            if not cand:
                break
            best = min(cand, key=lambda p: (len(nb[p]) - cand[p], p))      # fewest NEW neighbours
            if len(u) + len(nb[best]) - cand[best] > KMAX:
                break
            tile.append(best)
            done[best] = True
            u.update(nb[best])
        tiles.append((len(tile), len(u)))
    return np.array(tiles), float(np.mean([len(x) for x in nb]))


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    KMAX = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    pts, faces, tabs, nums = bench.build_tables(6890)
    for name in ["pool1", "pool2", "pool3", "pool4", "pool5", "pool6", "unpool6", "unpool5", "unpool4", "unpool3",
                 "unpool2", "unpool1"]:
        a, avg = greedy(tabs[name], nums, T, KMAX)
        print("%-8s P_out %5d, %4.1f neighbours/vertex: %4d tiles, %.2f vertices/tile, distinct neighbours mean %5.1f "
              "max %3d, accumulator rows used %3.0f %%" % (name, tabs[name].shape[0], avg, len(a), a[:, 0].mean(),
                                                          a[:, 1].mean(), a[:, 1].max(),
                                                          100 * (a[:, 0] * 17).sum() / (len(a) * 128.0)))


if __name__ == "__main__":
    main()
