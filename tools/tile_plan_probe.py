#!/usr/bin/env python
"""Feasibility probe for the tcgen05 gather/contract kernel of DESIGN.md section 7 (CPU only): group the output
vertices of every C2 layer into tiles of at most T vertices (T*17 accumulator rows <= 128) whose UNION of neighbour
ids has at most KMAX entries (the K extent of the dense local coefficient matrix), greedily by shared neighbours, and
report tile fill and K. In table / processing order consecutive vertices share almost no neighbours (7 vertices ->
7 x the per-vertex count), so the kernel needs such a spatial tile plan next to the degree-sorted order.

    python tools/tile_plan_probe.py [T] [KMAX]
    python tools/tile_plan_probe.py check        numerical check of the dense-tile formulation on one layer
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench


plan = []          # [(output vertices of the tile, sorted distinct neighbour ids)] of the last greedy() call


def check(T=7, KMAX=64, M=17, I=8, B=3):
    """z[b,p,m,i] = sum_n A[p,n,m] x[b,idx[p,n],i] computed tile by tile as ONE dense product per tile,
    z_tile[(p,m), (b,i)] = S_tile[(p,m), k] @ X_tile[k, (b,i)] with S_tile the zero-padded local coefficient matrix
    (k = position of a neighbour in the tile's distinct-neighbour list), against the direct sum."""
    pts, faces, tabs, nums = bench.build_tables(6890)
    table = tabs["unpool2"]                                   # 1811 -> 1811 convolution layer
    greedy(table, nums, T, KMAX)
    P, ids = table.shape[0], table[:, 1::2]
    mx = int(ids.max())
    p_in = mx if mx in nums else mx + 1
    rng = np.random.default_rng(0)
    A = rng.standard_normal((P, ids.shape[1], M))
    x = rng.standard_normal((B, p_in, I))
    xpad = np.concatenate([x, np.zeros((B, 1, I))], 1)
    mask = (ids != p_in)[:, :, None]
    want = np.einsum("pnm,bpni->bpmi", A * mask, xpad[:, ids])               # graphVAESSW.py:111-123
    got = np.zeros_like(want)
    for verts, nbrs in plan:
        K = (len(nbrs) + 7) // 8 * 8                                          # MMA k granularity
        col = {q: k for k, q in enumerate(nbrs)}
        S = np.zeros((len(verts) * M, K))
        for r, p in enumerate(verts):
            for n in range(ids.shape[1]):
                q = int(ids[p, n])
                if q != p_in:
                    S[r * M:(r + 1) * M, col[q]] += A[p, n]
        X = np.zeros((K, B * I))
        X[:len(nbrs)] = x[:, nbrs].transpose(1, 0, 2).reshape(len(nbrs), B * I)
        Z = S @ X                                                             # [(p,m), (b,i)]
        got[:, verts] = Z.reshape(len(verts), M, B, I).transpose(2, 0, 1, 3)
    err = np.abs(got - want).max() / np.abs(want).max()
    print("dense-tile formulation vs direct sum on unpool2 (%d tiles): max rel err %.2e" % (len(plan), err))
    assert err < 1e-12


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
    plan.clear()
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
        plan.append((tile, sorted(u)))
    return np.array(tiles), float(np.mean([len(x) for x in nb]))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "check":
        return check()
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
