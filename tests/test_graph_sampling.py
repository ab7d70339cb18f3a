"""Host GraphSampling (libvcmesh_sampling.so) must reproduce the reference tool's
tables bit for bit (including neighbour order inside a row). CPU only."""
import os
import tempfile

import numpy as np
import pytest

from helpers import load
from oracle import graphsampling_ref as ref
from vcmeshconv_b200 import graph_sampling as gs
from vcmeshconv_b200 import synthetic


def test_matches_reference_tables_in_golden():
    # tables recorded from the reference binary by oracle/gen_golden.py (sphere(300, seed=1), 5 levels)
    g = load("stack_sphere300")
    pts, faces = synthetic.sphere(300, seed=1)
    assert np.array_equal(faces, g["faces"])
    tabs, nums = gs.build_tables(faces, 300, synthetic.alternating_levels(5))
    for l in range(5):
        for kind in ("pool", "unpool"):
            want = g["table.%s%d" % (kind, l)]
            got = tabs["%s%d" % (kind, l)]
            assert got.dtype == np.int32 and want.dtype == np.int32
            assert np.array_equal(got, want), (kind, l)
    assert nums[0] == 300 and nums[1] == 300


def test_table_format_invariants():
    pts, faces = synthetic.sphere(500, seed=4)
    tabs, nums = gs.build_tables(faces, 500, synthetic.alternating_levels(4))
    for l in range(4):
        pool, unpool = tabs["pool%d" % l], tabs["unpool%d" % l]
        n_before, n_after = nums[l], nums[l + 1]
        assert pool.shape[0] == n_after and unpool.shape[0] == n_before
        for tab, pad in ((pool, n_before), (unpool, n_after)):
            ids, hops, cnt = tab[:, 1::2], tab[:, 2::2], tab[:, 0]
            n = ids.shape[1]
            real = np.arange(n)[None, :] < cnt[:, None]
            assert np.all(ids[~real] == pad) and np.all(hops[~real] == -1)      # padding = (P_in, -1)
            assert np.all(ids[real] < pad) and np.all(ids[real] >= 0) and np.all(hops[real] >= 0)
            assert cnt.max() == n                                                # N = max count
            for row, c in zip(ids, cnt):                                         # no duplicate ids in a row
                assert len(set(row[:c].tolist())) == c
        assert np.all(pool[:, 2] == 0)                                           # self first, hop 0
        # unpool is the transpose of the radius-r_unpool balls: same edge multiset when radii are equal
        pe = {(p, q) for p in range(n_after) for q in pool[p, 1::2][:pool[p, 0]]}
        ue = {(c, q) for q in range(n_before) for c in unpool[q, 1::2][:unpool[q, 0]]}
        assert pe == ue
    # stride 1 keeps every vertex, in order
    assert nums[1] == nums[0]
    assert np.array_equal(tabs["pool0"][:, 1], np.arange(500))


def test_error_paths():
    with pytest.raises(gs.GraphSamplingError):
        gs.MeshCNN(np.array([[0, 1, 7]], np.int32), 3)                          # index out of range
    cnn = gs.MeshCNN(np.array([[0, 1, 1]], np.int32), 2)                        # < 3 vertices
    with pytest.raises(gs.GraphSamplingError):
        cnn.add_pool_layer(1, 1, 1)
    cnn = gs.MeshCNN(np.array([[0, 1, 2]], np.int32), 3)
    with pytest.raises(gs.GraphSamplingError):
        cnn.add_pool_layer(0, 1, 1)
    with pytest.raises(gs.GraphSamplingError):
        cnn.table(3)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/graphsampling_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["sphere2000", "torus", "tets", "green"])
def test_bit_identical_to_reference_binary(name):
    green = ()
    if name == "sphere2000":
        pts, faces = synthetic.sphere(2000, seed=0)
        levels = synthetic.alternating_levels(7)
    elif name == "torus":
        pts, faces = synthetic.torus(24, 40)
        levels = synthetic.alternating_levels(5)
    elif name == "tets":
        pts, faces = synthetic.tets(700, seed=1)
        levels = [(1, 1, 1), (2, 2, 2), (1, 1, 1)]
    else:
        pts, faces = synthetic.sphere(800, seed=5)
        levels = [(1, 1, 1), (2, 1, 3), (2, 2, 1), (1, 2, 2)]                    # unequal radii too
        green = (0, 17, 333, 799)
    with tempfile.TemporaryDirectory() as d:
        obj = os.path.join(d, "m.obj")
        synthetic.write_obj(obj, pts, faces, green)
        want = ref.run(obj, levels)
    got, _ = gs.build_tables(faces, len(pts), levels, green)
    assert set(got) == set(want)
    for k in want:
        assert got[k].dtype == want[k].dtype == np.int32
        assert np.array_equal(got[k], want[k]), k
