"""Connectivity loader: reference-format table -> canonical buffers and the
derived device tables' host contract. CPU only."""
import numpy as np

from helpers import layer_cases, load
from vcmeshconv_b200 import connectivity as conn


def test_split_table_matches_reference_buffers():
    for case in layer_cases():
        g = load(case)
        cnt, ids = conn.split_table(g["table"])
        assert cnt.dtype == np.float32 and ids.dtype == np.int64
        assert np.array_equal(ids, g["param.neighbor_id_lstlst"])     # bit-exact vs LASMConvssw buffers
        assert np.array_equal(cnt, g["param.neighbor_num_lst"])


def test_reverse_csr_is_exact_transpose():
    for case in layer_cases():
        g = load(case)
        p_in = int(g["in_point_num"])
        d = conn.derive_host(g["table"], p_in)
        idx, n = d["idx"], d["idx"].shape[1]
        edges = {(p, k) for p in range(idx.shape[0]) for k in range(n) if idx[p, k] != p_in}
        seen = set()
        for q in range(p_in):
            slots = d["rev_slot"][d["rev_ptr"][q]:d["rev_ptr"][q + 1]]
            assert np.all(np.diff(slots) > 0)                          # ascending = deterministic order
            for s in slots:
                p, k = divmod(int(s), n)
                assert idx[p, k] == q
                seen.add((p, k))
        assert seen == edges and d["rev_ptr"][-1] == len(edges)
        # degree buckets: a permutation, non-increasing in `last`
        assert sorted(d["perm"].tolist()) == list(range(idx.shape[0]))
        assert np.all(np.diff(d["last"][d["perm"]]) <= 0)
        # GraphSampling tables keep real entries first: last == count
        assert np.array_equal(d["last"], g["table"][:, 0])


def test_sentinel_in_the_middle():
    # tables are masked by the sentinel id, not by the count column (graphVAESSW.py:111-113)
    t = np.array([[2, 0, 0, 3, -1, 1, 1], [1, 3, -1, 3, -1, 2, 1]], np.int32)   # P_in = 3
    d = conn.derive_host(t, 3)
    assert d["last"].tolist() == [3, 3]
    assert d["rev_ptr"].tolist() == [0, 1, 2, 3]
    assert d["rev_slot"].tolist() == [0, 2, 5]
