"""Connectivity loader: reference-format neighbour table -> device tables.

Replaces the chain graphAE_param_iso.py:95-127 (name -> file) ->
MCStructure.__init__ graphVAESSW.py:365-371 (np.load) ->
LASMConvssw.__init__ graphVAESSW.py:44-57 / MCVcoeffs.__init__ :385-396
(column 0 = count, columns 1:: = (id, hop) pairs). The derived tables (int32
padded ids, degree-bucketed processing order, reverse CSR) are built once on
the host by libvcmesh (vcm_tables_create) and stay resident in HBM.
"""
import ctypes as C

import numpy as np

from . import _lib


def split_table(table):
    """(counts float32 [P_out], ids int64 [P_out, N]) exactly as the reference
    keeps them in its `neighbor_num_lst` / `neighbor_id_lstlst` buffers."""
    table = np.asarray(table)
    if table.ndim != 2 or table.shape[1] < 1 or (table.shape[1] - 1) % 2:
        raise ValueError("connection_info must be [P_out, 1+2N], got %s" % (table.shape,))
    counts = table[:, 0].astype(np.float32)
    ids = np.ascontiguousarray(table[:, 1:].reshape(table.shape[0], -1, 2)[:, :, 0]).astype(np.int64)
    return counts, ids


def avg_neighbor_num(table):
    """round(mean count) with the reference's float32 mean (graphVAESSW.py:56)."""
    import torch
    return round(torch.from_numpy(np.asarray(table)[:, 0].astype(np.float32)).mean().item())


class DeviceTables:
    """Owns one vcm_tables handle (device-resident, immutable)."""

    def __init__(self, table, in_point_num, device=None):
        import torch
        table = np.ascontiguousarray(np.asarray(table), dtype=np.int32)
        if table.ndim != 2:
            raise ValueError("connection_info must be 2-D")
        if not torch.cuda.is_available():
            raise RuntimeError("vcmeshconv_b200 needs a CUDA device: there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().vcm_tables_create(table.ctypes.data, table.shape[0], table.shape[1],
                                                    int(in_point_num), C.byref(h)))
        self._h = h
        info = _lib.TablesInfo()
        _lib.check(_lib.lib().vcm_tables_get_info(h, C.byref(info)))
        self.p_in, self.p_out, self.n_max = info.p_in, info.p_out, info.n_max
        self.nnz, self.max_last, self.max_rev = info.nnz, info.max_last, info.max_rev

    @property
    def handle(self):
        return self._h

    def export(self, which):
        sizes = {_lib.TAB_IDX: self.p_out * self.n_max, _lib.TAB_LAST: self.p_out, _lib.TAB_PERM: self.p_out,
                 _lib.TAB_REV_PTR: self.p_in + 1, _lib.TAB_REV_SLOT: self.nnz,
                 _lib.TAB_IDX_SORTED: self.p_out * self.n_max, _lib.TAB_META_SORTED: 2 * self.p_out}
        out = np.empty(sizes[which], dtype=np.int32)
        _lib.check(_lib.lib().vcm_tables_export(self._h, which, out.ctypes.data, out.size))
        return out

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().vcm_tables_destroy(h)
            except Exception:
                pass
            self._h = None


# host restatement of what vcm_tables_create derives (used by the CPU tests and
# documented here as the contract of the device tables)
def derive_host(table, in_point_num):
    table = np.asarray(table)
    ids = table[:, 1:].reshape(table.shape[0], -1, 2)[:, :, 0].astype(np.int64)
    p_out, n = ids.shape
    real = ids != in_point_num
    pos = np.where(real, np.arange(1, n + 1)[None, :], 0)
    last = pos.max(1).astype(np.int32) if n else np.zeros(p_out, np.int32)
    perm = np.argsort(-last.astype(np.int64), kind="stable").astype(np.int32)
    slots = np.nonzero(real.reshape(-1))[0]
    q = ids.reshape(-1)[slots]
    order = np.argsort(q, kind="stable")
    rev_slot = slots[order].astype(np.int32)
    rev_ptr = np.zeros(in_point_num + 1, np.int32)
    np.add.at(rev_ptr, q + 1, 1)
    rev_ptr = np.cumsum(rev_ptr).astype(np.int32)
    return {"idx": ids.astype(np.int32), "last": last, "perm": perm, "rev_ptr": rev_ptr, "rev_slot": rev_slot,
            "idx_sorted": ids.astype(np.int32)[perm], "meta_sorted": np.stack([perm, last[perm]], 1).astype(np.int32)}
