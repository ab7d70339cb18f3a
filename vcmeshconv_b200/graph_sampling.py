"""ctypes front-end of libvcmesh_sampling.so (include/vcmesh_sampling.h).

Replaces running the reference's GraphSampling executable
(GraphSampling/main.cpp:286-328) + np.load of its `_pool{i}.npy` /
`_unpool{i}.npy` files: the tables come back as int32 numpy arrays in the
same [rows, 1+2N] layout (meshCNN.h:54-92).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB_SAMPLING
        if not os.path.exists(path):
            _build.build_sampling()
        L = C.CDLL(path)
        L.vcm_gs_last_error.restype = C.c_char_p
        L.vcm_gs_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.vcm_gs_destroy.argtypes = [C.c_void_p]
        L.vcm_gs_destroy.restype = None
        L.vcm_gs_add_layer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.vcm_gs_num_levels.argtypes = [C.c_void_p]
        L.vcm_gs_table_shape.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.vcm_gs_table_copy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vcm_gs_centres.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib = L
    return _lib


class GraphSamplingError(RuntimeError):
    pass


def _check(code):
    if code < 0:
        raise GraphSamplingError(lib().vcm_gs_last_error().decode())
    return code


class MeshCNN:
    """Same role as the reference's MeshCNN (meshCNN.h:19-49): chain levels,
    then read the per-level pool / unpool tables."""

    def __init__(self, faces, n_vert, must_include=()):
        faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
        must = np.ascontiguousarray(list(must_include), dtype=np.int32)
        h = C.c_void_p()
        _check(lib().vcm_gs_create(faces.ctypes.data, len(faces), int(n_vert),
                                   must.ctypes.data if len(must) else None, len(must), C.byref(h)))
        self._h = h
        self.point_nums = [int(n_vert)]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vcm_gs_destroy(self._h)
            self._h = None

    def add_pool_layer(self, stride, radius_pool, radius_unpool):
        n = _check(lib().vcm_gs_add_layer(self._h, stride, radius_pool, radius_unpool))
        self.point_nums.append(n)
        return n

    @property
    def num_levels(self):
        return lib().vcm_gs_num_levels(self._h)

    def table(self, level, unpool=False):
        r, c = C.c_int(), C.c_int()
        _check(lib().vcm_gs_table_shape(self._h, level, int(unpool), C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), dtype=np.int32)
        _check(lib().vcm_gs_table_copy(self._h, level, int(unpool), out.ctypes.data))
        return out

    def centres(self, level):
        n = _check(lib().vcm_gs_centres(self._h, level, None, 0))
        out = np.empty(n, dtype=np.int32)
        _check(lib().vcm_gs_centres(self._h, level, out.ctypes.data, n))
        return out


def build_tables(faces, n_vert, levels, must_include=()):
    """levels: [(stride, pool_radius, unpool_radius), ...] ->
    {'pool0': arr, 'unpool0': arr, 'pool1': ...} and the per-level vertex counts."""
    cnn = MeshCNN(faces, n_vert, must_include)
    for s, rp, ru in levels:
        cnn.add_pool_layer(s, rp, ru)
    tabs = {}
    for l in range(cnn.num_levels):
        tabs["pool%d" % l] = cnn.table(l, False)
        tabs["unpool%d" % l] = cnn.table(l, True)
    return tabs, list(cnn.point_nums)


def save_tables(tabs, folder):
    """Writes the reference's file names (`_pool{i}.npy`, meshCNN.h:121,129) so the
    reference's config resolution (graphAE_param_iso.py:95-127) finds them."""
    os.makedirs(folder, exist_ok=True)
    paths = {}
    for k, v in tabs.items():
        paths[k] = os.path.join(folder, "_%s.npy" % k)
        np.save(paths[k], v)
    return paths
