"""The C-ABI libraries load and export exactly what include/*.h declares.
No compute call is made: this runs without a GPU."""
import ctypes
import os
import re

from vcmeshconv_b200 import _lib, build
from vcmeshconv_b200 import graph_sampling as gs


def test_cuda_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _lib.declared_symbols()
    assert len(declared) >= 19
    assert set(declared) == set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(L, name), name
    assert L.vcm_version() >= 100
    assert isinstance(L.vcm_last_error(), bytes)


def test_no_torch_types_in_the_abi():
    hdr = open(os.path.join(build.ROOT, "include", "vcmesh.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)            # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and "Tensor" not in code


def test_bad_arguments_return_codes_not_exceptions():
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.vcm_tables_create(None, 1, 3, 1, ctypes.byref(h)) == -1
    assert b"NULL" in L.vcm_last_error()
    arr = (ctypes.c_int32 * 4)(1, 0, 0, 0)
    assert L.vcm_tables_create(arr, 1, 4, 1, ctypes.byref(h)) == -1       # cols must be 1 + 2N
    assert L.vcm_basis_bwd_dw_workspace(0, 17, 32, 32, 0) == 0
    assert L.vcm_basis_bwd_dw_workspace(4096, 17, 32, 32, 0) > 0
    assert L.vcm_basis_uses_tensor_cores(4096, 17, 32, 32, 0) == 0        # fp32 mode never uses tensor cores


def test_sampling_library_exports():
    L = gs.lib()
    hdr = open(os.path.join(build.ROOT, "include", "vcmesh_sampling.h")).read()
    names = sorted(set(re.findall(r"VCM_GS_API[^;(]*?\b(vcm_gs_\w+)\s*\(", hdr)))
    assert len(names) == 8
    for n in names:
        assert hasattr(L, n), n
