"""vcmeshconv_b200 — B200-native vc-conv / vd-pool / vd-unpool hot path of
VCMeshConv: reference module API (graphVAESSW.py) over a C-ABI library of
hand-written sm_100a CUDA kernels. See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"


def __getattr__(name):
    # lazy: importing the package must not require torch / a GPU (build.py, graph_sampling)
    import importlib
    if name in ("graphVAESSW", "functional", "connectivity", "graph_sampling", "synthetic", "build",
                "graphVAE_train", "graphAE_param_iso", "_lib"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
