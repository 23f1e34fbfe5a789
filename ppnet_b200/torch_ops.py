"""torch.ops.ppnet_b200.* -- the dispatcher-registered form of the hot-path entry points.

    from ppnet_b200 import torch_ops          # loads ppnet_b200/lib/libppnet_torch.so (TORCH_LIBRARY registration)
    b64, b32 = torch.ops.ppnet_b200.verdict_fused(pts_rc, obs, obs_cnt, 4.48)

The ops are registered for the CUDA dispatch key only: CPU tensors raise (there is no CPU path).  Each op switches to its
tensors' device (CUDAGuard) and launches on torch's current stream there.  `ppnet_b200.ops` (ctypes) calls the same C ABI
with explicit output buffers; measured per-call overheads are in DESIGN.md."""
import os

import torch

from ._lib import PPNetError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libppnet_torch.so")
_loaded = False

OPS = ("segcheck_edage_f64", "segcheck_mpnet_f32", "verdict_fused", "dda_gridcheck", "dda_gridcheck_rc64", "compact_bits",
       "grid_index_f64", "clearance_filter_f64", "raster_circles_bits", "gmm_sample", "propose_segments")


def load():
    """Registers the ops (idempotent).  Fails loudly when the extension has not been built."""
    global _loaded
    if not _loaded:
        if not os.path.exists(LIB_PATH):
            raise PPNetError("ppnet_b200: %s is missing -- build it with `python ppnet_b200/csrc/torch/build.py` "
                             "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        torch.ops.load_library(LIB_PATH)
        _loaded = True
    return torch.ops.ppnet_b200


load()
