"""ppnet_b200 -- B200-native (sm_100a) implementation of PPNet's EDaGe-PP data-generation hot path.

`ppnet_b200.ops` holds the batched device-tensor entry points (thin ctypes shims over the C ABI in
include/ppnet_b200.h); the modules PathSeg / Path / PathGenerate / MapGenerate / GMM / process_map /
mpnet mirror the reference's Python surface (same names, argument meaning and return values) and
dispatch to the same kernels.  No CPU fallback anywhere."""
from ._lib import PPNetError, launch_count  # noqa: F401

__version__ = "0.1.0"
