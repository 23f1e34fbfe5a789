"""Drop-in mirrors of the reference's EDaGe-PP modules: same class / function names, argument order, defaults and
return types; the numerical bodies run as sm_100a CUDA kernels through the C ABI
(include/ppnet_b200.h).  There is no CPU path: a CUDA device is required.

    reference                                   here
    EDaGe-PP/GMM.py           GMM               ppnet_b200.edage.GMM.GMM
    EDaGe-PP/PathSeg.py       PathSeg           ppnet_b200.edage.PathSeg.PathSeg
    EDaGe-PP/Path.py          Path, plot_obstacles   ppnet_b200.edage.Path
    EDaGe-PP/PathGenerate.py  PathGroup         ppnet_b200.edage.PathGenerate.PathGroup
    EDaGe-PP/MapGenerate.py   MapGenerate       ppnet_b200.edage.MapGenerate.MapGenerate
    EDaGe-PP/process_map.py   collision_check_circle_edge, add_init_end_single   ppnet_b200.edage.process_map

Randomness: the reference draws from the process-global np.random / torch streams.  Here every draw is
counter-based Philox keyed by (seed, object id); `seed(s)` below replaces `np.random.seed(s); torch.manual_seed(s)`.
"""
from . import _state
from ._state import seed

__all__ = ["seed", "_state"]
