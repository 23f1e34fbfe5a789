"""EDaGe-PP/process_map.py -- the two functions on the hot path: the segment-vs-circles verdict (:383-425) and the
init/end stamps (:119-145).  Scalar signatures as in the reference, plus batched forms."""
import numpy as np
import torch

from .. import ops
from .Path import _dev


def _obs_tensor(obs, dev):
    o = np.asarray([[float(v[0]), float(v[1]), float(v[2])] for v in obs], dtype=np.float64).reshape(-1, 3)
    omax = max(len(o), 1)
    t = torch.zeros([1, omax, 3], dtype=torch.float64, device=dev)
    if len(o):
        t[0, :len(o)] = torch.from_numpy(o).to(dev)
    return t, torch.tensor([len(o)], dtype=torch.int32, device=dev)


def collision_check_circle_edge(s, e, obs, clearance):
    """process_map.py:383-425: s, e in (row, col); obs = [[x, y, r], ...]; -> bool (True = collision)."""
    dev = _dev()
    pts = torch.tensor([[float(s[0]), float(s[1]), float(e[0]), float(e[1])]], dtype=torch.float64, device=dev)
    o, c = _obs_tensor(obs, dev)
    return bool(ops.segcheck_edage_f64(pts, o, c, float(clearance)).item())


def collision_check_path(path, obs, clearance):
    """The checker loop of extract_path_image (process_map.py:491-495) in one launch: path = [[row, col], ...]
    -> bool array, one verdict per consecutive pair."""
    dev = _dev()
    p = np.asarray(path, dtype=np.float64).reshape(-1, 2)
    if len(p) < 2:
        return np.zeros(0, dtype=bool)
    pts = torch.from_numpy(np.ascontiguousarray(np.concatenate([p[:-1], p[1:]], axis=1))).to(dev)
    o, c = _obs_tensor(obs, dev)
    return ops.segcheck_edage_f64(pts, o, c, float(clearance)).cpu().numpy().astype(bool)


def add_init_end_single(image, init, end):
    """process_map.py:119-145: 7x7 red squares at round(init), round(end) on image Tensor[3,R,R] (in place when the
    tensor already lives on the GPU; a CPU tensor is updated through a device copy)."""
    assert len(image.shape) == 3, "Image shape incorrect"
    assert init is not None, "Init is None"
    assert end is not None, "End is None"
    dev = _dev()
    img = image.to(device=dev, dtype=torch.float32)
    work = img.contiguous()[None]
    mk = lambda p: torch.tensor([[float(p[0]), float(p[1])]], dtype=torch.float64, device=dev)
    ops.add_init_end(work, mk(init), mk(end))
    if work.data_ptr() != image.data_ptr():
        image.copy_(work[0].to(device=image.device, dtype=image.dtype))
    return image
