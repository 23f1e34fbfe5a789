"""experiments/MPNet/neuralplanner.py:43-138 -- the "corrected" collision checker of the MPNet experiment (float32
flavour).  As in the reference, the functions read two module globals: `obc` (per-problem obstacle lists
[[x, y, r], ...], my_dataset.py:37,49) and `clearance` (= 1/50*224, neuralplanner.py:18).

    from ppnet_b200 import mpnet
    mpnet.obc = obc
    mpnet.steerTo(start, end, idx); mpnet.feasibility_check(path, idx); mpnet.lvc(path, idx)
"""
import numpy as np
import torch

from . import ops

clearance = 1 / 50 * 224
obc = []

_snapshot = {"obs": None, "cnt": None}


def _pack(problems):
    """[[x, y, r], ...] per problem -> (obs f64[P, omax, 3], cnt i32[P]) on the GPU."""
    if not torch.cuda.is_available():
        raise ops.PPNetError("ppnet_b200 needs a CUDA device (there is no CPU path)")
    omax = max([len(o) for o in problems] + [1])
    arr = np.zeros([max(len(problems), 1), omax, 3])
    cnt = np.zeros(max(len(problems), 1), dtype=np.int32)
    for i, o in enumerate(problems):
        cnt[i] = len(o)
        if len(o):
            arr[i, :len(o)] = np.asarray(o, dtype=np.float64).reshape(-1, 3)
    return torch.from_numpy(arr).cuda(), torch.from_numpy(cnt).cuda()


def set_obc(problems=None):
    """Explicit device snapshot of the obstacle lists for the *_batch entry points (`problems` defaults to the module
    global `obc`).  The reference-signature functions below do NOT use it: like the reference they read `obc[idx]` at
    every call, so in-place edits of `obc` are always seen."""
    src = obc if problems is None else problems
    _snapshot["obs"], _snapshot["cnt"] = _pack(src)
    return _snapshot["obs"], _snapshot["cnt"]


def _device_obc():
    """Batch entry points: the snapshot taken by set_obc(), or a fresh upload of the whole global."""
    if _snapshot["obs"] is not None:
        return _snapshot["obs"], _snapshot["cnt"]
    return _pack(obc)


def _problem(idx):
    """obc[idx] read live (neuralplanner.py:52 `for ox, oy, size in obc[idx]`) -> (obs f64[1, n, 3], cnt i32[1])."""
    return _pack([obc[idx]])


def _f32(p):
    if isinstance(p, torch.Tensor):
        return p.detach().to(torch.float32).cpu().numpy().reshape(-1)
    return np.asarray(p, dtype=np.float32).reshape(-1)


def _one(s, e, idx, want_steer):
    obs, cnt = _problem(idx)
    s, e = _f32(s), _f32(e)
    pts = torch.from_numpy(np.asarray([[s[0], s[1], e[0], e[1]]], dtype=np.float32)).cuda()
    return ops.segcheck_mpnet_f32(pts, obs, cnt, clearance, want_steer=want_steer)


def collision_check_circle_edge(s, e, idx):
    """neuralplanner.py:43-69 -> True on collision."""
    return bool(_one(s, e, idx, False).item())


def steerTo(start, end, idx):
    """neuralplanner.py:86-92 -> 0 if the edge is blocked (and not degenerate), else 1."""
    return int(_one(start, end, idx, True)[1].item())


def _path_tensors(path, idx):
    wp = np.stack([_f32(p)[:2] for p in path]).astype(np.float32)
    off = torch.tensor([0, len(wp)], dtype=torch.int64).cuda()
    return torch.from_numpy(wp).cuda(), off, torch.tensor([0], dtype=torch.int32).cuda()


def feasibility_check(path, idx):
    """neuralplanner.py:96-102 -> 1 if every consecutive pair steers, else 0."""
    obs, cnt = _problem(idx)
    wp, off, pm = _path_tensors(path, idx)
    return int(ops.path_feasible_f32(wp, off, pm, obs, cnt, clearance)[0].item())


def lvc(path, idx):
    """neuralplanner.py:123-138 (lazy vertex contraction) -> the contracted list of float32 waypoint tensors."""
    if len(path) < 2:
        return path
    obs, cnt = _problem(idx)
    wp, off, pm = _path_tensors(path, idx)
    out, n = ops.lvc_f32(wp, off, pm, obs, cnt, clearance)
    out = out[:int(n.item())].cpu()
    return [out[i] for i in range(out.shape[0])]


def feasibility_check_batch(wp, path_off, path_map):
    """Many paths in one launch: wp f32[total,2], CSR offsets, problem index per path -> (feasible u8[P], n_checked).
    Obstacles: the set_obc() snapshot if one was taken, else `obc` uploaded afresh."""
    obs, cnt = _device_obc()
    return ops.path_feasible_f32(wp, path_off, path_map, obs, cnt, clearance)


def lvc_batch(wp, path_off, path_map):
    obs, cnt = _device_obc()
    return ops.lvc_f32(wp, path_off, path_map, obs, cnt, clearance)
