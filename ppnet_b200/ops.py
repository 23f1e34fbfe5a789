"""Batched device entry points: torch CUDA tensors in, torch CUDA tensors out, one kernel launch each
on torch's current stream.  Thin ctypes shims over include/ppnet_b200.h -- PyTorch is only the
allocator and the stream owner here."""
import ctypes

import torch

from ._lib import PPNetError, check, lib

DOT_FUSED_SKX = 0
DOT_UNFUSED = 1
DEFAULT_BOUND = 224.0        # hard-coded in the reference (process_map.py:384-387)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise PPNetError("%s must be a CUDA tensor (no CPU fallback)" % name)
    if t.dtype != dtype:
        raise PPNetError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise PPNetError("%s must be contiguous" % name)
    return t


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _seg_grouping(n_segs, n_maps, seg_off):
    """Returns (seg_off tensor or None, segs_per_map_or_max)."""
    if seg_off is None:
        if n_maps == 0 or n_segs % max(n_maps, 1):
            raise PPNetError("uniform grouping needs n_segs divisible by n_maps (or pass seg_off)")
        return None, n_segs // n_maps
    _need(seg_off, torch.int64, "seg_off")
    if seg_off.numel() != n_maps + 1:
        raise PPNetError("seg_off must have n_maps + 1 entries")
    longest = int((seg_off[1:] - seg_off[:-1]).max().item()) if n_maps else 0
    return seg_off, max(longest, 1)


def segcheck_edage_f64(pts_rc, obs, obs_cnt, clearance, seg_off=None, bound=DEFAULT_BOUND,
                       dot_mode=DOT_FUSED_SKX, out=None):
    """A11 batched (EDaGe-PP/process_map.py:383-425).  pts_rc f64[N,4] = (s_row,s_col,e_row,e_col),
    obs f64[M,omax,3], obs_cnt i32[M]; segments grouped by map (uniform or CSR).  -> u8[N]."""
    _need(pts_rc, torch.float64, "pts_rc")
    _need(obs, torch.float64, "obs")
    _need(obs_cnt, torch.int32, "obs_cnt")
    n, m, omax = pts_rc.shape[0], obs.shape[0], obs.shape[1]
    so, spm = _seg_grouping(n, m, seg_off)
    out = torch.empty(n, dtype=torch.uint8, device=pts_rc.device) if out is None else _need(out, torch.uint8, "out")
    check(lib().ppnet_segcheck_edage_f64(
        _ptr(pts_rc), ctypes.c_int64(n), _ptr(so), ctypes.c_int64(spm), ctypes.c_int64(m), _ptr(obs),
        _ptr(obs_cnt), ctypes.c_int32(omax), ctypes.c_double(clearance), ctypes.c_double(bound),
        ctypes.c_int32(dot_mode), _ptr(out), _stream()), "ppnet_segcheck_edage_f64")
    return out


def segcheck_mpnet_f32(pts_xy, obs, obs_cnt, clearance, seg_off=None, bound=DEFAULT_BOUND, want_steer=False,
                       out=None):
    """A12 batched (experiments/MPNet/neuralplanner.py:43-69, steerTo :86-92).  pts_xy f32[N,4].
    -> verdict u8[N] (and steer u8[N] when want_steer)."""
    _need(pts_xy, torch.float32, "pts_xy")
    _need(obs, torch.float64, "obs")
    _need(obs_cnt, torch.int32, "obs_cnt")
    n, m, omax = pts_xy.shape[0], obs.shape[0], obs.shape[1]
    so, spm = _seg_grouping(n, m, seg_off)
    out = torch.empty(n, dtype=torch.uint8, device=pts_xy.device) if out is None else _need(out, torch.uint8, "out")
    steer = torch.empty(n, dtype=torch.uint8, device=pts_xy.device) if want_steer else None
    check(lib().ppnet_segcheck_mpnet_f32(
        _ptr(pts_xy), ctypes.c_int64(n), _ptr(so), ctypes.c_int64(spm), ctypes.c_int64(m), _ptr(obs),
        _ptr(obs_cnt), ctypes.c_int32(omax), ctypes.c_double(clearance), ctypes.c_double(bound), _ptr(out),
        _ptr(steer), _stream()),
        "ppnet_segcheck_mpnet_f32")
    return (out, steer) if want_steer else out


def path_feasible_f32(wp, path_off, path_map, obs, obs_cnt, clearance, bound=DEFAULT_BOUND):
    """feasibility_check batched (neuralplanner.py:96-102) -> (feasible u8[P], n_checked i32[P])."""
    _need(wp, torch.float32, "wp")
    _need(path_off, torch.int64, "path_off")
    _need(path_map, torch.int32, "path_map")
    _need(obs, torch.float64, "obs")
    _need(obs_cnt, torch.int32, "obs_cnt")
    p = path_map.numel()
    feas = torch.empty(p, dtype=torch.uint8, device=wp.device)
    chk = torch.empty(p, dtype=torch.int32, device=wp.device)
    check(lib().ppnet_path_feasible_f32(_ptr(wp), _ptr(path_off), _ptr(path_map), ctypes.c_int64(p), _ptr(obs),
                                        _ptr(obs_cnt), ctypes.c_int32(obs.shape[1]), ctypes.c_double(clearance),
                                        ctypes.c_double(bound), _ptr(feas), _ptr(chk), _stream()),
          "ppnet_path_feasible_f32")
    return feas, chk


def lvc_f32(wp, path_off, path_map, obs, obs_cnt, clearance, bound=DEFAULT_BOUND):
    """lvc batched (neuralplanner.py:123-138) -> (out_wp f32 like wp, out_len i32[P])."""
    _need(wp, torch.float32, "wp")
    _need(path_off, torch.int64, "path_off")
    _need(path_map, torch.int32, "path_map")
    _need(obs, torch.float64, "obs")
    _need(obs_cnt, torch.int32, "obs_cnt")
    p = path_map.numel()
    out = torch.zeros_like(wp)
    out_len = torch.empty(p, dtype=torch.int32, device=wp.device)
    check(lib().ppnet_lvc_f32(_ptr(wp), _ptr(path_off), _ptr(path_map), ctypes.c_int64(p), _ptr(obs),
                              _ptr(obs_cnt), ctypes.c_int32(obs.shape[1]), ctypes.c_double(clearance),
                              ctypes.c_double(bound), _ptr(out), _ptr(out_len), _stream()), "ppnet_lvc_f32")
    return out, out_len


def clearance_filter_f64(pathpt, cand, map_size, resolution, clearance):
    """A14 batched (EDaGe-PP/MapGenerate.py:132-143).  pathpt f64[M,Np,2] (row,col), cand f64[M,O,3]
    (x,y,r map units) -> (accept u8[M,O], out f64[M,O,3] compacted [col,row,r] px, out_cnt i32[M])."""
    _need(pathpt, torch.float64, "pathpt")
    _need(cand, torch.float64, "cand")
    m, np_, _ = pathpt.shape
    O = cand.shape[1]
    acc = torch.empty([m, O], dtype=torch.uint8, device=pathpt.device)
    out = torch.zeros([m, O, 3], dtype=torch.float64, device=pathpt.device)
    cnt = torch.empty(m, dtype=torch.int32, device=pathpt.device)
    check(lib().ppnet_clearance_filter_f64(_ptr(pathpt), ctypes.c_int32(np_), _ptr(cand), ctypes.c_int32(O),
                                           ctypes.c_int64(m), ctypes.c_double(map_size),
                                           ctypes.c_double(resolution), ctypes.c_double(clearance), _ptr(acc),
                                           _ptr(out), _ptr(cnt), _stream()), "ppnet_clearance_filter_f64")
    return acc, out, cnt


def grid_index_f64(pts, map_size, resolution, mapoffset):
    """A4 batched (EDaGe-PP/Path.py:378-386): f64[...] -> i32[...] same shape."""
    _need(pts, torch.float64, "pts")
    idx = torch.empty(pts.shape, dtype=torch.int32, device=pts.device)
    check(lib().ppnet_grid_index_f64(_ptr(pts), ctypes.c_int64(pts.numel()), ctypes.c_double(map_size),
                                     ctypes.c_double(resolution), ctypes.c_double(mapoffset), _ptr(idx),
                                     _stream()), "ppnet_grid_index_f64")
    return idx


def corridor_paint(x0, dirs, step_num, map_size, resolution, mapoffset, width, height, value=255, space=None):
    """A5 batched (EDaGe-PP/Path.py:113-134, 397-404): x0/dirs f64[P,rays,2], step_num f64[P]
    -> space u8[P,W,H]."""
    _need(x0, torch.float64, "x0")
    _need(dirs, torch.float64, "dirs")
    _need(step_num, torch.float64, "step_num")
    p, rays, _ = x0.shape
    if space is None:
        space = torch.zeros([p, width, height], dtype=torch.uint8, device=x0.device)
    check(lib().ppnet_corridor_paint(_ptr(x0), _ptr(dirs), _ptr(step_num), ctypes.c_int64(p), ctypes.c_int32(rays),
                                     ctypes.c_double(map_size), ctypes.c_double(resolution),
                                     ctypes.c_double(mapoffset), ctypes.c_int32(width), ctypes.c_int32(height),
                                     ctypes.c_uint8(value), _ptr(space), _stream()), "ppnet_corridor_paint")
    return space
