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
    if t.device.index != torch.cuda.current_device():
        # the launch goes to the current device's current stream: a tensor elsewhere would fault (or be reached through a
        # peer mapping from the wrong GPU).  torch.ops.ppnet_b200.* (torch_ops.py) switches devices itself.
        raise PPNetError("%s lives on %s but the current device is cuda:%d -- wrap the call in torch.cuda.device(%r)"
                         % (name, t.device, torch.cuda.current_device(), str(t.device)))
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
    if n_maps:
        d = seg_off[1:] - seg_off[:-1]
        lo, hi, first, last = (int(v) for v in torch.stack([d.min(), d.max(), seg_off[0], seg_off[-1]]).tolist())
        if first != 0 or lo < 0 or last != n_segs:          # the kernels index pts / verdict with it
            raise PPNetError("seg_off must start at 0, be non-decreasing and end at n_segs (%d)" % n_segs)
    else:
        hi = 0
    return seg_off, max(hi, 1)


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


CMP_F32_NEP50, CMP_F64_NUMPY1 = 0, 1


def verdict_fused(pts_rc, obs, obs_cnt, clearance, seg_off=None, bound=DEFAULT_BOUND, dot_mode=DOT_FUSED_SKX,
                  cmp_mode=CMP_F32_NEP50, want=("bits64", "bits32"), out=None):
    """A11 + A12 fused on one read of the f64 (row, col) segments (ppnet_verdict_fused).  The A12 flavour runs on
    (x, y) = (float32(col), float32(row)).  `want` picks the outputs: "u8_64" / "u8_32" (uint8[N]) and "bits64" /
    "bits32" (int32[ceil(N/32)], bit i & 31 of word i >> 5 = segment i).  -> dict of tensors (pass `out` to reuse)."""
    _need(pts_rc, torch.float64, "pts_rc")
    _need(obs, torch.float64, "obs")
    _need(obs_cnt, torch.int32, "obs_cnt")
    n, m, omax = pts_rc.shape[0], obs.shape[0], obs.shape[1]
    so, spm = _seg_grouping(n, m, seg_off)
    out = {} if out is None else out
    for k in want:
        if k not in ("u8_64", "u8_32", "bits64", "bits32"):
            raise PPNetError("verdict_fused: unknown output %r" % (k,))
        if k not in out:
            out[k] = (torch.empty(n, dtype=torch.uint8, device=pts_rc.device) if k.startswith("u8") else
                      torch.empty((n + 31) // 32, dtype=torch.int32, device=pts_rc.device))
        _need(out[k], torch.uint8 if k.startswith("u8") else torch.int32, k)
    g = lambda k: _ptr(out[k]) if k in want else ctypes.c_void_p(0)
    check(lib().ppnet_verdict_fused(
        _ptr(pts_rc), ctypes.c_int64(n), _ptr(so), ctypes.c_int64(spm), ctypes.c_int64(m), _ptr(obs), _ptr(obs_cnt),
        ctypes.c_int32(omax), ctypes.c_double(clearance), ctypes.c_double(bound), ctypes.c_int32(dot_mode),
        ctypes.c_int32(cmp_mode), g("u8_64"), g("u8_32"), g("bits64"), g("bits32"), _stream()), "ppnet_verdict_fused")
    return out


def unpack_bits(words, n):
    """int32/uint32 words [ceil(n/32)] -> uint8[n] (torch, on the words' device): bit i & 31 of word i >> 5."""
    w = words.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    sh = torch.arange(32, device=words.device, dtype=torch.int64)
    return ((w[:, None] >> sh[None, :]) & 1).to(torch.uint8).reshape(-1)[:n]


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


# ------------------------------------------------------------------------------------------------
STREAM_PLACE, STREAM_OBST, STREAM_GMM_PARAM, STREAM_GMM_SAMPLE, STREAM_UNIFORM, STREAM_PATH = 1, 2, 3, 4, 5, 6
DEFAULT_SEED = 0x5050_4E45_54          # "PPNET"


def hull2d_i32(pts, hmax=64):
    """A6 batched (EDaGe-PP/Path.py:388-395): pts i32[P,Np,2] -> (hull i32[P,hmax,2] CCW, cnt i32[P])."""
    _need(pts, torch.int32, "pts")
    p, np_, _ = pts.shape
    hull = torch.zeros([p, hmax, 2], dtype=torch.int32, device=pts.device)
    cnt = torch.empty(p, dtype=torch.int32, device=pts.device)
    check(lib().ppnet_hull2d_i32(_ptr(pts), ctypes.c_int32(np_), ctypes.c_int64(p), ctypes.c_int32(hmax), _ptr(hull),
                                 _ptr(cnt), _stream()), "ppnet_hull2d_i32")
    return hull, cnt


def boundary_check(hull, hull_cnt, path_idx, angle_arg, trans_arg, resolution, want_hull=False):
    """A10 batched (EDaGe-PP/Path.py:100-111): hull f64[B,hmax,2], queries (path_idx i32[N] or None,
    angle_arg f64[N] degrees as passed to the method, trans_arg f64[N,2]) -> ok u8[N] (, hull' f64[N,hmax,2])."""
    _need(hull, torch.float64, "hull")
    _need(hull_cnt, torch.int32, "hull_cnt")
    _need(angle_arg, torch.float64, "angle_arg")
    _need(trans_arg, torch.float64, "trans_arg")
    if path_idx is not None:
        _need(path_idx, torch.int32, "path_idx")
    n, hmax = angle_arg.numel(), hull.shape[1]
    ok = torch.empty(n, dtype=torch.uint8, device=hull.device)
    ho = torch.zeros([n, hmax, 2], dtype=torch.float64, device=hull.device) if want_hull else None
    check(lib().ppnet_boundary_check(_ptr(hull), _ptr(hull_cnt), ctypes.c_int32(hmax), _ptr(path_idx), _ptr(angle_arg),
                                     _ptr(trans_arg), ctypes.c_int64(n), ctypes.c_double(resolution), _ptr(ok), _ptr(ho),
                                     _stream()), "ppnet_boundary_check")
    return (ok, ho) if want_hull else ok


def uniform_f64(seed, stream_id, unit0, n_units, per_unit, device="cuda"):
    """Philox4x32-10 uniforms: f64[n_units, per_unit] in [0,1), a pure function of (seed, stream, unit, k)."""
    out = torch.empty([n_units, per_unit], dtype=torch.float64, device=device)
    check(lib().ppnet_uniform_f64(ctypes.c_uint64(seed), ctypes.c_uint32(stream_id), ctypes.c_uint64(unit0),
                                  ctypes.c_int64(n_units), ctypes.c_int32(per_unit), _ptr(out), _stream()),
          "ppnet_uniform_f64")
    return out


def gmm_params(seed, order=10, dim=2, mean_range=70.0, std_range=5.0, device="cuda"):
    """A17 GMM.__init__ draws (EDaGe-PP/GMM.py:11-13) -> (mean f32[K,D], std f32[K,D], weights f32[K])."""
    mean = torch.empty([order, dim], dtype=torch.float32, device=device)
    std = torch.empty([order, dim], dtype=torch.float32, device=device)
    w = torch.empty([order], dtype=torch.float32, device=device)
    check(lib().ppnet_gmm_params(ctypes.c_uint64(seed), ctypes.c_int32(order), ctypes.c_int32(dim),
                                 ctypes.c_float(mean_range), ctypes.c_float(std_range), _ptr(mean), _ptr(std), _ptr(w),
                                 _stream()), "ppnet_gmm_params")
    return mean, std, w


def gmm_sample(seed, sample0, n, mean, std, weights, want_comp=False, out=None):
    """A17 Distribution.sample([n]) (EDaGe-PP/GMM.py:14-16) -> f32[n, D] (, comp i32[n])."""
    _need(mean, torch.float32, "mean")
    _need(std, torch.float32, "std")
    _need(weights, torch.float32, "weights")
    k, d = mean.shape
    out = torch.empty([n, d], dtype=torch.float32, device=mean.device) if out is None else _need(out, torch.float32, "out")
    comp = torch.empty([n], dtype=torch.int32, device=mean.device) if want_comp else None
    check(lib().ppnet_gmm_sample(ctypes.c_uint64(seed), ctypes.c_uint64(sample0), ctypes.c_int64(n), ctypes.c_int32(k),
                                 ctypes.c_int32(d), _ptr(mean), _ptr(std), _ptr(weights), _ptr(out), _ptr(comp),
                                 _stream()), "ppnet_gmm_sample")
    return (out, comp) if want_comp else out


def raster_circles_bits(obs, obs_cnt, resolution, inflate=0.0, out=None):
    """A15 restated: obs f64[M,omax,3], obs_cnt i32[M] -> bits u32 (as int32 tensor) [M,R,ceil(R/32)]."""
    _need(obs, torch.float64, "obs")
    _need(obs_cnt, torch.int32, "obs_cnt")
    m, omax, _ = obs.shape
    w = (resolution + 31) // 32
    bits = torch.empty([m, resolution, w], dtype=torch.int32, device=obs.device) if out is None else out
    check(lib().ppnet_raster_circles_bits(_ptr(obs), _ptr(obs_cnt), ctypes.c_int32(omax), ctypes.c_int64(m),
                                          ctypes.c_int32(resolution), ctypes.c_double(inflate), _ptr(bits), _stream()),
          "ppnet_raster_circles_bits")
    return bits


def raster_canvas_bits(obs, obs_cnt, size, resolution, inflate=0.0):
    """A15 in the canvas model (ppnet_raster_canvas_bits): obs f64[M,omax,3] in the units of `size` = (w, h)
    -> bits i32[M,R,ceil(R/32)]."""
    _need(obs, torch.float64, "obs")
    _need(obs_cnt, torch.int32, "obs_cnt")
    m, omax, _ = obs.shape
    bits = torch.empty([m, resolution, (resolution + 31) // 32], dtype=torch.int32, device=obs.device)
    check(lib().ppnet_raster_canvas_bits(_ptr(obs), _ptr(obs_cnt), ctypes.c_int32(omax), ctypes.c_int64(m),
                                         ctypes.c_double(size[0]), ctypes.c_double(size[1]), ctypes.c_int32(resolution),
                                         ctypes.c_double(inflate), _ptr(bits), _stream()), "ppnet_raster_canvas_bits")
    return bits


def dda_gridcheck(bits, resolution, segs_xy, seg_off=None, want_first=True, max_segs_per_map=None, out=None,
                  first_out=None):
    """Integer DDA vs bit-packed maps: bits i32[M,R,W], segs f32[N,4] grouped by map -> (verdict u8[N], first i32[N])."""
    _need(bits, torch.int32, "bits")
    _need(segs_xy, torch.float32, "segs_xy")
    m, n = bits.shape[0], segs_xy.shape[0]
    if seg_off is None:
        so, spm = None, n // max(m, 1)
    else:
        so, longest = _seg_grouping(n, m, seg_off)              # validated: starts at 0, non-decreasing, ends at n
        spm = max(max_segs_per_map, longest) if max_segs_per_map is not None else longest
    v = torch.empty(n, dtype=torch.uint8, device=bits.device) if out is None else _need(out, torch.uint8, "out")
    fh = None
    if want_first:
        fh = torch.empty(n, dtype=torch.int32, device=bits.device) if first_out is None else _need(first_out, torch.int32, "first_out")
    check(lib().ppnet_dda_gridcheck(_ptr(bits), ctypes.c_int32(resolution), ctypes.c_int64(m), _ptr(segs_xy),
                                    ctypes.c_int64(n), _ptr(so), ctypes.c_int64(spm), _ptr(v), _ptr(fh), _stream()),
          "ppnet_dda_gridcheck")
    return (v, fh) if want_first else v


def dda_gridcheck_rc64(bits, resolution, segs_rc, seg_off=None, want=("bits",), max_segs_per_map=None, out=None):
    """The DDA on the A11 array read directly (ppnet_dda_gridcheck_rc64): segs_rc f64[N,4] = (s_row, s_col, e_row,
    e_col), walked as (x, y) = (float32(col), float32(row)).  `want`: any of "u8" (uint8[N]), "bits"
    (int32[ceil(N/32)]), "first" (int32[N]).  -> dict of tensors."""
    _need(bits, torch.int32, "bits")
    _need(segs_rc, torch.float64, "segs_rc")
    m, n = bits.shape[0], segs_rc.shape[0]
    if seg_off is None:
        so, spm = None, n // max(m, 1)
    else:
        so, longest = _seg_grouping(n, m, seg_off)              # validated: starts at 0, non-decreasing, ends at n
        spm = max(max_segs_per_map, longest) if max_segs_per_map is not None else longest
    out = {} if out is None else out
    shapes = {"u8": (n, torch.uint8), "bits": ((n + 31) // 32, torch.int32), "first": (n, torch.int32)}
    for k in want:
        if k not in shapes:
            raise PPNetError("dda_gridcheck_rc64: unknown output %r" % (k,))
        if k not in out:
            out[k] = torch.empty(shapes[k][0], dtype=shapes[k][1], device=bits.device)
        _need(out[k], shapes[k][1], k)
    g = lambda k: _ptr(out[k]) if k in want else ctypes.c_void_p(0)
    check(lib().ppnet_dda_gridcheck_rc64(_ptr(bits), ctypes.c_int32(resolution), ctypes.c_int64(m), _ptr(segs_rc),
                                         ctypes.c_int64(n), _ptr(so), ctypes.c_int64(spm), g("u8"), g("first"), g("bits"),
                                         _stream()), "ppnet_dda_gridcheck_rc64")
    return out


class GenParams(ctypes.Structure):
    """Mirror of `ppnet_gen_params` (include/ppnet_b200.h)."""
    _fields_ = [
        ("bank_pathpt", ctypes.c_void_p), ("bank_segpt", ctypes.c_void_p), ("bank_hull", ctypes.c_void_p),
        ("bank_hull_cnt", ctypes.c_void_p), ("bank_obs", ctypes.c_void_p), ("bank_obs_cnt", ctypes.c_void_p),
        ("n_bank", ctypes.c_int32), ("np", ctypes.c_int32), ("nseg1", ctypes.c_int32), ("hmax", ctypes.c_int32),
        ("pomax", ctypes.c_int32),
        ("map0", ctypes.c_int64), ("n_maps", ctypes.c_int64),
        ("reps", ctypes.c_int32), ("obstacles_num", ctypes.c_int32), ("max_tries", ctypes.c_int32),
        ("reserved0", ctypes.c_int32),
        ("resolution", ctypes.c_double), ("map_size", ctypes.c_double), ("obstacle_size", ctypes.c_double),
        ("clearance", ctypes.c_double), ("raster_inflate", ctypes.c_double),
        ("seed", ctypes.c_uint64),
        ("in_angle", ctypes.c_void_p), ("in_trans", ctypes.c_void_p), ("in_cand", ctypes.c_void_p),
        ("out_angle", ctypes.c_void_p), ("out_trans", ctypes.c_void_p), ("out_segpt", ctypes.c_void_p),
        ("out_pathpt", ctypes.c_void_p), ("out_obs", ctypes.c_void_p), ("out_obs_cnt", ctypes.c_void_p),
        ("out_rand_cnt", ctypes.c_void_p), ("out_bits", ctypes.c_void_p), ("out_tries", ctypes.c_void_p),
        ("out_valid", ctypes.c_void_p), ("counters", ctypes.c_void_p),
    ]


class PathBank:
    """Device-resident target-path bank consumed by generate_maps (what PathGroup.generate produces):
    PathPoint f64[B,Np,2], SegPointImage f64[B,S+1,2], ConvexHull f64[B,hmax,2] + cnt, obstacles f64[B,pomax,3] + cnt."""

    def __init__(self, pathpt, segpt, hull, hull_cnt, obs=None, obs_cnt=None, length=None):
        self.pathpt = _need(pathpt, torch.float64, "pathpt")
        self.segpt = _need(segpt, torch.float64, "segpt")
        self.hull = _need(hull, torch.float64, "hull")
        self.hull_cnt = _need(hull_cnt, torch.int32, "hull_cnt")
        b = pathpt.shape[0]
        if obs is None:
            obs = torch.zeros([b, 1, 3], dtype=torch.float64, device=pathpt.device)
            obs_cnt = torch.zeros([b], dtype=torch.int32, device=pathpt.device)
        self.obs = _need(obs, torch.float64, "obs")
        self.obs_cnt = _need(obs_cnt, torch.int32, "obs_cnt")
        self.length = length
        self.n_bank, self.np, self.nseg1 = b, pathpt.shape[1], segpt.shape[1]
        self.hmax, self.pomax = hull.shape[1], self.obs.shape[1]


class MapBatch:
    """Outputs of one generate_maps launch (device tensors)."""
    __slots__ = ("angle", "trans", "segpt", "pathpt", "obs", "obs_cnt", "rand_cnt", "bits", "tries", "valid",
                 "counters", "map0", "n_maps")


def generate_maps(bank, map0, n_maps, reps, obstacles_num, resolution=224, map_size=50.0, obstacle_size=5.0,
                  clearance=1.0, seed=DEFAULT_SEED, max_tries=4096, want_bits=True, raster_inflate=0.0,
                  in_angle=None, in_trans=None, in_cand=None, out=None, counters=None, want_labels=True):
    """Fused A10+A13+A14(+A15) generator (EDaGe-PP/MapGenerate.py:58-151): one launch, one CTA per map.
    Deterministic in (seed, global map index): any split of [map0, map0+n_maps) over ranks gives the same maps."""
    dev_ = bank.pathpt.device
    R, O = int(resolution), int(obstacles_num)
    if out is None:
        out = MapBatch()
        out.angle = torch.empty([n_maps], dtype=torch.float64, device=dev_)
        out.trans = torch.empty([n_maps, 2], dtype=torch.int32, device=dev_)
        out.segpt = torch.empty([n_maps, bank.nseg1, 2], dtype=torch.float64, device=dev_) if want_labels else None
        out.pathpt = torch.empty([n_maps, bank.np, 2], dtype=torch.float64, device=dev_) if want_labels else None
        out.obs = torch.zeros([n_maps, O + bank.pomax, 3], dtype=torch.float64, device=dev_)
        out.obs_cnt = torch.empty([n_maps], dtype=torch.int32, device=dev_)
        out.rand_cnt = torch.empty([n_maps], dtype=torch.int32, device=dev_)
        out.bits = torch.empty([n_maps, R, (R + 31) // 32], dtype=torch.int32, device=dev_) if want_bits else None
        out.tries = torch.empty([n_maps], dtype=torch.int32, device=dev_)
        out.valid = torch.empty([n_maps], dtype=torch.uint8, device=dev_)
        out.counters = counters if counters is not None else torch.zeros([4], dtype=torch.int64, device=dev_)
    out.map0, out.n_maps = map0, n_maps
    if in_angle is not None:
        _need(in_angle, torch.float64, "in_angle")
        _need(in_trans, torch.int32, "in_trans")
    if in_cand is not None:
        _need(in_cand, torch.float64, "in_cand")
    p = GenParams()
    p.bank_pathpt, p.bank_segpt, p.bank_hull = bank.pathpt.data_ptr(), bank.segpt.data_ptr(), bank.hull.data_ptr()
    p.bank_hull_cnt, p.bank_obs, p.bank_obs_cnt = bank.hull_cnt.data_ptr(), bank.obs.data_ptr(), bank.obs_cnt.data_ptr()
    p.n_bank, p.np, p.nseg1, p.hmax, p.pomax = bank.n_bank, bank.np, bank.nseg1, bank.hmax, bank.pomax
    p.map0, p.n_maps, p.reps, p.obstacles_num, p.max_tries = map0, n_maps, reps, O, max_tries
    p.resolution, p.map_size, p.obstacle_size, p.clearance = float(R), float(map_size), float(obstacle_size), float(clearance)
    p.raster_inflate, p.seed = float(raster_inflate), seed
    p.in_angle = in_angle.data_ptr() if in_angle is not None else None
    p.in_trans = in_trans.data_ptr() if in_trans is not None else None
    p.in_cand = in_cand.data_ptr() if in_cand is not None else None
    for name in ("angle", "trans", "segpt", "pathpt", "obs", "obs_cnt", "rand_cnt", "bits", "tries", "valid"):
        t = getattr(out, name)
        setattr(p, "out_" + name, t.data_ptr() if t is not None else None)
    p.counters = out.counters.data_ptr()
    check(lib().ppnet_generate_maps(ctypes.byref(p), _stream()), "ppnet_generate_maps")
    return out


# ------------------------------------------------------------------------------------------------
# target-path synthesis (A1-A3, A5-A9)
_PATH_FIELDS = [
    ("path0", ctypes.c_int64), ("n_paths", ctypes.c_int64),
    ("seg_num", ctypes.c_int32), ("poly_order", ctypes.c_int32), ("hmax", ctypes.c_int32), ("pomax", ctypes.c_int32),
    ("max_obst_iter", ctypes.c_int32), ("max_obst_rand", ctypes.c_int32),
    ("clearance", ctypes.c_double), ("map_size", ctypes.c_double), ("resolution", ctypes.c_double),
    ("width_coef", ctypes.c_double),
    ("seed", ctypes.c_uint64),
]
_PATH_INPUTS = ["force_straight", "in_straight", "in_y", "in_uend", "in_poly", "in_obst_rand", "in_obst_rand_cnt", "in_hull",
                "in_hull_cnt"]
# name, dtype, shape as a function of dims d = dict(n, S, C, Np, Nb, h, po, R)
_PATH_OUTPUTS = [
    ("poly", torch.float64, lambda d: [d["n"], d["S"], d["C"]]),
    ("endpoint", torch.float64, lambda d: [d["n"], d["S"]]),
    ("is_straight", torch.uint8, lambda d: [d["n"], d["S"]]),
    ("path_straight", torch.uint8, lambda d: [d["n"]]),
    ("seg_trans_local", torch.float64, lambda d: [d["n"], d["S"], 2]),
    ("grad_st", torch.float64, lambda d: [d["n"], d["S"]]),
    ("grad_end", torch.float64, lambda d: [d["n"], d["S"]]),
    ("seg_length", torch.float64, lambda d: [d["n"], d["S"]]),
    ("seg_rot", torch.float64, lambda d: [d["n"], d["S"]]),
    ("seg_trans", torch.float64, lambda d: [d["n"], d["S"], 2]),
    ("segpoint_raw", torch.float64, lambda d: [d["n"], d["S"] + 1, 2]),
    ("pathpoint_raw", torch.float64, lambda d: [d["n"], d["Np"], 2]),
    ("length", torch.float64, lambda d: [d["n"]]),
    ("cells", torch.int32, lambda d: [d["n"], d["Np"], 2]),
    ("up", torch.float64, lambda d: [d["n"], d["S"], 50, 2]),
    ("up_dir", torch.float64, lambda d: [d["n"], d["S"], 50, 2]),
    ("down", torch.float64, lambda d: [d["n"], d["S"], 50, 2]),
    ("cap_init", torch.float64, lambda d: [d["n"], 50, 2]),
    ("cap_end", torch.float64, lambda d: [d["n"], 50, 2]),
    ("boundary_raw", torch.float64, lambda d: [d["n"], d["Nb"], 2]),
    ("ray_x0", torch.float64, lambda d: [d["n"], d["Nb"], 2]),
    ("ray_dir", torch.float64, lambda d: [d["n"], d["Nb"], 2]),
    ("step_num", torch.float64, lambda d: [d["n"]]),
    ("space_raw", torch.uint8, lambda d: [d["n"], 2 * d["R"], 2 * d["R"]]),
    ("space", torch.uint8, lambda d: [d["n"], d["R"], d["R"]]),
    ("hull_raw", torch.int32, lambda d: [d["n"], d["h"], 2]),
    ("hull_cnt", torch.int32, lambda d: [d["n"]]),
    ("rotation", torch.float64, lambda d: [d["n"]]),
    ("translation", torch.float64, lambda d: [d["n"], 2]),
    ("neg_rotation_ws", torch.float64, lambda d: [d["n"]]),
    ("hull", torch.float64, lambda d: [d["n"], d["h"], 2]),
    ("segpoint_img", torch.float64, lambda d: [d["n"], d["S"] + 1, 2]),
    ("pathpoint", torch.float64, lambda d: [d["n"], d["Np"], 2]),
    ("boundary", torch.float64, lambda d: [d["n"], d["Nb"], 2]),
    ("isle", torch.int32, lambda d: [d["n"], d["h"], 2]),
    ("isle_cnt", torch.int32, lambda d: [d["n"]]),
    ("obs", torch.float64, lambda d: [d["n"], d["po"], 3]),
    ("obs_cnt", torch.int32, lambda d: [d["n"]]),
    ("obst_rand_used", torch.int32, lambda d: [d["n"]]),
    ("status", torch.int32, lambda d: [d["n"]]),
]


class PathParams(ctypes.Structure):
    """Mirror of `ppnet_path_params` (include/ppnet_b200.h)."""
    _fields_ = (_PATH_FIELDS + [(k, ctypes.c_void_p) for k in _PATH_INPUTS] +
                [(k, ctypes.c_void_p) for k, _, _ in _PATH_OUTPUTS])


class PathBatch:
    """Outputs of one path_synthesize launch (device tensors, one attribute per `ppnet_path_params` output).  The tensors
    are views of ONE flat device buffer (`_flat`), so `to_host()` is a single device->host copy."""

    def to_host(self):
        """-> dict name -> numpy array (views of one host copy of the flat buffer)."""
        flat = self._flat.cpu().numpy()
        out = {}
        for name, (off, nbytes, dt, shape) in self._layout.items():
            out[name] = flat[off:off + nbytes].view(dt).reshape(shape)
        return out

    def to_bank(self):
        """The target-path bank generate_maps consumes (PathPoint, SegPointImage, ConvexHull, obstacles)."""
        return PathBank(self.pathpoint, self.segpoint_img, self.hull, self.hull_cnt, self.obs, self.obs_cnt,
                        length=self.length)


def path_synthesize(path0, n_paths, seg_num=10, poly_order=4, clearance=1.0, map_size=50.0, resolution=224,
                    seed=DEFAULT_SEED, hmax=64, pomax=32, max_obst_iter=256, want_space=False, device="cuda", width_coef=0.2,
                    force_straight=None, in_straight=None, in_y=None, in_uend=None, in_poly=None, in_obst_rand=None,
                    in_obst_rand_cnt=None, in_hull=None, in_hull_cnt=None):
    """PathGroup.generate's per-path work for global path ids [path0, path0 + n_paths)
    (EDaGe-PP/PathGenerate.py:33-50): PathSeg.random x S -> Path.generate -> draw_boundary -> path_space rays
    (+ paint) -> convexhull -> space_normalization (points) -> search_isle -> set_obstacles.
    Deterministic in (seed, path id)."""
    R = int(resolution)
    S, C = int(seg_num), int(poly_order) + 1
    dims = dict(n=n_paths, S=S, C=C, Np=100 * S, Nb=100 * S + 100, h=hmax, po=pomax, R=R)
    out = PathBatch()
    p = PathParams()
    p.path0, p.n_paths, p.seg_num, p.poly_order, p.hmax, p.pomax = path0, n_paths, S, int(poly_order), hmax, pomax
    p.max_obst_iter, p.clearance, p.map_size, p.resolution = max_obst_iter, float(clearance), float(map_size), float(R)
    p.seed = seed
    p.width_coef = float(width_coef)
    ins = dict(force_straight=(force_straight, torch.uint8), in_straight=(in_straight, torch.uint8),
               in_y=(in_y, torch.float64), in_uend=(in_uend, torch.float64), in_poly=(in_poly, torch.float64),
               in_obst_rand=(in_obst_rand, torch.float32), in_obst_rand_cnt=(in_obst_rand_cnt, torch.int32),
               in_hull=(in_hull, torch.int32), in_hull_cnt=(in_hull_cnt, torch.int32))
    for name, (t, dt) in ins.items():
        if t is not None:
            _need(t, dt, name)
            setattr(p, name, t.data_ptr())
    p.max_obst_rand = in_obst_rand.shape[1] if in_obst_rand is not None else 0
    if in_hull is not None and in_hull.shape[1] != hmax:
        raise PPNetError("in_hull must be [n, hmax, 2]")
    # one zeroed flat buffer, 256-byte aligned sub-ranges viewed as the typed outputs (one allocation, one memset, and one
    # copy when the host mirror wants everything)
    layout, off = {}, 0
    np_dt = {torch.float64: "<f8", torch.int32: "<i4", torch.uint8: "u1"}
    for name, dt, shp in _PATH_OUTPUTS:
        if name in ("space_raw", "space") and not want_space:
            continue
        shape = shp(dims)
        nbytes = int(torch.tensor([], dtype=dt).element_size())
        for d in shape:
            nbytes *= int(d)
        layout[name] = (off, nbytes, np_dt[dt], tuple(int(d) for d in shape))
        off = (off + nbytes + 255) & ~255
    flat = torch.zeros([max(off, 256)], dtype=torch.uint8, device=device)
    out._flat, out._layout = flat, layout
    for name, dt, shp in _PATH_OUTPUTS:
        if name not in layout:
            setattr(out, name, None)
            continue
        o, nbytes, _, shape = layout[name]
        t = flat[o:o + nbytes].view(dt).reshape(shape)
        setattr(out, name, t)
        setattr(p, name, t.data_ptr())
    check(lib().ppnet_path_synthesize(ctypes.byref(p), _stream()), "ppnet_path_synthesize")
    out.n_paths, out.path0 = n_paths, path0
    return out


PATH_STATUS_BITS = {1: "max_obst_iter reached in set_obstacles", 2: "supplied torch.rand draws exhausted",
                    4: "path-obstacle capacity (pomax) overflow", 8: "hull / isle capacity (hmax) overflow"}


def path_synthesize_checked(path0, n_paths, hmax=64, pomax=32, max_obst_iter=256, **kw):
    """path_synthesize, then the guards the kernels report instead of acting on (`status`, `hull_cnt`): on a capacity
    overflow or an exhausted iteration guard the launch is repeated with larger capacities (draws are keyed by the path
    id, so the repeat reproduces every path that was fine); what still fails raises.  The reference has no capacities --
    its lists grow and its set_obstacles loops until it succeeds (EDaGe-PP/Path.py:463-500)."""
    while True:
        out = path_synthesize(path0, n_paths, hmax=hmax, pomax=pomax, max_obst_iter=max_obst_iter, **kw)
        st = bits = 0
        if n_paths:                                          # one device read for all the guards
            s_ = out.status
            st1, st2, st4, st8, hmx = torch.stack([(s_ & 1).max(), (s_ & 2).max(), (s_ & 4).max(), (s_ & 8).max(),
                                                   out.hull_cnt.max()]).tolist()
            bits = (1 if st1 else 0) | (2 if st2 else 0) | (4 if st4 else 0) | (8 if st8 or hmx > hmax else 0)
            st = bits
        if bits == 0:
            return out
        if bits & 2:
            raise PPNetError("path_synthesize: %s" % PATH_STATUS_BITS[2])
        grown = False
        if bits & 8 and hmax < 128:
            hmax, grown = 128, True
        if bits & 4 and pomax < 512:
            pomax, grown = pomax * 2, True
        if bits & 1 and max_obst_iter < (1 << 16):
            max_obst_iter, grown = max_obst_iter * 8, True
        if not grown:
            raise PPNetError("path_synthesize: " + "; ".join(v for k, v in PATH_STATUS_BITS.items() if bits & k) +
                             " (status %d) at the largest supported capacities" % st)


def bits_to_image(bits, resolution, add=None):
    """A15's return value: bits i32[M,R,W] -> image f32[M,3,R,R] (1 free / 0 obstacle), optionally + `add`."""
    _need(bits, torch.int32, "bits")
    m = bits.shape[0]
    if add is not None:
        _need(add, torch.float32, "add")
    img = torch.empty([m, 3, resolution, resolution], dtype=torch.float32, device=bits.device)
    check(lib().ppnet_bits_to_image(_ptr(bits), ctypes.c_int32(resolution), ctypes.c_int64(m), _ptr(add), _ptr(img),
                                    _stream()), "ppnet_bits_to_image")
    return img


def add_init_end(image, init, end):
    """A16 batched (EDaGe-PP/process_map.py:119-145): image f32[M,3,R,R] in place, init/end f64[M,2] (row, col)."""
    _need(image, torch.float32, "image")
    _need(init, torch.float64, "init")
    _need(end, torch.float64, "end")
    m, _, r, _ = image.shape
    check(lib().ppnet_add_init_end(_ptr(image), ctypes.c_int32(r), _ptr(init), _ptr(end), ctypes.c_int64(m), _stream()),
          "ppnet_add_init_end")
    return image


def mask_rigid(src, angle_deg, translate, out_size):
    """torchvision's RandomRotation(degrees=(d, d)) + functional.affine(translate) + top-left crop on uint8 masks
    (EDaGe-PP/Path.py:160-178, MapGenerate.py:102-106): src u8[n,Ws,Ws], angle_deg f64[n], translate f64[n,2] (tx, ty)
    -> u8[n,out_size,out_size]."""
    _need(src, torch.uint8, "src")
    _need(angle_deg, torch.float64, "angle_deg")
    _need(translate, torch.float64, "translate")
    n, ws, _ = src.shape
    out = torch.empty([n, out_size, out_size], dtype=torch.uint8, device=src.device)
    check(lib().ppnet_mask_rigid(_ptr(src), ctypes.c_int32(ws), _ptr(angle_deg), _ptr(translate), ctypes.c_int64(n),
                                 ctypes.c_int32(out_size), _ptr(out), _stream()), "ppnet_mask_rigid")
    return out


def path_mask(pathpt, resolution=224, stride=5):
    """N2 generate_gen_path batched (EDaGe-PP/process_map.py:148-163): pathpt f64[M,Np,2] -> u8[M,R,R]."""
    _need(pathpt, torch.float64, "pathpt")
    m, np_, _ = pathpt.shape
    out = torch.empty([m, resolution, resolution], dtype=torch.uint8, device=pathpt.device)
    check(lib().ppnet_path_mask(_ptr(pathpt), ctypes.c_int32(np_), ctypes.c_int64(m), ctypes.c_int32(stride),
                                ctypes.c_int32(resolution), _ptr(out), _stream()), "ppnet_path_mask")
    return out


def extract_path(mask, init_state, end_state, down_sample_rate, max_len=4096):
    """N3 extract_path batched (EDaGe-PP/process_map.py:293-365): mask f32[n,h,w] (already down-sampled), init/end
    f64[n,2] -> (path f64[n,max_len+2,2], length i32[n], ok u8[n])."""
    _need(mask, torch.float32, "mask")
    _need(init_state, torch.float64, "init_state")
    _need(end_state, torch.float64, "end_state")
    n, h, w = mask.shape
    out = torch.zeros([n, max_len + 2, 2], dtype=torch.float64, device=mask.device)
    ln = torch.empty([n], dtype=torch.int32, device=mask.device)
    ok = torch.empty([n], dtype=torch.uint8, device=mask.device)
    check(lib().ppnet_extract_path(_ptr(mask), ctypes.c_int32(h), ctypes.c_int32(w), _ptr(init_state), _ptr(end_state),
                                   ctypes.c_double(down_sample_rate), ctypes.c_int64(n), ctypes.c_int32(max_len), _ptr(out),
                                   _ptr(ln), _ptr(ok), _stream()), "ppnet_extract_path")
    return out, ln, ok


def planner_masks(wp, path_off, clearance=1 / 50 * 224, resolution=224, points_per_seg=100):
    """N4 (EDaGe-PP/gerated_by_planners.py:88-157): wp f64[total,2] (x, y), CSR path_off i64[n+1]
    -> (mask_space u8[n,R,R], mask_path u8[n,R,R]) in file orientation (row = y, col = x), 1 = painted."""
    _need(wp, torch.float64, "wp")
    _need(path_off, torch.int64, "path_off")
    n = path_off.numel() - 1
    longest = int((path_off[1:] - path_off[:-1]).max().item()) if n else 0
    sp = torch.empty([n, resolution, resolution], dtype=torch.uint8, device=wp.device)
    pm = torch.empty_like(sp)
    check(lib().ppnet_planner_masks(_ptr(wp), _ptr(path_off), ctypes.c_int64(n), ctypes.c_int64(longest),
                                    ctypes.c_double(clearance), ctypes.c_int32(resolution), ctypes.c_int32(points_per_seg),
                                    _ptr(sp), _ptr(pm), _stream()), "ppnet_planner_masks")
    return sp, pm


def compact_u8(flags, keep=0):
    """Ordered stream compaction: indices i (ascending) with flags[i] == keep -> (idx i64[n] (first `count` valid),
    count i64[1]) on the device, e.g. the free segments of a verdict array or the feasible paths."""
    _need(flags, torch.uint8, "flags")
    n = flags.numel()
    L = lib()
    L.ppnet_compact_workspace_elems.restype = ctypes.c_int64
    ws = torch.empty([int(L.ppnet_compact_workspace_elems(ctypes.c_int64(n)))], dtype=torch.int64, device=flags.device)
    idx = torch.empty([n], dtype=torch.int64, device=flags.device)
    cnt = torch.empty([1], dtype=torch.int64, device=flags.device)
    check(L.ppnet_compact_u8(_ptr(flags), ctypes.c_int64(n), ctypes.c_uint8(keep), _ptr(idx), _ptr(cnt), _ptr(ws), _stream()),
          "ppnet_compact_u8")
    return idx, cnt


def propose_segments(map0, n_maps, segs_per_map, resolution=224, sigma=15.0, seed=DEFAULT_SEED, out=None, device="cuda"):
    """Device-side segment source (ppnet_propose_segments): the config-2 candidate segments of global maps
    [map0, map0 + n_maps) -> f64[n_maps * segs_per_map, 4] (s_row, s_col, e_row, e_col), a pure function of
    (seed, global map index, k)."""
    n = n_maps * segs_per_map
    out = torch.empty([n, 4], dtype=torch.float64, device=device) if out is None else _need(out, torch.float64, "out")
    check(lib().ppnet_propose_segments(ctypes.c_uint64(seed), ctypes.c_uint64(map0), ctypes.c_int64(n_maps),
                                       ctypes.c_int64(segs_per_map), ctypes.c_double(resolution), ctypes.c_double(sigma),
                                       _ptr(out), _stream()), "ppnet_propose_segments")
    return out


def compact_u8_i32(flags, keep=1, idx_base=0):
    """One-CTA ordered compaction of a short flag array (ppnet_compact_u8_i32) -> (idx i32[n], count i64[1])."""
    _need(flags, torch.uint8, "flags")
    n = flags.numel()
    idx = torch.empty([n], dtype=torch.int32, device=flags.device)
    cnt = torch.empty([1], dtype=torch.int64, device=flags.device)
    check(lib().ppnet_compact_u8_i32(_ptr(flags), ctypes.c_int64(n), ctypes.c_uint8(keep), ctypes.c_int32(idx_base), _ptr(idx),
                                     _ptr(cnt), _stream()), "ppnet_compact_u8_i32")
    return idx, cnt


def compact_bits(a, b=None, c=None, n=None, out=None, idx_base=0):
    """Ordered compaction of bit-packed verdicts (ppnet_compact_bits): survivors = segments whose bit is clear in every
    given array -> (idx i32[n] (first `count` valid, ascending), count i64[1]) on the device."""
    _need(a, torch.int32, "a")
    for t, nm in ((b, "b"), (c, "c")):
        if t is not None:
            _need(t, torch.int32, nm)
            if t.numel() != a.numel():
                raise PPNetError("compact_bits: arrays must have the same number of words")
    n = a.numel() * 32 if n is None else int(n)
    if (n + 31) // 32 != a.numel():
        raise PPNetError("compact_bits: n does not match the number of words")
    L = lib()
    L.ppnet_compact_bits_workspace_elems.restype = ctypes.c_int64
    if out is None:
        out = (torch.empty([n], dtype=torch.int32, device=a.device), torch.empty([1], dtype=torch.int64, device=a.device),
               torch.empty([int(L.ppnet_compact_bits_workspace_elems(ctypes.c_int64(n)))], dtype=torch.int64, device=a.device))
    idx, cnt, ws = out
    check(L.ppnet_compact_bits(_ptr(a), _ptr(b), _ptr(c), ctypes.c_int64(n), ctypes.c_int32(idx_base), _ptr(idx), _ptr(cnt),
                               _ptr(ws), _stream()),
          "ppnet_compact_bits")
    return idx, cnt, ws


def digest(t, unit0, acc, rows=None, row_elems=0, salt=0):
    """*acc += 64-bit content digest of the per-unit (per-map) array `t` [n_units, ...] (ppnet_digest_u32): additive over any
    split of the global unit range, so N ranks digesting their shards and summing (mod 2^64) must reproduce the digest of one
    rank doing everything.  `acc` int64[1] device tensor; `rows` int32[n_units] limits unit u to its first rows[u] rows of
    `row_elems` elements."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda or not t.is_contiguous():
        raise PPNetError("digest: needs a contiguous CUDA tensor")
    _need(acc, torch.int64, "acc")
    n = t.shape[0]
    if n == 0:
        return acc
    bytes_per_unit = t.numel() // n * t.element_size()
    if bytes_per_unit % 4:
        raise PPNetError("digest: bytes per unit must be a multiple of 4")
    if rows is not None:
        _need(rows, torch.int32, "rows")
    check(lib().ppnet_digest_u32(_ptr(t), ctypes.c_int64(bytes_per_unit // 4), ctypes.c_int64(n), ctypes.c_uint64(unit0), _ptr(rows),
                                 ctypes.c_int32(row_elems * t.element_size() // 4), ctypes.c_uint64(salt), _ptr(acc), _stream()),
          "ppnet_digest_u32")
    return acc


def digest_maps(gen, acc):
    """Digest of everything one generate_maps launch produced for maps [gen.map0, gen.map0 + gen.n_maps): labels
    (angle, translation, SegPoint, PathPoint), the accepted obstacle sets (obs_cnt rows) and the bit-packed maps."""
    g0 = gen.map0
    digest(gen.angle, g0, acc, salt=1)
    digest(gen.trans, g0, acc, salt=2)
    if gen.segpt is not None:
        digest(gen.segpt, g0, acc, salt=3)
    if gen.pathpt is not None:
        digest(gen.pathpt, g0, acc, salt=4)
    digest(gen.obs, g0, acc, rows=gen.obs_cnt, row_elems=3, salt=5)
    digest(gen.obs_cnt, g0, acc, salt=6)
    if gen.bits is not None:
        digest(gen.bits, g0, acc, salt=7)
    digest(gen.valid.view(torch.uint8).to(torch.int32), g0, acc, salt=8)
    return acc


def write_problems_jsonl(path, index, init, end, length, obs, obs_cnt, append=True):
    """N1 dataset writer (ppnet_write_problems_jsonl, host code): one json.dumps-identical line per problem
    {"Index", "Init", "End", "Length", "Obstacles"} (EDaGe-PP/MapGenerate.py:144-149).  numpy arrays: index i64[n],
    init / end f64[n,2], length f64[n], obs f64[n,omax,3], obs_cnt i32[n].  -> bytes written."""
    import numpy as np

    def c(a, dt, name, size):
        a = np.asarray(a)
        if a.dtype != np.dtype(dt) or not a.flags.c_contiguous or a.size != size:
            raise PPNetError("%s must be a C-contiguous %s array with %d elements" % (name, np.dtype(dt), size))
        return a
    n = len(np.asarray(index))
    obs = np.asarray(obs)
    omax = obs.shape[1] if obs.ndim == 3 else 0
    index, init, end = c(index, np.int64, "index", n), c(init, np.float64, "init", 2 * n), c(end, np.float64, "end", 2 * n)
    length, obs, obs_cnt = c(length, np.float64, "length", n), c(obs, np.float64, "obs", 3 * omax * n), c(obs_cnt, np.int32, "obs_cnt", n)
    nb = ctypes.c_int64()
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    check(lib().ppnet_write_problems_jsonl(str(path).encode(), ctypes.c_int32(1 if append else 0), ctypes.c_int64(n), p(index), p(init),
                                           p(end), p(length), p(obs), p(obs_cnt), ctypes.c_int32(omax), ctypes.byref(nb)),
          "ppnet_write_problems_jsonl")
    return nb.value
