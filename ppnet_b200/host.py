"""Host-buffer boundary: numpy arrays (pageable or pinned) in, numpy arrays out, through the `*_host` entry
points of the C ABI.  This is the call a user of the reference makes -- host data in, host data out; the
copies and the kernels are inside.  No CPU fallback."""
import ctypes

import numpy as np

from ._lib import PPNetError, check, lib
from .ops import DEFAULT_BOUND, DEFAULT_SEED, DOT_FUSED_SKX, GenParams


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def _c(a, dt, name):
    a = np.asarray(a)
    if a.dtype != dt or not a.flags.c_contiguous:
        raise PPNetError("%s must be a C-contiguous %s array" % (name, np.dtype(dt)))
    return a


def _seg_off(seg_off, n_segs, n_maps):
    """CSR offsets as the kernels index with them: int64[n_maps + 1], starting at 0, non-decreasing, ending at n_segs."""
    if seg_off is None:
        if n_maps == 0 or n_segs % max(n_maps, 1):
            raise PPNetError("uniform grouping needs n_segs divisible by n_maps (or pass seg_off)")
        return None
    so = _c(seg_off, np.int64, "seg_off")
    if so.size != n_maps + 1 or (so.size and (so[0] != 0 or so[-1] != n_segs or (np.diff(so) < 0).any())):
        raise PPNetError("seg_off must be int64[n_maps + 1], start at 0, be non-decreasing and end at n_segs (%d)" % n_segs)
    return so


def _out_arr(a, dt, size, name):
    """Caller-supplied output array or None -> a checked array (exact dtype / size / contiguity: it is handed to
    cudaMemcpyAsync as a raw pointer)."""
    if a is None:
        return np.empty(size, dtype=dt)
    if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dt) or not a.flags.c_contiguous or a.size != size or not a.flags.writeable:
        raise PPNetError("%s must be a writeable C-contiguous %s array with %d elements" % (name, np.dtype(dt), size))
    return a


class HostContext:
    """Two CUDA streams + a grow-only device arena (ppnet_ctx_create)."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        check(lib().ppnet_ctx_create(ctypes.c_int32(device), ctypes.byref(self._h)), "ppnet_ctx_create")
        self.device = device

    def close(self):
        if self._h:
            lib().ppnet_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bytes_moved(self):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        check(lib().ppnet_ctx_bytes(self._h, ctypes.byref(a), ctypes.byref(b)), "ppnet_ctx_bytes")
        return a.value, b.value

    # ---- A11
    def segcheck_edage_f64(self, pts_rc, obs, obs_cnt, clearance, seg_off=None, bound=DEFAULT_BOUND,
                           dot_mode=DOT_FUSED_SKX, out=None):
        pts = _c(pts_rc, np.float64, "pts_rc").reshape(-1, 4)
        ob, oc = _c(obs, np.float64, "obs"), _c(obs_cnt, np.int32, "obs_cnt")
        n, m = len(pts), len(oc)
        so = _seg_off(seg_off, n, m)
        out = _out_arr(out, np.uint8, n, "out")
        check(lib().ppnet_segcheck_edage_f64_host(self._h, _p(pts), ctypes.c_int64(n), _p(so),
                                                  ctypes.c_int64(0 if so is not None else n // max(m, 1)),
                                                  ctypes.c_int64(m), _p(ob), _p(oc), ctypes.c_int32(ob.shape[1]),
                                                  ctypes.c_double(clearance), ctypes.c_double(bound),
                                                  ctypes.c_int32(dot_mode), _p(out)), "ppnet_segcheck_edage_f64_host")
        return out

    # ---- A12
    def segcheck_mpnet_f32(self, pts_xy, obs, obs_cnt, clearance, seg_off=None, bound=DEFAULT_BOUND, out=None,
                           steer=None):
        pts = _c(pts_xy, np.float32, "pts_xy").reshape(-1, 4)
        ob, oc = _c(obs, np.float64, "obs"), _c(obs_cnt, np.int32, "obs_cnt")
        n, m = len(pts), len(oc)
        so = _seg_off(seg_off, n, m)
        out = _out_arr(out, np.uint8, n, "out")
        if steer is not None:
            steer = _out_arr(steer, np.uint8, n, "steer")
        check(lib().ppnet_segcheck_mpnet_f32_host(self._h, _p(pts), ctypes.c_int64(n), _p(so),
                                                  ctypes.c_int64(0 if so is not None else n // max(m, 1)),
                                                  ctypes.c_int64(m), _p(ob), _p(oc), ctypes.c_int32(ob.shape[1]),
                                                  ctypes.c_double(clearance), ctypes.c_double(bound), _p(out),
                                                  _p(steer)), "ppnet_segcheck_mpnet_f32_host")
        return out

    # ---- A14
    def clearance_filter_f64(self, pathpt, cand, map_size, resolution, clearance):
        pp, cd = _c(pathpt, np.float64, "pathpt"), _c(cand, np.float64, "cand")
        m, np_, _ = pp.shape
        O = cd.shape[1]
        acc = np.empty([m, O], dtype=np.uint8)
        out = np.zeros([m, O, 3], dtype=np.float64)
        cnt = np.empty(m, dtype=np.int32)
        check(lib().ppnet_clearance_filter_f64_host(self._h, _p(pp), ctypes.c_int32(np_), _p(cd), ctypes.c_int32(O),
                                                    ctypes.c_int64(m), ctypes.c_double(map_size),
                                                    ctypes.c_double(resolution), ctypes.c_double(clearance), _p(acc),
                                                    _p(out), _p(cnt)), "ppnet_clearance_filter_f64_host")
        return acc, out, cnt

    # ---- DDA
    def dda_gridcheck(self, bits, resolution, segs_xy, seg_off=None, out=None, first_hit=None):
        b = np.asarray(bits)
        if b.dtype not in (np.uint32, np.int32) or not b.flags.c_contiguous:
            raise PPNetError("bits must be a C-contiguous uint32 array")
        sg = _c(segs_xy, np.float32, "segs_xy").reshape(-1, 4)
        n, m = len(sg), b.shape[0]
        so = _seg_off(seg_off, n, m)
        out = _out_arr(out, np.uint8, n, "out")
        if first_hit is not None:
            first_hit = _out_arr(first_hit, np.int32, n, "first_hit")
        check(lib().ppnet_dda_gridcheck_host(self._h, _p(b), ctypes.c_int32(resolution), ctypes.c_int64(m), _p(sg),
                                             ctypes.c_int64(n), _p(so),
                                             ctypes.c_int64(0 if so is not None else n // max(m, 1)), _p(out),
                                             _p(first_hit)), "ppnet_dda_gridcheck_host")
        return out

    # ---- GMM
    def gmm_sample(self, seed, sample0, n, mean, std, weights, out=None):
        mean, std, w = _c(mean, np.float32, "mean"), _c(std, np.float32, "std"), _c(weights, np.float32, "weights")
        k, d = mean.shape
        out = _out_arr(out, np.float32, n * d, "out").reshape(n, d) if out is not None else np.empty([n, d], dtype=np.float32)
        check(lib().ppnet_gmm_sample_host(self._h, ctypes.c_uint64(seed), ctypes.c_uint64(sample0), ctypes.c_int64(n),
                                          ctypes.c_int32(k), ctypes.c_int32(d), _p(mean), _p(std), _p(w), _p(out)),
              "ppnet_gmm_sample_host")
        return out


class PipelineIO(ctypes.Structure):
    """Mirror of `ppnet_pipeline_io` (include/ppnet_b200.h)."""
    _fields_ = [("segs_rc_f64", ctypes.c_void_p), ("segs_xy_f32", ctypes.c_void_p), ("segs_per_map", ctypes.c_int64),
                ("clearance_px", ctypes.c_double), ("bound", ctypes.c_double), ("dot_mode", ctypes.c_int32),
                ("cmp_mode", ctypes.c_int32), ("verdict_f64", ctypes.c_void_p), ("verdict_f32", ctypes.c_void_p),
                ("verdict_dda", ctypes.c_void_p),
                ("vbits_f64", ctypes.c_void_p), ("vbits_f32", ctypes.c_void_p), ("vbits_dda", ctypes.c_void_p),
                ("free_idx", ctypes.c_void_p), ("free_count", ctypes.c_void_p), ("valid_idx", ctypes.c_void_p),
                ("valid_count", ctypes.c_void_p), ("propose_sigma", ctypes.c_double), ("out_segs_rc", ctypes.c_void_p)]


class HostBank:
    """Device copy of a target-path bank built from host arrays (ppnet_bank_upload)."""

    def __init__(self, pathpt, segpt, hull, hull_cnt, obs=None, obs_cnt=None, device=0):
        pp, sp = _c(pathpt, np.float64, "pathpt"), _c(segpt, np.float64, "segpt")
        hl, hc = _c(hull, np.float64, "hull"), _c(hull_cnt, np.int32, "hull_cnt")
        b = pp.shape[0]
        if obs is None:
            obs, obs_cnt = np.zeros([b, 1, 3]), np.zeros(b, dtype=np.int32)
        ob, oc = _c(obs, np.float64, "obs"), _c(obs_cnt, np.int32, "obs_cnt")
        self.n_bank, self.np, self.nseg1, self.hmax, self.pomax = b, pp.shape[1], sp.shape[1], hl.shape[1], ob.shape[1]
        self._h = ctypes.c_void_p()
        check(lib().ppnet_bank_upload(ctypes.c_int32(device), _p(pp), _p(sp), _p(hl), _p(hc), _p(ob), _p(oc),
                                      ctypes.c_int32(b), ctypes.c_int32(self.np), ctypes.c_int32(self.nseg1),
                                      ctypes.c_int32(self.hmax), ctypes.c_int32(self.pomax), ctypes.byref(self._h)),
              "ppnet_bank_upload")

    def close(self):
        if self._h:
            lib().ppnet_bank_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _out(a, dt, size, name):
    """Caller-supplied output array: exact dtype, size and contiguity (it is handed to cudaMemcpyAsync as a raw pointer)."""
    if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dt) or not a.flags.c_contiguous or a.size != size:
        raise PPNetError("%s must be a C-contiguous %s array with %d elements" % (name, np.dtype(dt), size))
    if not a.flags.writeable:
        raise PPNetError("%s must be writeable" % name)
    return a


def generate_maps_host(ctx, bank, map0, n_maps, reps, obstacles_num, out, resolution=224, map_size=50.0,
                       obstacle_size=5.0, clearance=1.0, seed=DEFAULT_SEED, max_tries=4096, raster_inflate=0.0, checks=None):
    """MapGenerate.generate into HOST arrays.  `out` is a dict of preallocated numpy arrays with any of the keys
    angle f64[n], trans i32[n,2], segpt f64[n,S+1,2], pathpt f64[n,Np,2], obs f64[n,O+pomax,3], obs_cnt i32[n],
    rand_cnt i32[n], bits u32[n,R,W], tries i32[n], valid u8[n], counters u64[4].
    `checks` (optional dict) fuses the verdicts on the fresh maps into the same call (ppnet_generate_and_check_host):
      segments    segs_rc_f64 f64[n*spm,4]; optionally segs_xy_f32 f32[n*spm,4] (round-1 two-array mode).  Without
                  segs_xy_f32 the float32 flavours run on the device-side cast + swap of segs_rc_f64 (one upload).
                  Or propose_sigma > 0 with segs_per_map: the segments are drawn on the device, nothing is uploaded
                  (out_segs_rc f64[n*spm,4] receives them if given).
      settings    clearance_px, bound, dot_mode, cmp_mode
      outputs     verdict_f64 / verdict_f32 / verdict_dda u8[n*spm]; vbits_f64 / vbits_f32 / vbits_dda
                  u32[ceil(n*spm/32)] bit-packed; free_idx i32[n*spm] + free_count i64[1] (segments free under every
                  requested verdict, ascending); valid_idx i32[n] + valid_count i64[1] (maps with a valid placement)."""
    p = GenParams()
    p.map0, p.n_maps, p.reps, p.obstacles_num, p.max_tries = map0, n_maps, reps, obstacles_num, max_tries
    p.resolution, p.map_size, p.obstacle_size = float(resolution), float(map_size), float(obstacle_size)
    p.clearance, p.raster_inflate, p.seed = float(clearance), float(raster_inflate), seed
    R, W = int(resolution), (int(resolution) + 31) // 32
    shapes = dict(angle=(np.float64, n_maps), trans=(np.int32, 2 * n_maps), segpt=(np.float64, 2 * bank.nseg1 * n_maps),
                  pathpt=(np.float64, 2 * bank.np * n_maps), obs=(np.float64, 3 * (obstacles_num + bank.pomax) * n_maps),
                  obs_cnt=(np.int32, n_maps), rand_cnt=(np.int32, n_maps), tries=(np.int32, n_maps), valid=(np.uint8, n_maps))
    for name in ("angle", "trans", "segpt", "pathpt", "obs", "obs_cnt", "rand_cnt", "bits", "tries", "valid"):
        a = out.get(name)
        if a is not None:
            if name == "bits":
                if a.dtype not in (np.uint32, np.int32):
                    raise PPNetError("bits must be a uint32 / int32 array")
                _out(a, a.dtype, n_maps * R * W, "bits")
            else:
                _out(a, shapes[name][0], shapes[name][1], name)
        setattr(p, "out_" + name, a.ctypes.data if a is not None else None)
    cts = out.get("counters")
    if cts is not None:
        _out(cts, np.uint64, 4, "counters")
    p.counters = cts.ctypes.data if cts is not None else None
    if checks is None:
        check(lib().ppnet_generate_maps_host(ctx._h, bank._h, ctypes.byref(p)), "ppnet_generate_maps_host")
        return out
    io = PipelineIO()
    s64, s32 = checks.get("segs_rc_f64"), checks.get("segs_xy_f32")
    if s64 is not None:
        s64 = _c(s64, np.float64, "segs_rc_f64")
        io.segs_rc_f64 = s64.ctypes.data
    if s32 is not None:
        s32 = _c(s32, np.float32, "segs_xy_f32")
        io.segs_xy_f32 = s32.ctypes.data
    if s64 is not None or s32 is not None:
        n_seg = s64.size // 4 if s64 is not None else s32.size // 4
        if s64 is not None and s32 is not None and s64.size != s32.size:
            raise PPNetError("segs_rc_f64 and segs_xy_f32 must describe the same segments")
        if n_maps == 0 or n_seg % max(n_maps, 1):
            raise PPNetError("segments must be uniformly grouped: len(segs) divisible by n_maps")
        io.segs_per_map = n_seg // n_maps
    else:
        io.propose_sigma = float(checks.get("propose_sigma", 0.0))
        io.segs_per_map = int(checks.get("segs_per_map", 0))
        if io.propose_sigma <= 0 or io.segs_per_map <= 0:
            raise PPNetError("checks need segments: segs_rc_f64 / segs_xy_f32, or propose_sigma and segs_per_map")
        n_seg = io.segs_per_map * n_maps
    io.clearance_px, io.bound, io.dot_mode = float(checks["clearance_px"]), float(checks.get("bound", DEFAULT_BOUND)), \
        int(checks.get("dot_mode", DOT_FUSED_SKX))
    io.cmp_mode = int(checks.get("cmp_mode", 0))
    n_words = (n_seg + 31) // 32
    for name, dt, size in (("verdict_f64", np.uint8, n_seg), ("verdict_f32", np.uint8, n_seg), ("verdict_dda", np.uint8, n_seg),
                           ("vbits_f64", np.uint32, n_words), ("vbits_f32", np.uint32, n_words), ("vbits_dda", np.uint32, n_words),
                           ("free_idx", np.int32, n_seg), ("free_count", np.int64, 1), ("valid_idx", np.int32, n_maps),
                           ("valid_count", np.int64, 1), ("out_segs_rc", np.float64, 4 * n_seg)):
        a = checks.get(name)
        if a is not None:
            setattr(io, name, _out(a, dt, size, name).ctypes.data)
    check(lib().ppnet_generate_and_check_host(ctx._h, bank._h, ctypes.byref(p), ctypes.byref(io)),
          "ppnet_generate_and_check_host")
    return out
