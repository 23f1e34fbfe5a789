"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the plain-C oracle (oracle/oracle_c.c).

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
`threads=N` fans contiguous ranges out over N host threads (ctypes releases the GIL)."""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "oracle_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_corridor_paint.restype = ctypes.c_long
        _lib.orc_boundary_check.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _fan(n, threads, fn):
    """Run fn(lo, hi) over [0, n) split into `threads` contiguous ranges."""
    threads = max(1, int(threads))
    if threads == 1 or n < 2 * threads:
        fn(0, n)
        return
    cuts = [n * t // threads for t in range(threads + 1)]
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda ab: fn(*ab), [(cuts[t], cuts[t + 1]) for t in range(threads)]))


def segcheck_f64(pts_rc, seg_map, obs, obs_cnt, clearance, bound=224.0, dot_mode=0, threads=1):
    pts = _c(pts_rc, np.float64).reshape(-1, 4)
    sm, ob, oc = _c(seg_map, np.int32), _c(obs, np.float64), _c(obs_cnt, np.int32)
    n, omax = len(pts), ob.shape[1]
    out = np.empty(n, dtype=np.uint8)
    L = lib()

    def run(lo, hi):
        L.orc_segcheck_f64(_p(pts[lo:hi]), _p(sm[lo:hi]), _p(ob), _p(oc), ctypes.c_int(omax),
                           ctypes.c_double(clearance), ctypes.c_double(bound), ctypes.c_int(dot_mode),
                           ctypes.c_long(hi - lo), _p(out[lo:hi]))

    _fan(n, threads, run)
    return out


def segcheck_f32(pts_xy, seg_map, obs, obs_cnt, clearance, bound=224.0, threads=1, want_steer=True):
    pts = _c(pts_xy, np.float32).reshape(-1, 4)
    sm, ob, oc = _c(seg_map, np.int32), _c(obs, np.float64), _c(obs_cnt, np.int32)
    n, omax = len(pts), ob.shape[1]
    out = np.empty(n, dtype=np.uint8)
    steer = np.empty(n, dtype=np.uint8) if want_steer else None
    L = lib()

    def run(lo, hi):
        L.orc_segcheck_f32(_p(pts[lo:hi]), _p(sm[lo:hi]), _p(ob), _p(oc), ctypes.c_int(omax),
                           ctypes.c_double(clearance), ctypes.c_double(bound), ctypes.c_long(hi - lo),
                           _p(out[lo:hi]), _p(steer[lo:hi]) if want_steer else None)

    _fan(n, threads, run)
    return (out, steer) if want_steer else out


def segcheck_f32_cmp(pts_xy, seg_map, obs, obs_cnt, clearance, cmp64, bound=224.0, threads=1):
    """A12 verdicts with the threshold comparison of either NumPy generation (cmp64 = 1: NumPy 1.x promotes the
    float32 offset to float64; 0: NEP 50, float32 compare)."""
    pts = _c(pts_xy, np.float32).reshape(-1, 4)
    sm, ob, oc = _c(seg_map, np.int32), _c(obs, np.float64), _c(obs_cnt, np.int32)
    n, omax = len(pts), ob.shape[1]
    out = np.empty(n, dtype=np.uint8)
    L = lib()

    def run(lo, hi):
        L.orc_segcheck_f32_cmp(_p(pts[lo:hi]), _p(sm[lo:hi]), _p(ob), _p(oc), ctypes.c_int(omax),
                               ctypes.c_double(clearance), ctypes.c_double(bound), ctypes.c_int(int(cmp64)),
                               ctypes.c_long(hi - lo), _p(out[lo:hi]))

    _fan(n, threads, run)
    return out


def feasible(wp, path_off, path_map, obs, obs_cnt, clearance, bound=224.0, threads=1):
    wp = _c(wp, np.float32).reshape(-1, 2)
    po, pm = _c(path_off, np.int64), _c(path_map, np.int32)
    ob, oc = _c(obs, np.float64), _c(obs_cnt, np.int32)
    n = len(pm)
    out = np.empty(n, dtype=np.uint8)
    chk = np.empty(n, dtype=np.int64)
    L = lib()

    def run(lo, hi):
        L.orc_feasible(_p(wp), _p(po[lo:hi + 1]), _p(pm[lo:hi]), _p(ob), _p(oc),
                       ctypes.c_int(ob.shape[1]), ctypes.c_double(clearance), ctypes.c_double(bound),
                       ctypes.c_long(hi - lo), _p(out[lo:hi]), _p(chk[lo:hi]))

    _fan(n, threads, run)
    return out, chk


def lvc(wp, path_off, path_map, obs, obs_cnt, clearance, bound=224.0, threads=1):
    wp = _c(wp, np.float32).reshape(-1, 2)
    po, pm = _c(path_off, np.int64), _c(path_map, np.int32)
    ob, oc = _c(obs, np.float64), _c(obs_cnt, np.int32)
    n = len(pm)
    out = np.zeros_like(wp)
    out_len = np.empty(n, dtype=np.int32)
    L = lib()

    def run(lo, hi):
        L.orc_lvc(_p(wp), _p(po[lo:hi + 1]), _p(pm[lo:hi]), _p(ob), _p(oc), ctypes.c_int(ob.shape[1]),
                  ctypes.c_double(clearance), ctypes.c_double(bound), ctypes.c_long(hi - lo), _p(out),
                  _p(out_len[lo:hi]))

    _fan(n, threads, run)
    return out, out_len


def clearance_filter(pathpt, cand, map_size, resolution, clearance, threads=1):
    pp = _c(pathpt, np.float64)
    cd = _c(cand, np.float64)
    n_maps, np_, _ = pp.shape
    O = cd.shape[1]
    acc = np.empty([n_maps, O], dtype=np.uint8)
    out = np.zeros([n_maps, O, 3], dtype=np.float64)
    cnt = np.empty(n_maps, dtype=np.int32)
    L = lib()

    def run(lo, hi):
        L.orc_clearance_filter(_p(pp[lo:hi]), ctypes.c_int(np_), _p(cd[lo:hi]), ctypes.c_int(O),
                               ctypes.c_double(map_size), ctypes.c_double(resolution),
                               ctypes.c_double(clearance), ctypes.c_long(hi - lo), _p(acc[lo:hi]),
                               _p(out[lo:hi]), _p(cnt[lo:hi]))

    _fan(n_maps, threads, run)
    return acc, out, cnt


def grid_index(pts, map_size, resolution, off):
    p = _c(pts, np.float64).reshape(-1, 2)
    out = np.empty(p.shape, dtype=np.int64)
    lib().orc_grid_index(_p(p), ctypes.c_long(len(p)), ctypes.c_double(map_size),
                         ctypes.c_double(resolution), ctypes.c_double(off), _p(out))
    return out


def corridor_paint(x0, dirs, step_num, map_size, resolution, off, W, H):
    x0, dirs = _c(x0, np.float64).reshape(-1, 2), _c(dirs, np.float64).reshape(-1, 2)
    sn = _c(np.broadcast_to(step_num, (len(x0),)), np.float64)
    space = np.zeros([W, H], dtype=np.uint8)
    n = lib().orc_corridor_paint(_p(x0), _p(dirs), _p(sn), ctypes.c_long(len(x0)),
                                 ctypes.c_double(map_size), ctypes.c_double(resolution),
                                 ctypes.c_double(off), ctypes.c_int(W), ctypes.c_int(H), _p(space))
    return space, n


def boundary_check(hull, angle_deg, t0, t1, resolution):
    h = _c(hull, np.float64).reshape(-1, 2)
    out = np.empty_like(h)
    ok = lib().orc_boundary_check(_p(h), ctypes.c_int(len(h)), ctypes.c_double(angle_deg),
                                  ctypes.c_double(t0), ctypes.c_double(t1),
                                  ctypes.c_double(resolution), _p(out))
    return bool(ok), out


def raster_circles_bits(obs, obs_cnt, resolution, inflate=0.0, threads=1):
    ob, oc = _c(obs, np.float64), _c(obs_cnt, np.int32)
    n_maps, omax, _ = ob.shape
    R = int(resolution)
    W = (R + 31) // 32
    bits = np.empty([n_maps, R, W], dtype=np.uint32)
    L = lib()

    def run(lo, hi):
        L.orc_raster_circles_bits(_p(ob[lo:hi]), _p(oc[lo:hi]), ctypes.c_int(omax),
                                  ctypes.c_long(hi - lo), ctypes.c_int(R), ctypes.c_double(inflate),
                                  _p(bits[lo:hi]))

    _fan(n_maps, threads, run)
    return bits


def dda_gridcheck(bits, resolution, segs_xy, seg_map, threads=1):
    b = _c(bits, np.uint32)
    sg, sm = _c(segs_xy, np.float32).reshape(-1, 4), _c(seg_map, np.int32)
    n = len(sg)
    v = np.empty(n, dtype=np.uint8)
    fh = np.empty(n, dtype=np.int32)
    L = lib()

    def run(lo, hi):
        L.orc_dda_gridcheck(_p(b), ctypes.c_int(int(resolution)), _p(sg[lo:hi]), _p(sm[lo:hi]),
                            ctypes.c_long(hi - lo), _p(v[lo:hi]), _p(fh[lo:hi]))

    _fan(n, threads, run)
    return v, fh
