"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of PPNet's EDaGe-PP hot path.

This file is the *checker*.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  The product (`ppnet_b200/`) never does and
fails loudly when its CUDA library is missing.

Every function restates one reference function op-for-op, each floating-point operation
individually rounded in the precision the reference uses (NumPy 2.3 / SciPy 1.18 / torch 2.11
semantics of the build container -- the reference pins none of them).  Citations are relative to
the PPNet tree (`/root/reference`).  The restatement is pinned by `tests/golden/*.npz`, which
`tests/golden/make_golden.py` produced by running the *real* reference code (through
`oracle/ref_loader.py`) in the build container; `tests/test_oracle_golden.py` replays them.

dot_mode (A11 only): NumPy's float64 `np.dot` / `np.linalg.norm` on 2-vectors go through
OpenBLAS `ddot`, which is `fma(a1, b1, rn(a0*b0))` with the SkylakeX kernel (AVX-512 hosts, the
default here) and the un-fused `rn(rn(a0*b0) + rn(a1*b1))` with the Haswell/Zen kernels.
`DOT_FUSED_SKX = 0`, `DOT_UNFUSED = 1`.
"""
import math
from fractions import Fraction

import numpy as np

DOT_FUSED_SKX = 0
DOT_UNFUSED = 1

f32 = np.float32
f64 = np.float64


# --------------------------------------------------------------------------------------------
# exact helpers
# --------------------------------------------------------------------------------------------
def fma64(a, b, c):
    """Correctly rounded a*b+c in binary64 (Python 3.12 has no math.fma)."""
    a, b, c = float(a), float(b), float(c)
    if not (math.isfinite(a) and math.isfinite(b) and math.isfinite(c)):
        return a * b + c
    r = Fraction(a) * Fraction(b) + Fraction(c)
    if r == 0:
        # sign of an exact zero sum: follow IEEE (round-to-nearest => +0 unless both -0)
        return a * b + c
    return float(r)


def dot2_f64(a0, a1, b0, b1, dot_mode):
    """np.dot of two float64 2-vectors (OpenBLAS ddot)."""
    if dot_mode == DOT_FUSED_SKX:
        return f64(fma64(a1, b1, f64(a0) * f64(b0)))
    return f64(f64(a0) * f64(b0) + f64(a1) * f64(b1))


def rint_half_even(v):
    """np.round on a scalar: round half to even (Path.py:383)."""
    return int(np.round(f64(v)))


# --------------------------------------------------------------------------------------------
# A4  Path.coord_euclidean2image  (EDaGe-PP/Path.py:378-386) -- the grid-index rule
# --------------------------------------------------------------------------------------------
def grid_index(points, map_size, resolution, mapoffset):
    """points f64[N,2] -> int64[N,2];  idx = int(np.round(v / step_len + mapoffset)),
    step_len = MapSize / Resolution in float64."""
    pts = np.reshape(np.asarray(points, dtype=np.float64), [-1, 2])
    step_len = f64(map_size) / f64(resolution)
    out = np.empty(pts.shape, dtype=np.int64)
    for i, p in enumerate(pts):
        out[i, 0] = int(np.round(p[0] / step_len + f64(mapoffset)))
        out[i, 1] = int(np.round(p[1] / step_len + f64(mapoffset)))
    return out


def grid_index_vec(points, map_size, resolution, mapoffset):
    """Vectorised form of grid_index (numpy elementwise ops round individually => same bits)."""
    pts = np.reshape(np.asarray(points, dtype=np.float64), [-1, 2])
    step_len = f64(map_size) / f64(resolution)
    return np.rint(pts / step_len + f64(mapoffset)).astype(np.int64)


# --------------------------------------------------------------------------------------------
# A5  Path.free_space_bydirection  (EDaGe-PP/Path.py:397-404) -- the float ray-march
# --------------------------------------------------------------------------------------------
def corridor_ray(x_init, direction, step_num, map_size, resolution, mapoffset, width, height):
    """Cells painted by one ray, in order.  v_i = x_init + i*dir (integer i times the f64
    vector, then one add -- not an accumulated +=); stop at the first cell that fails the
    strict test 0 < cx < width and 0 < cy < height."""
    x_init = np.asarray(x_init, dtype=np.float64).reshape(2)
    direction = np.asarray(direction, dtype=np.float64).reshape(2)
    cells = []
    for i in range(int(np.round(step_num))):
        v = x_init + i * direction
        c = grid_index(v, map_size, resolution, mapoffset)[0]
        if 0 < c[0] < width and 0 < c[1] < height:
            cells.append((int(c[0]), int(c[1])))
        else:
            break
    return cells


def corridor_paint(x_inits, directions, step_num, map_size, resolution):
    """Driver loops of Path.path_space (Path.py:117-134) given the 1100 ray origins/directions:
    returns the painted uint8 [2R,2R] (value 255)."""
    w = 2 * resolution
    space = np.zeros([w, w], dtype=np.uint8)
    for x0, d in zip(x_inits, directions):
        for cx, cy in corridor_ray(x0, d, step_num, map_size, resolution, resolution, w, w):
            space[cx, cy] = 255
    return space


# --------------------------------------------------------------------------------------------
# A11  process_map.collision_check_circle_edge  (EDaGe-PP/process_map.py:383-425), float64
# --------------------------------------------------------------------------------------------
def segcheck_edage_f64(s_rc, e_rc, obs, clearance, dot_mode=DOT_FUSED_SKX, bound=224.0):
    """s_rc, e_rc: (row, col) float64 pairs as the reference receives them; obs: iterable of
    [x, y, r] Python floats; returns bool."""
    s_r, s_c = f64(s_rc[0]), f64(s_rc[1])
    e_r, e_c = f64(e_rc[0]), f64(e_rc[1])
    # :384-387 bounds test on the raw inputs *before* the swap
    if s_r < 0 or s_c > bound:
        return True
    if e_r < 0 or e_c > bound:
        return True
    # :388-389 swap to (x, y)
    s0, s1 = s_c, s_r
    e0, e1 = e_c, e_r
    d0, d1 = e0 - s0, e1 - s1                                    # :390
    L = np.sqrt(dot2_f64(d0, d1, d0, d1, dot_mode))             # np.linalg.norm :391
    with np.errstate(all="ignore"):
        n0, n1 = d1 / L, (-d0) / L                              # :391
        for ox, oy, size in obs:
            o0, o1 = f64(f32(ox)), f64(f32(oy))                 # torch.tensor([ox,oy]) => f32 :396
            thr = f64(float(size) + float(clearance) / 2)       # :397
            # vertex test on e only, scipy euclidean (un-fused) :397
            v0, v1 = e0 - o0, e1 - o1
            if np.sqrt(v0 * v0 + v1 * v1) < thr:
                return True
            dis = dot2_f64(n0, n1, o0 - s0, o1 - s1, dot_mode)  # :406
            if dis > 0:                                         # :407-408
                n0, n1 = -n0, -n1
            a = abs(dis)                                        # :409
            p0, p1 = o0 + a * n0, o1 + a * n1                   # :410
            u0, u1 = p0 - s0, p1 - s1                           # :411
            nu = np.sqrt(dot2_f64(u0, u1, u0, u1, dot_mode))    # :412
            u0, u1 = u0 / nu, u1 / nu
            w0, w1 = p0 - e0, p1 - e1                           # :413
            nw = np.sqrt(dot2_f64(w0, w1, w0, w1, dot_mode))    # :414
            w0, w1 = w0 / nw, w1 / nw
            if a < thr and dot2_f64(u0, u1, w0, w1, dot_mode) < 0:   # :415
                return True
    return False


# --------------------------------------------------------------------------------------------
# A12  experiments/MPNet/neuralplanner.py:43-138 -- MPNet flavour, all float32
# --------------------------------------------------------------------------------------------
MPNET_CLEARANCE = 1 / 50 * 224      # neuralplanner.py:18


def segcheck_mpnet_f32(s, e, obs, clearance=MPNET_CLEARANCE, bound=224.0, cmp_mode=0):
    """s, e: (x, y) float32 pairs (no swap); obs: iterable of [x, y, r] Python floats.
    cmp_mode 0: NumPy >= 2 (NEP 50) -- the np.float32 offsets are compared with float32(size + clearance/2), what this
    container runs and the goldens record; cmp_mode 1: NumPy 1.x (the reference's requirements.txt era) -- a np.float32
    scalar against a Python float promotes to float64."""
    s0, s1 = f32(s[0]), f32(s[1])
    e0, e1 = f32(e[0]), f32(e[1])
    if s0 < 0 or s1 > bound:                                    # :44-47
        return True
    if e0 < 0 or e1 > bound:
        return True
    d0, d1 = f32(e0 - s0), f32(e1 - s1)                         # :50
    with np.errstate(all="ignore"):
        L = np.sqrt(f32(f32(d0 * d0) + f32(d1 * d1)))           # np.linalg.norm f32 :51
        n0, n1 = f32(d1 / L), f32(f32(-d0) / L)
        for ox, oy, size in obs:
            o0, o1 = f32(ox), f32(oy)                           # :53
            # Python-float threshold, cast to f32 for the compare (NumPy 2 / NEP 50) :54,66
            thr64 = float(size) + float(clearance) / 2
            thr = f32(thr64)
            lt = (lambda x: float(x) < thr64) if cmp_mode else (lambda x: x < thr)
            v0, v1 = f32(e0 - o0), f32(e1 - o1)
            if lt(np.sqrt(f32(f32(v0 * v0) + f32(v1 * v1)))):   # :54
                return True
            q0, q1 = f32(o0 - s0), f32(o1 - s1)
            dis = f32(f32(n0 * q0) + f32(n1 * q1))              # np.dot f32, un-fused :57
            if dis > 0:                                         # :58-59
                n0, n1 = f32(-n0), f32(-n1)
            a = f32(abs(dis))                                   # :60
            p0, p1 = f32(o0 + f32(a * n0)), f32(o1 + f32(a * n1))   # :61
            u0, u1 = f32(p0 - s0), f32(p1 - s1)                 # :62
            nu = np.sqrt(f32(f32(u0 * u0) + f32(u1 * u1)))      # :63
            u0, u1 = f32(u0 / nu), f32(u1 / nu)
            w0, w1 = f32(p0 - e0), f32(p1 - e1)                 # :64
            nw = np.sqrt(f32(f32(w0 * w0) + f32(w1 * w1)))      # :65
            w0, w1 = f32(w0 / nw), f32(w1 / nw)
            if lt(a) and f32(f32(u0 * w0) + f32(u1 * w1)) < 0:      # :66
                return True
    return False


def steer_to(start, end, obs, clearance=MPNET_CLEARANCE):
    """neuralplanner.py:86-92.  dist = scipy euclidean in f32 (un-fused)."""
    d0 = f32(f32(start[0]) - f32(end[0]))
    d1 = f32(f32(start[1]) - f32(end[1]))
    dist = np.sqrt(f32(f32(d0 * d0) + f32(d1 * d1)))
    if dist > 0:
        if segcheck_mpnet_f32(start, end, obs, clearance):
            return 0
    return 1


def feasibility_check(path, obs, clearance=MPNET_CLEARANCE):
    """neuralplanner.py:96-102."""
    for i in range(0, len(path) - 1):
        if steer_to(path[i], path[i + 1], obs, clearance) == 0:
            return 0
    return 1


def lvc(path, obs, clearance=MPNET_CLEARANCE):
    """Lazy vertex contraction, literally recursive as neuralplanner.py:123-138."""
    for i in range(0, len(path) - 1):
        for j in range(len(path) - 1, i + 1, -1):
            if steer_to(path[i], path[j], obs, clearance) == 1:
                pc = [path[k] for k in range(0, i + 1)] + [path[k] for k in range(j, len(path))]
                return lvc(pc, obs, clearance)
    return path


# --------------------------------------------------------------------------------------------
# A14  MapGenerate.generate_map_randomly  (EDaGe-PP/MapGenerate.py:126-151), float64
# --------------------------------------------------------------------------------------------
def clearance_filter(path_point, cand, map_size, resolution, clearance):
    """path_point f64[Np,2] (row, col); cand f64[O,3] = (x, y, r) in map units (the three
    np.random.random draws already scaled as MapGenerate.py:128-130).
    Returns (accept bool[O], accepted f64[K,3] rows [coord_img[1], coord_img[0], radius_img])."""
    pp = np.asarray(path_point, dtype=np.float64)
    odd = pp[1::2]                                              # `if i % 2` :139
    cand = np.asarray(cand, dtype=np.float64)
    accept = np.zeros(len(cand), dtype=bool)
    out = []
    M, R, c = f64(map_size), f64(resolution), f64(clearance)
    for k, item in enumerate(cand):
        q0 = item[0] / M * R                                    # :134
        q1 = item[1] / M * R
        radius_img = item[2] / M * R                            # :136
        dx = odd[:, 0] - q0                                     # scipy euclidean, un-fused :140
        dy = odd[:, 1] - q1
        m = np.sqrt(np.min(dx * dx + dy * dy))                  # sqrt monotone => == min of sqrt
        if m > radius_img + c / M * R:                          # :142
            accept[k] = True
            out.append([q1, q0, radius_img])                    # :143
    return accept, np.asarray(out, dtype=np.float64).reshape(-1, 3)


# --------------------------------------------------------------------------------------------
# rotation helper: Path.coord_rotation (Path.py:271-274) = 2x2 . 2xN np.dot (OpenBLAS dgemm)
# --------------------------------------------------------------------------------------------
def rot2(c, s, x0, x1, dot_mode=DOT_FUSED_SKX):
    """[[c,-s],[s,c]] . [x0,x1]; dgemm SkylakeX accumulates with FMA: fma(r01,x1, r00*x0)."""
    ms = -s
    if dot_mode == DOT_FUSED_SKX:
        return f64(fma64(ms, x1, f64(c) * f64(x0))), f64(fma64(c, x1, f64(s) * f64(x0)))
    return f64(f64(c) * f64(x0) + f64(ms) * f64(x1)), f64(f64(s) * f64(x0) + f64(c) * f64(x1))


# --------------------------------------------------------------------------------------------
# A10  Path.boundary_check  (EDaGe-PP/Path.py:100-111)
# --------------------------------------------------------------------------------------------
def boundary_check(hull, angle_deg, translation, resolution, dot_mode=DOT_FUSED_SKX):
    """hull f64[H,2] (row, col); `angle_deg` and `translation` exactly as passed to the
    reference method (MapGenerate passes -angle and [t1, t0]).
    h' = Rot(angle/180*pi).(h - R/2) + t + R/2 ; ok iff all 0 <= h' < R."""
    hull = np.asarray(hull, dtype=np.float64).reshape(-1, 2)
    R = f64(resolution)
    off = R / 2
    theta = f64(angle_deg) / 180 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    out = np.empty_like(hull)
    ok = True
    for i, h in enumerate(hull):
        r0, r1 = rot2(c, s, h[0] - off, h[1] - off, dot_mode)
        out[i, 0] = (r0 + f64(translation[0])) + off
        out[i, 1] = (r1 + f64(translation[1])) + off
    for h in out:                                              # :108-110
        if h[0] < 0 or h[0] >= R or h[1] < 0 or h[1] >= R:
            ok = False
            break
    return ok, out


# --------------------------------------------------------------------------------------------
# A13  MapGenerate.generate placement block  (EDaGe-PP/MapGenerate.py:63-93)
# --------------------------------------------------------------------------------------------
def place_translation(u2, resolution):
    """translation = np.array(np.random.random([2]) * R - R/2, dtype=int)  (:64) -- C cast,
    truncation toward zero."""
    R = f64(resolution)
    return [int(f64(u2[0]) * R - R / 2), int(f64(u2[1]) * R - R / 2)]


def place_angle(u1):
    """angle = np.random.random([1]) * 360 - 180  (:63)."""
    return f64(u1) * 360 - 180


def place_points(points, angle_deg, translation, resolution, dot_mode=DOT_FUSED_SKX):
    """Rigid transform of SegPointImage / PathPoint (MapGenerate.py:70-80):
    p' = (Rot(-angle/180*pi).(p - R/2) + R/2) + [t1, t0], translation = [t0, t1] as drawn."""
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 2)
    R = f64(resolution)
    off = R / 2
    theta = -f64(angle_deg) / 180 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    out = np.empty_like(pts)
    t_r, t_c = f64(translation[1]), f64(translation[0])
    for i, p in enumerate(pts):
        r0, r1 = rot2(c, s, p[0] - off, p[1] - off, dot_mode)
        out[i, 0] = (r0 + off) + t_r
        out[i, 1] = (r1 + off) + t_c
    return out


def place_obstacles(path_obs, angle_deg, translation, resolution, dot_mode=DOT_FUSED_SKX):
    """Path-hugging obstacles [x, y, r] -> placed [x', y', r] (MapGenerate.py:83-89)."""
    out = []
    R = f64(resolution)
    off = R / 2
    theta = -f64(angle_deg) / 180 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    for ob in path_obs:
        c0, c1 = f64(ob[1]) - off, f64(ob[0]) - off            # coord = [obs[1], obs[0]] - R/2
        r0, r1 = rot2(c, s, c0, c1, dot_mode)
        r0 = (r0 + off) + f64(translation[1])
        r1 = (r1 + off) + f64(translation[0])
        out.append([r1, r0, float(ob[2])])
    return np.asarray(out, dtype=np.float64).reshape(-1, 3)


# --------------------------------------------------------------------------------------------
# A6  Path.convexhull (EDaGe-PP/Path.py:388-395) -- Qhull on integer points == strict
#     monotone chain (vertex sets identical on 300/300 integer walks)
# --------------------------------------------------------------------------------------------
def hull2d(points):
    """points int[N,2] -> CCW hull vertices int[H,2] (strict corners only, no collinear points),
    starting from the lexicographically smallest point.  Degenerate (all collinear) inputs
    return the two extreme points (Qhull would raise; the reference never handles it)."""
    pts = sorted(set((int(p[0]), int(p[1])) for p in np.asarray(points).reshape(-1, 2)))
    if len(pts) <= 2:
        return np.asarray(pts, dtype=np.int64).reshape(-1, 2)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower = []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    upper = []
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return np.asarray(lower[:-1] + upper[:-1], dtype=np.int64).reshape(-1, 2)


def hull_signed_area2(h):
    h = np.asarray(h, dtype=np.int64)
    x, y = h[:, 0], h[:, 1]
    return int(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


# --------------------------------------------------------------------------------------------
# A1  PathSeg (EDaGe-PP/PathSeg.py:10-58) given the random draws
# --------------------------------------------------------------------------------------------
SEG_LEN_RANGE = 7       # PathSeg.py:5
MIN_LEN = 0             # PathSeg.py:6


def polyval(p, x):
    """np.polyval: Horner, y = y*x + p[i] starting from zeros (un-fused)."""
    y = np.zeros_like(np.asarray(x, dtype=np.float64))
    for pv in p:
        y = y * x + pv
    return y


def polyder(p):
    n = len(p) - 1
    return np.asarray([p[i] * (n - i) for i in range(n)], dtype=np.float64)


def pathseg_from_draws(y_noise, u_end, polyorder=4, is_straight=False):
    """PathSeg.random (:21-36) given y_noise = np.random.random(1000) and u_end =
    np.random.random(1).  Returns dict(Poly, EndPoint, Translation, GradSt, GradEnd, Length)."""
    x = np.arange(0, 1000) / 100
    y = np.asarray(y_noise, dtype=np.float64) * 10 - 5           # :23
    poly = np.polyfit(x, y, polyorder)                           # :24 (LAPACK lstsq; 1e-5 parity)
    return pathseg_from_poly(poly, f64(u_end) * (SEG_LEN_RANGE - MIN_LEN) + MIN_LEN, is_straight)


def pathseg_from_poly(poly, endpoint, is_straight=False):
    poly = np.array(poly, dtype=np.float64)
    poly[-1] = 0                                                 # :28
    if is_straight:                                              # :29-31
        poly[:len(poly) - 2] = 0
    end = f64(endpoint)
    y_end = polyval(poly, end)                                   # :39
    pd = polyder(poly)                                           # :44
    grad_st = polyval(pd, f64(0))
    grad_end = polyval(pd, end)
    xs = np.arange(0, 100) / 100 * (end - 0)                     # :50
    ys = polyval(poly, xs)
    length = f64(0)
    for i in range(99):                                          # :52-53 scipy euclidean
        dx, dy = xs[i + 1] - xs[i], ys[i + 1] - ys[i]
        length = length + np.sqrt(dx * dx + dy * dy)
    dx, dy = end - xs[99], y_end - ys[99]                        # :54-57
    length = length + np.sqrt(dx * dx + dy * dy)
    return dict(Poly=poly, EndPoint=end, Translation=np.array([end, y_end]), GradSt=grad_st,
                GradEnd=grad_end, Length=length)


# --------------------------------------------------------------------------------------------
# A2  Path.generate / transform / plot (EDaGe-PP/Path.py:78-98, 224-233, 253-316)
# --------------------------------------------------------------------------------------------
def path_chain(segs):
    """segs: list of pathseg dicts.  Returns dict(Rotation[S], Translation[S,2], SegPoint[S+1,2],
    PathPoint[100*S,2], Length)."""
    S = len(segs)
    # angle(i+1, i) = atan(GradEnd_i) - atan(GradSt_{i+1})  (:276-280); angle_abs = running sum
    rot = np.zeros(S)
    acc = 0.0
    for i in range(1, S):
        acc = acc + (math.atan(segs[i - 1]["GradEnd"]) - math.atan(segs[i]["GradSt"]))
        rot[i] = acc
    # translation_seg(index) = sum_{i<index} R(angle_abs(i)) . T_i   (i=0 un-rotated) (:291-299)
    trans = np.zeros([S, 2])
    t = np.zeros(2)
    for i in range(S):
        trans[i] = t
        Ti = segs[i]["Translation"]
        if i != 0:
            c, s = np.cos(rot[i]), np.sin(rot[i])
            t = t + np.array(rot2(c, s, Ti[0], Ti[1]))
        else:
            t = t + Ti
    seg_point = [[0.0, 0.0]]
    for i in range(S):                                           # :86-90
        px, py = segs[i]["EndPoint"], polyval(segs[i]["Poly"], segs[i]["EndPoint"])
        if i != 0:
            c, s = np.cos(rot[i]), np.sin(rot[i])
            px, py = rot2(c, s, px, py)
        seg_point.append([px + trans[i][0], py + trans[i][1]])
    pts = []
    for i in range(S):                                           # plot() :256-260
        xs = np.arange(0, 100) / 100 * segs[i]["EndPoint"]
        ys = polyval(segs[i]["Poly"], xs)
        if i != 0:
            c, s = np.cos(rot[i]), np.sin(rot[i])
            px = np.empty(100)
            py = np.empty(100)
            for k in range(100):
                px[k], py[k] = rot2(c, s, xs[k], ys[k])
            xs, ys = px, py
        pts.append(np.stack([xs + trans[i][0], ys + trans[i][1]], axis=1))
    path_point = np.concatenate(pts, axis=0)
    d = path_point[1:] - path_point[:-1]
    length = float(sum(np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])))   # :93-94
    return dict(Rotation=rot, Translation=trans, SegPoint=np.asarray(seg_point),
                PathPoint=path_point, Length=length)


# --------------------------------------------------------------------------------------------
# A3  Path.draw_boundary (EDaGe-PP/Path.py:318-356)
# --------------------------------------------------------------------------------------------
def draw_boundary(segs, chain, clearance):
    """Returns dict(up[S,50,2], up_dir[S,50,2], down, down_dir, init[50,2], end[50,2],
    BoundaryPoint[100*S+100, 2])."""
    S = len(segs)
    rot, trans = chain["Rotation"], chain["Translation"]
    end_point = chain["SegPoint"][-1]
    up, upd, down, downd = [], [], [], []
    for i in range(S):
        xs = np.arange(0, 50) / 50 * segs[i]["EndPoint"]          # :320
        ys = polyval(segs[i]["Poly"], xs)
        yd = polyval(polyder(segs[i]["Poly"]), xs)
        c, s = np.cos(rot[i]), np.sin(rot[i])
        pt = np.empty([50, 2])
        nrm = np.empty([50, 2])
        for k in range(50):
            if i != 0:
                px, py = rot2(c, s, xs[k], ys[k])
            else:
                px, py = xs[k], ys[k]
            pt[k] = [px + trans[i][0], py + trans[i][1]]
            # coord_rotation is applied for every segment, including 0 (Rotation_0 == False == 0)
            c0, s0 = (c, s) if i != 0 else (np.cos(0.0), np.sin(0.0))
            n0, n1 = rot2(c0, s0, yd[k], -1.0)
            nn = np.sqrt(dot2_f64(n0, n1, n0, n1, DOT_FUSED_SKX))   # np.linalg.norm :329
            nrm[k] = [n0 / nn, n1 / nn]
        up.append(pt - 0.5 * clearance * nrm)                     # :330
        upd.append(nrm.copy())
        down.append(pt + 0.5 * clearance * nrm)                   # :332
        downd.append(-1 * nrm)
    up, upd, down, downd = map(np.asarray, (up, upd, down, downd))
    init, end = [], []
    for i in range(50):                                           # :334-336
        a = np.pi / 50 * (i + 1)
        init.append(rot2(np.cos(a), np.sin(a), up[0][0][0], up[0][0][1]))
    last = up[S - 1][49] - end_point
    for i in range(50):                                           # :337-343
        a = -np.pi / 50 * (i + 1)
        r = rot2(np.cos(a), np.sin(a), last[0], last[1])
        end.append([r[0] + end_point[0], r[1] + end_point[1]])
    init, end = np.asarray(init, dtype=np.float64), np.asarray(end, dtype=np.float64)
    bp = np.concatenate([init[::-1], up.reshape(-1, 2), end, down.reshape(-1, 2)[::-1]], axis=0)
    return dict(up=up, up_dir=upd, down=down, down_dir=downd, init=init, end=end, BoundaryPoint=bp)


def corridor_rays(bnd, end_point, map_size, resolution):
    """Ray origins and directions in the order Path.path_space paints them (Path.py:118-134)."""
    step_len = 1 / resolution * map_size                          # :118
    x0, dr = [], []
    for p in bnd["init"]:                                         # :120-122
        nn = np.sqrt(dot2_f64(p[0], p[1], p[0], p[1], DOT_FUSED_SKX))
        x0.append(p)
        dr.append(-step_len * p / nn)
    ep = np.reshape(end_point, [2])
    for p in bnd["end"]:                                          # :123-126
        v = ep - p
        nn = np.sqrt(dot2_f64(v[0], v[1], v[0], v[1], DOT_FUSED_SKX))
        x0.append(p)
        dr.append(step_len * v / nn)
    for p, d in zip(bnd["up"].reshape(-1, 2), bnd["up_dir"].reshape(-1, 2)):      # :127-130
        x0.append(p)
        dr.append(step_len * d)
    for p, d in zip(bnd["down"].reshape(-1, 2), bnd["down_dir"].reshape(-1, 2)):  # :131-134
        x0.append(p)
        dr.append(step_len * d)
    return np.asarray(x0), np.asarray(dr)


# --------------------------------------------------------------------------------------------
# A7  Path.space_normalization, point part (EDaGe-PP/Path.py:157-193)
# --------------------------------------------------------------------------------------------
def space_normalization_points(seg_point, path_point, boundary_point, hull_raw, map_size, resolution):
    """Given the raw (map-unit) SegPoint/PathPoint/BoundaryPoint and the raw integer hull (cells at offset R):
    Rotation = atan(Ey/Ex)/pi*180 - 135 (:159; `atan`, not atan2); hull rotated about (R, R) by -Rotation and
    shifted so that its vertex mean sits at (R/2, R/2) (:162-176); every point set goes rotate -> A4 rounding
    (offset R) -> + the same shift (:180-188).  Returns dict(Rotation, Translation (as stored: swapped),
    ConvexHull, SegPointImage, PathPoint, BoundaryPoint)."""
    R = f64(resolution)
    seg_point = np.asarray(seg_point, dtype=np.float64)
    e = seg_point[-1]
    rotation = math.atan(e[1] / e[0]) / np.pi * 180 + (-135)
    rad = -rotation / 180 * np.pi
    c, s = np.cos(rad), np.sin(rad)
    hull = np.asarray(hull_raw, dtype=np.float64) - R
    hr = np.empty_like(hull)
    for i, h in enumerate(hull):
        hr[i] = rot2(c, s, h[0], h[1])
    hr = hr + R
    center = hr.mean(axis=0)
    shift = np.array([R / 2, R / 2]) - center                   # [translation[1], translation[0]] of the stored value

    def norm(pts):
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
        out = np.empty_like(pts)
        for i, q in enumerate(pts):
            out[i] = rot2(c, s, q[0], q[1])
        return grid_index_vec(out, map_size, resolution, R).astype(np.float64) + shift

    return dict(Rotation=rotation, Translation=np.array([shift[1], shift[0]]), ConvexHull=hr + shift,
                SegPointImage=norm(seg_point), PathPoint=norm(path_point), BoundaryPoint=norm(boundary_point))


# --------------------------------------------------------------------------------------------
# A8  Path.search_isle (EDaGe-PP/Path.py:502-537)
# --------------------------------------------------------------------------------------------
def search_isle(path_point, hull, clearance, map_size, resolution, width_coef=0.2):
    """-> list of (lo, hi): the isle is path_point[lo:hi].  For every hull edge longer than 5/step: nearest path
    point (first minimum) to each end; the slice between them is an isle iff the first point whose distance to
    the chord normal exceeds int(round(c/step*width_coef)) differs (coordinate-wise) from the slice's last point."""
    pp = np.asarray(path_point, dtype=np.float64)
    hull = np.asarray(hull, dtype=np.float64)
    step_len = 1 / f64(resolution) * f64(map_size)
    thr = int(np.round(f64(clearance) / step_len * width_coef))
    H = len(hull)
    out = []
    for i in range(H):
        j = 0 if i == H - 1 else i + 1
        dv = hull[j] - hull[i]
        if not np.sqrt(dot2_f64(dv[0], dv[1], dv[0], dv[1], DOT_FUSED_SKX)) > 5 / step_len:
            continue
        d0 = pp - hull[i]
        d1 = pp - hull[j]
        # np.linalg.norm = sqrt(ddot(x, x)) (fused on SkylakeX); compared by strict <: first minimum
        n0 = np.sqrt([dot2_f64(v[0], v[1], v[0], v[1], DOT_FUSED_SKX) for v in d0])
        n1 = np.sqrt([dot2_f64(v[0], v[1], v[0], v[1], DOT_FUSED_SKX) for v in d1])
        i0, i1 = int(np.argmin(n0)), int(np.argmin(n1))
        lo, hi = min(i0, i1), max(i0, i1)
        b = pp[lo:hi]
        if len(b) == 0:
            continue                                             # (the reference would raise IndexError)
        v = b[0] - b[-1]
        with np.errstate(invalid="ignore", divide="ignore"):
            v = v / np.sqrt(dot2_f64(v[0], v[1], v[0], v[1], DOT_FUSED_SKX))
        nrm = np.array([v[1], -v[0]])
        p = b[-1]
        for q in b:
            w = q - b[0]
            dis = abs(dot2_f64(w[0], w[1], nrm[0], nrm[1], DOT_FUSED_SKX))
            if dis > thr:                                        # NaN never breaks
                p = q
                break
        if (p != b[-1]).any():
            out.append((lo, hi))
    return out


# --------------------------------------------------------------------------------------------
# A9  Path.set_obstacles (EDaGe-PP/Path.py:463-500) given the torch.rand(1) draws it consumes
# --------------------------------------------------------------------------------------------
def set_obstacles(path_point, isles, clearance, map_size, resolution, draws, max_iter=100000):
    """path_point f64[Np,2] (normalised), isles = list of (lo, hi), draws = iterator of torch.rand(1) values
    (float32-valued).  Returns (obstacles [[x, y, r], ...], number of draws used).  float32 where the
    reference computes on float32 tensors (radius, motion, their sums), float64 elsewhere."""
    pp = np.asarray(path_point, dtype=np.float64)
    odd = pp[1::2]
    c_px = f64(clearance) / f64(map_size) * f64(resolution)
    size_clearance = c_px * 1.1
    draws = list(draws)
    used = 0

    def rand():
        nonlocal used
        v = f32(draws[used])
        used += 1
        return v

    obstacles = []
    for lo, hi in isles:
        isle = pp[lo:hi]
        center = (isle[0] + isle[-1]) / 2
        v = isle[0] - isle[-1]
        dt = v / np.sqrt(dot2_f64(v[0], v[1], v[0], v[1], DOT_FUSED_SKX))
        dn = np.array([dt[1], -dt[0]])
        w = isle[int(len(isle) / 2)] - center
        if not dot2_f64(w[0], w[1], dn[0], dn[1], DOT_FUSED_SKX) < 0:
            dn = -dn
        dis = []
        for q in isle:
            w = q - isle[0]
            dis.append(abs(dot2_f64(w[0], w[1], dn[0], dn[1], DOT_FUSED_SKX)))
        size_max = max(dis) * 2
        peak = isle[dis.index(max(dis))]
        obs_size = []                                            # float32 motions
        size_pre = f32(0)
        coord = None
        it = 0
        while (sum(obs_size, f32(0)) if obs_size else 0) < size_max:
            it += 1
            if it > max_iter:
                break
            first = len(obs_size) == 0
            radius = f32(f32(rand() * f32(size_max)) / f32(2))               # torch.rand(1) * size_max / 2
            random_normal = f32(1) if first else rand()
            motion = f32(random_normal * f32(f32(radius + f32(size_pre)) + (f32(size_clearance) if first else f32(0))))
            alt = f32(radius - (sum(obs_size, f32(0)) if obs_size else f32(0)))
            if alt > motion:                                     # Python max(motion, alt): alt only if strictly greater
                motion = alt
            base = peak if first else coord
            coord = base + f64(motion) * dn
            if not first:
                k = f32(f32(f32(f32(rand() - f32(0.5)) / f32(0.5)) * radius) / f32(2))
                coord = coord + f64(k) * dt
            dd = odd - coord
            m = float(np.sqrt(dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]).min())
            r_out = radius
            if m < f64(f32(radius + f32(c_px))):
                r_out = f64(m) - c_px
            if r_out > 0:
                obs_size.append(motion)
                size_pre = r_out
                obstacles.append([float(coord[1]), float(coord[0]), float(r_out)])
    return obstacles, used


# --------------------------------------------------------------------------------------------
# PathGroup.generate's per-path work end to end (PathGenerate.py:33-50), given every random draw
# --------------------------------------------------------------------------------------------
def synthesize_path(y_noise, u_end, straight, clearance, map_size, resolution, obst_draws, path_straight=False,
                    hull_raw=None, polyorder=4):
    """y_noise f64[S,1000], u_end f64[S], straight bool[S], obst_draws = torch.rand(1) values for set_obstacles.
    Returns a dict with every intermediate (names as in `ppnet_path_params`)."""
    S = len(u_end)
    segs = [pathseg_from_draws(y_noise[i], u_end[i], polyorder, bool(straight[i])) for i in range(S)]
    chain = path_chain(segs)
    bnd = draw_boundary(segs, chain, clearance)
    x0, dr = corridor_rays(bnd, chain["SegPoint"][-1], map_size, resolution)
    cells = grid_index_vec(chain["PathPoint"], map_size, resolution, resolution)
    hr = np.asarray(hull2d(cells)) if hull_raw is None else np.asarray(hull_raw)
    nrm = space_normalization_points(chain["SegPoint"], chain["PathPoint"], bnd["BoundaryPoint"], hr, map_size, resolution)
    isles = [] if path_straight else search_isle(nrm["PathPoint"], nrm["ConvexHull"], clearance, map_size, resolution)
    obs, used = set_obstacles(nrm["PathPoint"], isles, clearance, map_size, resolution, obst_draws, max_iter=256)
    step_len = 1 / f64(resolution) * f64(map_size)
    return dict(segs=segs, chain=chain, bnd=bnd, ray_x0=x0, ray_dir=dr, step_num=0.8 * f64(clearance) / step_len,
                cells=cells, hull_raw=hr, norm=nrm, isles=isles, obstacles=np.asarray(obs).reshape(-1, 3), used=used)


# --------------------------------------------------------------------------------------------
# A16  process_map.add_init_end_single (EDaGe-PP/process_map.py:119-145)
# --------------------------------------------------------------------------------------------
def add_init_end_single(image, init, end):
    """image f32[3,R,R] modified in place: 7x7 red (255,0,0) squares at round(init), round(end)."""
    res = image.shape[1]
    for pt in (init, end):
        r0, c0 = int(np.round(pt[0])), int(np.round(pt[1]))
        for j in range(-3, 4):
            for k in range(-3, 4):
                if 0 <= r0 + j < res and 0 <= c0 + k < res:
                    image[0, r0 + j, c0 + k] = 255
                    image[1, r0 + j, c0 + k] = 0
                    image[2, r0 + j, c0 + k] = 0
    return image


# --------------------------------------------------------------------------------------------
# N1 / A7 (mask): torchvision's RandomRotation(degrees=(d, d)) + functional.affine(translate) + crop on a mask
#   (Path.py:160-161, 175-178; MapGenerate.py:102-106; process_map.py:174-178)
# --------------------------------------------------------------------------------------------
def mask_rigid(src, angle_deg, translate, out_size):
    """src u8[Ws,Ws]; two nearest-neighbour passes: out(i, j) = rot(rint(i - ty), rint(j - tx)), rot(i2, j2) =
    src(rint(sin*x + cos*y + c), rint(cos*x - sin*y + c)), x = j2 - c, y = i2 - c, c = (Ws-1)/2, angle d."""
    src = np.asarray(src)
    ws = src.shape[0]
    ii, jj = np.meshgrid(np.arange(out_size), np.arange(out_size), indexing="ij")
    j2 = np.rint(jj - f64(translate[0]))
    i2 = np.rint(ii - f64(translate[1]))
    c = 0.5 * (ws - 1)
    th = f64(angle_deg) / 180.0 * np.pi
    x, y = j2 - c, i2 - c
    js = np.rint(np.cos(th) * x - np.sin(th) * y + c)
    is_ = np.rint(np.sin(th) * x + np.cos(th) * y + c)
    ok = (j2 >= 0) & (j2 < ws) & (i2 >= 0) & (i2 < ws) & (js >= 0) & (js < ws) & (is_ >= 0) & (is_ < ws)
    out = np.zeros([out_size, out_size], dtype=src.dtype)
    out[ok] = src[is_[ok].astype(int), js[ok].astype(int)]
    return out


# --------------------------------------------------------------------------------------------
# N2  process_map.generate_gen_path (EDaGe-PP/process_map.py:148-163): every 5th label point -> 255
# --------------------------------------------------------------------------------------------
def gen_path_mask(path_point, resolution=224):
    m = np.zeros([resolution, resolution], dtype=np.uint8)
    for step, q in enumerate(np.asarray(path_point, dtype=np.float64)):
        r0, c0 = int(np.round(q[0])), int(np.round(q[1]))
        if step % 5 == 0 and 0 < r0 < resolution and 0 < c0 < resolution:      # strict: row / col 0 are never painted
            m[r0, c0] = 255
    return m


# --------------------------------------------------------------------------------------------
# N3  process_map.extract_path (EDaGe-PP/process_map.py:293-365): greedy 8-neighbour walk on the down-sampled
#     heat-map (the PIL bilinear down-sampling itself is third-party and stays with the caller)
# --------------------------------------------------------------------------------------------
_MOTIONS = [[0, 1], [0, -1], [1, 0], [-1, 0], [1, 1], [1, -1], [-1, 1], [-1, -1]]


def extract_path_walk(mask_small, init_state, end_state, down_sample_rate, max_steps=100000):
    """mask_small f32[h,w]; returns (ok, path f64[L,2]) with path = [init_state, ds * walk..., end_state]."""
    mask = np.asarray(mask_small, dtype=np.float32)
    h, w = mask.shape
    ds = f64(down_sample_rate)
    init = np.asarray(init_state, dtype=np.float64) / ds
    end = np.asarray(end_state, dtype=np.float64) / ds
    path = []
    nxt = init
    for _ in range(max_steps):
        cand = [np.asarray(m, dtype=np.float64) + nxt for m in _MOTIONS]
        val = []
        for cpt in cand:
            r0, c0 = int(np.round(cpt[0])), int(np.round(cpt[1]))
            val.append(mask[r0, c0] if (0 <= r0 < h and 0 <= c0 < w) else np.float32(0))
        while max(val) > 0:
            ci = val.index(max(val))
            nxt = cand[ci]
            fresh = True
            for i, q in enumerate(path):
                d = np.sqrt((nxt[0] - q[0]) * (nxt[0] - q[0]) + (nxt[1] - q[1]) * (nxt[1] - q[1]))
                if (nxt[0] == q[0] and nxt[1] == q[1]) or (d <= 1.5 and i < len(path) - 2):
                    val[ci] = np.float32(0)
                    fresh = False
                    break
            if fresh:
                break
        if max(val) == 0:
            return False, np.zeros([0, 2])
        path.append(nxt)
        de = np.sqrt((nxt[0] - end[0]) * (nxt[0] - end[0]) + (nxt[1] - end[1]) * (nxt[1] - end[1]))
        if de <= 2.5:
            pts = [np.asarray(init_state, dtype=np.float64)] + [q * ds for q in path] + [np.asarray(end_state, dtype=np.float64)]
            return True, np.asarray(pts)
    return False, np.zeros([0, 2])


# --------------------------------------------------------------------------------------------
# N4  gerated_by_planners.generated_by_planners, the two label masks (EDaGe-PP/gerated_by_planners.py:88-157)
# --------------------------------------------------------------------------------------------
def planner_masks(waypoints, clearance=1 / 50 * 224, resolution=224, points_per_seg=100):
    """waypoints f64[L,2] (x, y) of one planner solution.  Returns (mask_space u8[R,R], mask_path u8[R,R]) in the
    orientation of the saved files (row = y, col = x), non-zero where the reference paints."""
    p = np.asarray(waypoints, dtype=np.float64)
    R = resolution
    step_img = 1 / 224                                             # hard-coded (:58)
    K = round(clearance / 2 / step_img)
    space = np.zeros([R, R], dtype=np.uint8)
    pathm = np.zeros([R, R], dtype=np.uint8)

    def paint(m, pts):
        x = np.rint(pts[:, 0])
        y = np.rint(pts[:, 1])
        ok = (x > 0) & (x < R - 1) & (y > 0) & (y < R - 1)          # 0 < x < 223 and 0 < y < 223
        m[y[ok].astype(int), x[ok].astype(int)] = 1

    ks = (np.arange(K) * step_img)[:, None]
    for a, b in ((p[0], p[-1]), (p[-1], p[0])):                    # 360 rays around each end point (:98-111)
        d = a - b
        d = d / np.sqrt(dot2_f64(d[0], d[1], d[0], d[1], DOT_FUSED_SKX))
        for l in range(360):
            rad = l / 180 * np.pi
            c, s_ = np.cos(rad), np.sin(rad)
            dl = np.array([c * d[0] + (-s_) * d[1], s_ * d[0] + c * d[1]])
            paint(space, a + ks * dl)
    wps = []
    for l in range(len(p) - 1):                                    # +-normal band along every edge (:114-134)
        d = p[l + 1] - p[l]
        nd = np.sqrt(dot2_f64(d[0], d[1], d[0], d[1], DOT_FUSED_SKX))
        d = d / nd
        step = np.sqrt(f64((p[l][0] - p[l + 1][0]) ** 2 + (p[l][1] - p[l + 1][1]) ** 2)) / points_per_seg
        nrm = np.array([d[1], -d[0]])
        for j in range(points_per_seg):
            w = p[l] + j * step * d
            wps.append(w)
            paint(space, w + ks * nrm)
            paint(space, w - ks * nrm)
    paint(pathm, np.asarray(wps).reshape(-1, 2))
    return space, pathm


# --------------------------------------------------------------------------------------------
# A15  plot_obstacles -- GEOMETRIC restatement (parity UNPINNED: matplotlib/Agg/JPEG/PIL dither
#      are not installed).  Pixel (row i, col j) is obstacle iff its centre
#      (j + 0.5, i + 0.5) lies inside a disk (x, y, r [+ inflate]).
# --------------------------------------------------------------------------------------------
def raster_circles_bits(obs, resolution, inflate=0.0):
    """obs [[x, y, r], ...] -> bit-packed occupancy uint32[R, ceil(R/32)] (bit j%32 of word j//32
    in row i set = obstacle).  Exact integer-free rule evaluated in float64:
    (j+0.5-x)^2 + (i+0.5-y)^2 <= (r+inflate)^2, each op rounded, un-fused."""
    R = int(resolution)
    W = (R + 31) // 32
    bits = np.zeros([R, W], dtype=np.uint32)
    jj = np.arange(R, dtype=np.float64) + 0.5
    for ox, oy, r in obs:
        rr = f64(r) + f64(inflate)
        if not rr > 0 or not np.isfinite(rr) or not np.isfinite(ox) or not np.isfinite(oy):
            continue
        r2 = rr * rr
        for i in range(max(0, int(math.floor(oy - rr - 1))), min(R, int(math.ceil(oy + rr + 1)))):
            dy = (f64(i) + 0.5) - f64(oy)
            dx = jj - f64(ox)
            hit = (dx * dx + dy * dy) <= r2
            for j in np.nonzero(hit)[0]:
                bits[i, j >> 5] |= np.uint32(1 << (j & 31))
    return bits


# A15, second mode: the CANVAS MODEL of plot_obstacles (EDaGe-PP/Path.py:36-49).  matplotlib's default figure is
# 6.4 x 4.8 in; savefig(dpi=90) -> 576 x 432 px; default axes [left 0.125, bottom 0.11, width 0.775, height 0.77] ->
# x in [72, 518.4], y (from the top, y axis inverted by ax.axis(ymin=size[1], ymax=0)) in [51.84, 384.48]; the reference
# then crops [:, 53:383, 73:517] and T.Resize(resolution)s the 330 x 444 crop (bilinear; torchvision 0.12 of the
# reference's requirements.txt does not antialias tensors).  Restated: canvas pixel black iff its centre is inside the
# ELLIPSE a data circle becomes under the two different axis scales; the real torch bilinear resize; occupied iff < 0.5.
# What is left out is what cannot be pinned here: Agg anti-aliasing, the JPEG round trip and PIL's dither.
def raster_canvas_bits(obs, size, resolution, inflate=0.0):
    import torch
    import torch.nn.functional as F
    R = int(resolution)
    W = (R + 31) // 32
    sx, sy = 446.4 / f64(size[0]), 332.64 / f64(size[1])
    canvas = np.ones([330, 444], dtype=np.float32)                       # white
    px = np.arange(73, 517, dtype=np.float64) + 0.5
    for ox, oy, r in obs:
        rr = f64(r) + f64(inflate)
        cx, cy = 72.0 + f64(ox) * sx, 51.84 + f64(oy) * sy
        if not rr > 0 or not np.isfinite(rr) or not np.isfinite(cx) or not np.isfinite(cy):
            continue
        ax, ay = rr * sx, rr * sy
        for v in range(330):
            ny = ((f64(53 + v) + 0.5) - cy) / ay
            nx = (px - cx) / ax
            canvas[v, (nx * nx + ny * ny) <= 1.0] = 0.0
    out = F.interpolate(torch.from_numpy(canvas)[None, None], size=(R, R), mode="bilinear", align_corners=False,
                        antialias=False)[0, 0].numpy()
    occ = out < 0.5
    bits = np.zeros([R, W], dtype=np.uint32)
    for i, j in zip(*np.nonzero(occ)):
        bits[i, j >> 5] |= np.uint32(1 << (j & 31))
    return bits


# --------------------------------------------------------------------------------------------
# DDA grid check (NEW functionality in the B200 build; no reference counterpart).
# Endpoints are snapped with the A4 rule (rint half-even, step 1, offset 0), then an all-integer
# DDA walks the major axis one cell at a time; minor = start + round_half_up(k*dminor/dmajor)
# computed in integers.  Verdict: first visited cell that is out of [0,R)^2 or occupied.
# --------------------------------------------------------------------------------------------
DDA_CLAMP = 1 << 29


def _snap(v):
    """A4 rule at step 1 / offset 0: rint half-to-even, clamped to +-2^29.  NaN -> None."""
    v = float(v)
    if v != v:
        return None
    return int(min(max(np.rint(f64(v)), -DDA_CLAMP), DDA_CLAMP))


def dda_cells(s_xy, e_xy):
    """Generator of the visited cells (x, y), k = 0 .. n."""
    x0, y0, x1, y1 = _snap(s_xy[0]), _snap(s_xy[1]), _snap(e_xy[0]), _snap(e_xy[1])
    dx, dy = x1 - x0, y1 - y0
    n = max(abs(dx), abs(dy))
    if n == 0:
        yield (x0, y0)
        return
    for k in range(n + 1):
        # round-half-up of k*d/n in integers: floor((2*k*d + n) / (2n))
        yield (x0 + (2 * k * dx + n) // (2 * n), y0 + (2 * k * dy + n) // (2 * n))


def dda_gridcheck(bits, resolution, s_xy, e_xy):
    """Returns (hit bool, first_hit_index int) -- index = k of the first blocked cell, -1 if free.
    A NaN coordinate is blocked at k = 0."""
    R = int(resolution)
    if any(float(v) != float(v) for v in (s_xy[0], s_xy[1], e_xy[0], e_xy[1])):
        return True, 0
    for k, (cx, cy) in enumerate(dda_cells(s_xy, e_xy)):
        if cx < 0 or cx >= R or cy < 0 or cy >= R:
            return True, k
        if (int(bits[cy, cx >> 5]) >> (cx & 31)) & 1:
            return True, k
    return False, -1


# --------------------------------------------------------------------------------------------
# A17  GMM (EDaGe-PP/GMM.py:7-16) -- analytic quantities for the statistical parity test
# --------------------------------------------------------------------------------------------
def gmm_marginal_cdf(x, weights, mean, std, dim):
    """CDF of marginal `dim` of the mixture at x (vectorised)."""
    from scipy.stats import norm
    w = np.asarray(weights, dtype=np.float64)
    w = w / w.sum()
    x = np.asarray(x, dtype=np.float64)[:, None]
    return (w[None, :] * norm.cdf(x, loc=np.asarray(mean)[None, :, dim],
                                   scale=np.asarray(std)[None, :, dim])).sum(axis=1)


# --------------------------------------------------------------------------------------------
# 64-bit content digest (ppnet_b200/csrc/digest.cu): cross-rank identity proof, additive over any split of the global
# unit range.  digest = sum over units u (global index g = unit0 + u) and 32-bit words k of
#   mix64(word ^ mix64(mix64(g * 0x9E3779B97F4A7C15 + salt) + k))   (mod 2^64), mix64 = splitmix64 finaliser.
# --------------------------------------------------------------------------------------------
def _mix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(30)); x = x * np.uint64(0xbf58476d1ce4e5b9)
        x = x ^ (x >> np.uint64(27)); x = x * np.uint64(0x94d049bb133111eb)
        x = x ^ (x >> np.uint64(31))
    return x


def digest_u32(data, unit0, rows=None, row_words=0, salt=0):
    """data: array [n_units, ...] of any dtype whose unit size is a multiple of 4 bytes -> Python int (mod 2^64)."""
    a = np.ascontiguousarray(data)
    n = a.shape[0]
    if n == 0:
        return 0
    w = a.reshape(n, -1).view(np.uint32).astype(np.uint64)
    wpu = w.shape[1]
    g = np.uint64(unit0) + np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = _mix64(g * np.uint64(0x9E3779B97F4A7C15) + np.uint64(salt))
        m = _mix64(w ^ _mix64(key[:, None] + np.arange(wpu, dtype=np.uint64)[None, :]))
    if rows is not None:
        keep = np.arange(wpu)[None, :] < (np.maximum(np.asarray(rows), 0)[:, None] * row_words)
        m = np.where(keep, m, np.uint64(0))
    return int(m.sum(dtype=np.uint64))
