"""Property tests (hypothesis) that cross-check the two restatements of the verdicts -- the line-by-line Python oracle
and the plain-C oracle used for bulk parity -- on adversarial inputs: degenerate segments, centres on the segment,
huge / tiny / non-finite values.  Both are test infrastructure; the golden fixtures pin them to the reference."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import c_oracle
from oracle import ppnet_oracle as orc

coord = st.one_of(st.floats(-50, 300, allow_nan=False), st.sampled_from([0.0, 224.0, 1e-300, 1e12, -1e12, np.inf, -np.inf, np.nan]))
radius = st.one_of(st.floats(0, 40, allow_nan=False), st.sampled_from([0.0, 1e-12, 1e6, np.inf, np.nan, -3.0]))
circle = st.tuples(coord, coord, radius)
seg = st.tuples(coord, coord, coord, coord)


@settings(max_examples=300, deadline=None)
@given(seg, st.lists(circle, min_size=0, max_size=5), st.sampled_from([0.0, 4.48, 13.44]), st.sampled_from([0, 1]))
def test_f64_python_and_c_oracles_agree(sg, circles, clearance, dot_mode):
    pts = np.asarray([sg], dtype=np.float64)
    obs = np.zeros([1, max(len(circles), 1), 3])
    for i, c in enumerate(circles):
        obs[0, i] = c
    cnt = np.asarray([len(circles)], dtype=np.int32)
    got = bool(c_oracle.segcheck_f64(pts, np.zeros(1, np.int32), obs, cnt, clearance, dot_mode=dot_mode)[0])
    with np.errstate(all="ignore"):
        want = orc.segcheck_edage_f64(pts[0, :2], pts[0, 2:], [list(c) for c in circles], clearance, dot_mode=dot_mode)
    assert got == bool(want)


@settings(max_examples=300, deadline=None)
@given(seg, st.lists(circle, min_size=0, max_size=5), st.sampled_from([0.0, 4.48]))
def test_f32_python_and_c_oracles_agree(sg, circles, clearance):
    with np.errstate(all="ignore"):
        pts = np.asarray([sg], dtype=np.float64).astype(np.float32)
    obs = np.zeros([1, max(len(circles), 1), 3])
    for i, c in enumerate(circles):
        obs[0, i] = c
    cnt = np.asarray([len(circles)], dtype=np.int32)
    v, steer = c_oracle.segcheck_f32(pts, np.zeros(1, np.int32), obs, cnt, clearance)
    with np.errstate(all="ignore"):
        want = orc.segcheck_mpnet_f32(pts[0, :2], pts[0, 2:], [list(c) for c in circles], clearance)
        want_st = orc.steer_to(pts[0, :2], pts[0, 2:], [list(c) for c in circles], clearance)
    assert bool(v[0]) == bool(want) and int(steer[0]) == int(want_st)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(st.integers(-40, 40), st.integers(-40, 40)), min_size=1, max_size=60))
def test_hull_is_convex_ccw_and_contains_every_point(pts):
    p = np.asarray(pts, dtype=np.int64)
    h = np.asarray(orc.hull2d(p), dtype=np.int64).reshape(-1, 2)
    assert len(h) >= 1 and set(map(tuple, h.tolist())) <= set(map(tuple, p.tolist()))
    if len(h) >= 3:
        for i in range(len(h)):
            a, b, c = h[i], h[(i + 1) % len(h)], h[(i + 2) % len(h)]
            assert (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]) > 0      # strict left turns: CCW, no collinear vertex
            for q in p:                                                                       # every point on the inner side of every edge
                assert (b[0] - a[0]) * (q[1] - a[1]) - (b[1] - a[1]) * (q[0] - a[0]) >= 0


def test_a15_canvas_model_affine_and_size():
    """The canvas-model restatement of plot_obstacles: a disk in the middle of the map comes out as a disk of the same
    area (to the +-1 px band), shifted by less than a pixel; nothing outside the map paints anything."""
    from oracle import ppnet_oracle as orc
    R = 224
    bits = orc.raster_canvas_bits([[112.0, 112.0, 20.0]], (R, R), R)
    occ = ((bits[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(R, -1)[:, :R]
    ii, jj = np.nonzero(occ)
    assert abs(occ.sum() - np.pi * 400) < 2 * np.pi * 20                 # area within one perimeter's worth of pixels
    assert abs(ii.mean() + 0.5 - 112) < 1.0 and abs(jj.mean() + 0.5 - 112) < 1.0
    ref = orc.raster_circles_bits([[112.0, 112.0, 20.0]], R)
    disk = ((ref[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(R, -1)[:, :R]
    assert 0 < (occ != disk).sum() < 2 * np.pi * 20
    assert orc.raster_canvas_bits([[-60.0, 40.0, 10.0], [float("nan"), 1.0, 1.0]], (R, R), R).sum() == 0
