"""Pin the oracle (oracle/ppnet_oracle.py + oracle/oracle_c.c) against the golden fixtures that the
REAL reference produced (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import ppnet_oracle as orc


# ------------------------------------------------------------------ A11 / A12 segment checks
def _rc(segs_xy):
    """(x, y) golden segments -> the (row, col) order process_map receives."""
    return np.stack([segs_xy[:, 1], segs_xy[:, 0], segs_xy[:, 3], segs_xy[:, 2]], axis=1)


import pytest  # noqa: E402


@pytest.mark.parametrize("fixture", ["segcheck_f64", "segcheck_f64_unfused"])
def test_segcheck_f64_python_oracle_vs_reference(golden, fixture):
    """`segcheck_f64` was produced by the reference under OpenBLAS' SkylakeX ddot (fused), `segcheck_f64_unfused` by the
    same code under OPENBLAS_CORETYPE=HASWELL (un-fused): both ddot models are pinned by real reference outputs."""
    g = golden(fixture)
    segs, mp, obs, cnt = g["segs_xy"], g["seg_map"], g["obs"], g["obs_cnt"]
    mode = int(g["dot_mode"])
    rc = _rc(segs)
    idx = np.concatenate([np.arange(0, len(segs), 5), np.arange(len(segs) - 711, len(segs))])
    bad = 0
    for i in idx:
        m = mp[i]
        r = orc.segcheck_edage_f64(rc[i, 0:2], rc[i, 2:4], obs[m, :cnt[m]].tolist(),
                                   float(g["clearance"]), mode)
        bad += int(r != bool(g["verdict"][i]))
    assert bad == 0


@pytest.mark.parametrize("fixture", ["segcheck_f64", "segcheck_f64_unfused"])
def test_segcheck_f64_c_oracle_vs_reference(golden, fixture):
    g = golden(fixture)
    assert int(g["dot_mode"]) == (0 if fixture == "segcheck_f64" else 1)
    v = c_oracle.segcheck_f64(_rc(g["segs_xy"]), g["seg_map"], g["obs"], g["obs_cnt"],
                              float(g["clearance"]), dot_mode=int(g["dot_mode"]), threads=2)
    assert np.array_equal(v, g["verdict"])
    # the other ddot model must NOT reproduce the fixture (the sharp cases hinge on it)
    v2 = c_oracle.segcheck_f64(_rc(g["segs_xy"]), g["seg_map"], g["obs"], g["obs_cnt"],
                               float(g["clearance"]), dot_mode=1 - int(g["dot_mode"]))
    assert (v2 != g["verdict"]).sum() > 0


def test_segcheck_kat_table(golden):
    """SURVEY 8(a) known-answer table (both flavours)."""
    want64 = dict(edge_hit=1, edge_miss=0, vertex_hit_e=1, start_inside_only=0, beyond_end_proj=0,
                  on_line_center=1, oob_xneg=0, oob_y225=0, x230=1)
    want32 = dict(edge_hit=1, edge_miss=0, vertex_hit_e=1, start_inside_only=0, beyond_end_proj=0,
                  on_line_center=1, oob_xneg=1, oob_y225=1, x230=0)
    for name, want in (("segcheck_f64", want64), ("segcheck_f32", want32)):
        g = golden(name)
        nk = int(g["n_kat"])
        got = dict(zip([str(s) for s in g["kat_names"]], g["verdict"][-nk:]))
        for k, v in want.items():
            assert got[k] == v, (name, k)


def test_segcheck_f32_python_oracle_vs_reference(golden):
    g = golden("segcheck_f32")
    segs, mp, obs, cnt = g["segs_xy"], g["seg_map"], g["obs"], g["obs_cnt"]
    idx = np.concatenate([np.arange(0, len(segs), 5), np.arange(len(segs) - 711, len(segs))])
    for i in idx:
        m = mp[i]
        ol = obs[m, :cnt[m]].tolist()
        assert orc.segcheck_mpnet_f32(segs[i, 0:2], segs[i, 2:4], ol) == bool(g["verdict"][i]), i
        assert orc.steer_to(segs[i, 0:2], segs[i, 2:4], ol) == int(g["steer"][i]), i


def test_segcheck_f32_c_oracle_vs_reference(golden):
    g = golden("segcheck_f32")
    v, st = c_oracle.segcheck_f32(g["segs_xy"], g["seg_map"], g["obs"], g["obs_cnt"],
                                  float(g["clearance"]), threads=2)
    assert np.array_equal(v, g["verdict"])
    assert np.array_equal(st, g["steer"])


def test_mpnet_feasibility_and_lvc(golden):
    g = golden("segcheck_f32")
    feas, _ = c_oracle.feasible(g["path_pts"], g["path_off"], g["path_map"], g["obs"], g["obs_cnt"],
                                float(g["clearance"]))
    assert np.array_equal(feas, g["feasible"])
    out, out_len = c_oracle.lvc(g["path_pts"], g["path_off"], g["path_map"], g["obs"], g["obs_cnt"],
                                float(g["clearance"]), threads=2)
    po, lo = g["path_off"], g["lvc_off"]
    for p in range(len(g["path_map"])):
        want = g["lvc_pts"][lo[p]:lo[p + 1]]
        assert out_len[p] == len(want)
        assert np.array_equal(out[po[p]:po[p] + out_len[p]], want)
    # python oracle on a subset (slow)
    for p in range(0, len(g["path_map"]), 9):
        m = g["path_map"][p]
        ol = g["obs"][m, :g["obs_cnt"][m]].tolist()
        path = [g["path_pts"][k] for k in range(po[p], po[p + 1])]
        assert orc.feasibility_check(path, ol) == int(g["feasible"][p])
        got = orc.lvc(path, ol)
        assert np.array_equal(np.asarray(got), g["lvc_pts"][lo[p]:lo[p + 1]])


# ------------------------------------------------------------------ A4 / A5
def test_grid_index_vs_reference(golden):
    g = golden("grid")
    pts = g["pts"]
    for res, off, key in ((224, 224, "idx224"), (224, 112.0, "idx112"), (1024, 1024, "idx1024")):
        assert np.array_equal(orc.grid_index_vec(pts, 50, res, off), g[key])
        assert np.array_equal(c_oracle.grid_index(pts, 50, res, off), g[key])
    assert np.array_equal(orc.grid_index(pts[:300], 50, 224, 224), g["idx224"][:300])
    assert np.array_equal(orc.grid_index_vec(g["halves"], 1, 1, 0), g["idx_half"])
    # round-half-to-even spot values
    assert [orc.rint_half_even(v) for v in (0.5, 1.5, 2.5, 3.5, -0.5, -1.5)] == [0, 2, 2, 4, 0, -2]


def test_corridor_rays_vs_reference(golden):
    g = golden("grid")
    off = g["ray_off"]
    for i in range(len(g["ray_x0"])):
        want = {tuple(c) for c in g["ray_cells"][off[i]:off[i + 1]]}
        if i % 3 == 0:
            got = set(orc.corridor_ray(g["ray_x0"][i], g["ray_dir"][i], g["ray_step_num"][i], 50, 224,
                                       224, 448, 448))
            assert got == want, i
        sp, _ = c_oracle.corridor_paint(g["ray_x0"][i], g["ray_dir"][i], g["ray_step_num"][i], 50,
                                        224, 224, 448, 448)
        assert {tuple(c) for c in np.argwhere(sp)} == want, i


# ------------------------------------------------------------------ A1-A3, A5, A6 on whole paths
def _segs_from_golden(g, pre):
    segs = []
    for i in range(10):
        segs.append(orc.pathseg_from_draws(g[pre + "draw_y"][i], g[pre + "draw_end"][i], 4,
                                           bool(g[pre + "is_straight"][i])))
    return segs


def test_path_synthesis_vs_reference(golden):
    g = golden("paths")
    for k in range(int(g["n_paths"])):
        pre = "p%d_" % k
        # is_straight rule: np.random.random(1) < 0.2 (PathSeg.py:19)
        assert np.array_equal(g[pre + "draw_straight"] < 0.2, g[pre + "is_straight"])
        segs = _segs_from_golden(g, pre)
        np.testing.assert_allclose(np.asarray([s["Poly"] for s in segs]), g[pre + "Poly"],
                                   rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose([s["EndPoint"] for s in segs], g[pre + "EndPoint"], rtol=1e-12)
        np.testing.assert_allclose([s["Length"] for s in segs], g[pre + "SegLength"], rtol=1e-6)
        np.testing.assert_allclose([s["GradSt"] for s in segs], g[pre + "GradSt"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose([s["GradEnd"] for s in segs], g[pre + "GradEnd"], rtol=1e-6, atol=1e-9)
        ch = orc.path_chain(segs)
        np.testing.assert_allclose(ch["Rotation"], g[pre + "SegRotation"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(ch["Translation"], g[pre + "SegTranslation"], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(ch["SegPoint"], g[pre + "SegPoint_raw"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(ch["PathPoint"], g[pre + "PathPoint_raw"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(ch["Length"], g[pre + "Length"], rtol=1e-6)
        bd = orc.draw_boundary(segs, ch, float(g[pre + "clearance"]))
        np.testing.assert_allclose(bd["up"], g[pre + "up"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(bd["up_dir"], g[pre + "up_dir"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(bd["down"], g[pre + "down"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(bd["init"], g[pre + "init"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(bd["end"], g[pre + "end"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(bd["BoundaryPoint"], g[pre + "BoundaryPoint_raw"], rtol=1e-5, atol=1e-8)


def test_corridor_paint_whole_path_vs_reference(golden):
    """Bit-exact painted-cell set when fed the reference's own boundary points (A5 contract:
    'bit-exact given the same points')."""
    g = golden("paths")
    for k in range(int(g["n_paths"])):
        pre = "p%d_" % k
        c = float(g[pre + "clearance"])
        bnd = dict(init=g[pre + "init"], end=g[pre + "end"], up=g[pre + "up"], up_dir=g[pre + "up_dir"],
                   down=g[pre + "down"], down_dir=-1 * g[pre + "up_dir"])
        x0, dr = orc.corridor_rays(bnd, g[pre + "SegPoint_raw"][-1], 50, 224)
        step_len = 1 / 224 * 50
        sp, _ = c_oracle.corridor_paint(x0, dr, 0.8 * c / step_len, 50, 224, 224, 448, 448)
        assert np.array_equal(sp, g[pre + "space_raw"]), k
        if k == 0:
            sp2 = orc.corridor_paint(x0, dr, 0.8 * c / step_len, 50, 224)
            assert np.array_equal(sp2, g[pre + "space_raw"])


def test_hull_vs_reference(golden):
    g = golden("paths")
    for k in range(int(g["n_paths"])):
        pre = "p%d_" % k
        pts = orc.grid_index_vec(g[pre + "PathPoint_raw"], 50, 224, 224)
        h = orc.hull2d(pts)
        want = g[pre + "hull_raw"].astype(np.int64)
        assert {tuple(p) for p in h} == {tuple(p) for p in want}
        assert orc.hull_signed_area2(h) > 0 and orc.hull_signed_area2(want) > 0      # both CCW
        # same cyclic order
        hl, wl = [tuple(p) for p in h], [tuple(p) for p in want]
        j = wl.index(hl[0])
        assert hl == wl[j:] + wl[:j]


def test_hull_convehull_fixture(golden):
    """convehull.py:5-7 fixture: seed(0), rand(30,2) -> hull.vertices."""
    g = golden("misc")
    assert list(g["hull_vertices"]) == [26, 10, 9, 6, 13, 8, 17, 7, 21]
    from scipy.spatial import ConvexHull
    np.random.seed(0)
    pts = np.random.rand(30, 2)
    assert np.array_equal(pts, g["hull_pts"])
    assert list(ConvexHull(pts).vertices) == list(g["hull_vertices"])
    # the strict monotone chain returns the same vertex set on scaled integer copies
    ipts = np.rint(pts * 1e6).astype(np.int64)
    h = orc.hull2d(ipts)
    want = ipts[g["hull_vertices"]]
    assert {tuple(p) for p in h} == {tuple(p) for p in want}


def test_pathseg_seeded_smoke(golden):
    g = golden("misc")
    np.testing.assert_allclose(g["seg0_poly"], [0.00155343, -0.03078496, 0.20270637, -0.49542406, 0],
                               rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(g["seg0_end"], 0.07044587, rtol=1e-6)
    np.testing.assert_allclose(g["seg0_len"], 0.0781822, rtol=1e-6)
    np.random.seed(0)
    u_s = np.random.random(1)
    y = np.random.random(1000)
    u_e = np.random.random(1)[0]
    s = orc.pathseg_from_draws(y, u_e, 4, bool(u_s < 0.2))
    np.testing.assert_allclose(s["Poly"], g["seg0_poly"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(s["Length"], g["seg0_len"], rtol=1e-6)


# ------------------------------------------------------------------ A10 / A13 / A14 via MapGenerate
def test_boundary_check_vs_reference(golden):
    g = golden("mapgen")
    tot = bad = 0
    for gi in range(int(g["n_groups"])):
        pre = "g%d_" % gi
        off = g[pre + "bc_hull_off"]
        for i in range(len(g[pre + "bc_ok"])):
            hull = g[pre + "tp%d_hull" % g[pre + "bc_path"][i]]
            ok, out = orc.boundary_check(hull, g[pre + "bc_angle_arg"][i], g[pre + "bc_trans_arg"][i], 224)
            ok_c, out_c = c_oracle.boundary_check(hull, g[pre + "bc_angle_arg"][i],
                                                  g[pre + "bc_trans_arg"][i][0],
                                                  g[pre + "bc_trans_arg"][i][1], 224)
            want = g[pre + "bc_hull_out"][off[i]:off[i + 1]]
            np.testing.assert_allclose(out, want, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(out_c, want, rtol=1e-9, atol=1e-9)
            tot += 1
            bad += int(ok != bool(g[pre + "bc_ok"][i])) + int(ok_c != bool(g[pre + "bc_ok"][i]))
    assert tot > 30 and bad == 0


def test_placement_labels_vs_reference(golden):
    g = golden("mapgen")
    for gi in range(int(g["n_groups"])):
        pre = "g%d_" % gi
        P = int(g[pre + "P"])
        for i, idx in enumerate(g[pre + "map_index"]):
            j = (int(idx) % (P * P)) // P                  # index = i*P^2 + j*P + k  (MapGenerate.py:68)
            ang, tr = g[pre + "label_angle"][i], g[pre + "label_translation"][i]
            sp = orc.place_points(g[pre + "tp%d_SegPointImage" % j], ang, tr, 224)
            pp = orc.place_points(g[pre + "tp%d_PathPoint" % j], ang, tr, 224)
            np.testing.assert_allclose(sp, g[pre + "label_segpoint"][i], rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(pp, g[pre + "label_pathpoint"][i], rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(pp, g[pre + "map_pathpoint"][i], rtol=1e-5, atol=1e-9)
            po = g[pre + "map_pathobs_off"]
            ob = orc.place_obstacles(g[pre + "tp%d_obstacles" % j], ang, tr, 224)
            np.testing.assert_allclose(ob, g[pre + "map_pathobs"][po[i]:po[i + 1]], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(g[pre + "prob_init"][i], sp[0], rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(g[pre + "prob_end"][i], sp[10], rtol=1e-5, atol=1e-9)


def test_clearance_filter_vs_reference(golden):
    """A14: bit-exact accept set + emitted [col,row,r] triples, fed the reference's own path points
    and candidate draws."""
    g = golden("mapgen")
    for gi in range(int(g["n_groups"])):
        pre = "g%d_" % gi
        c = float(g[pre + "clearance"])
        pp, cand = g[pre + "map_pathpoint"], g[pre + "map_cand"]
        acc_c, out_c, cnt_c = c_oracle.clearance_filter(pp, cand, 50, 224, c, threads=2)
        off, pof = g[pre + "prob_obs_off"], g[pre + "map_pathobs_off"]
        for i in range(len(pp)):
            n_path = pof[i + 1] - pof[i]
            want = g[pre + "prob_obs"][off[i]:off[i + 1] - n_path]
            acc, out = orc.clearance_filter(pp[i], cand[i], 50, 224, c)
            assert np.array_equal(out, want), (gi, i)
            assert cnt_c[i] == len(want)
            assert np.array_equal(out_c[i, :cnt_c[i]], want)
            assert np.array_equal(acc_c[i].astype(bool), acc)


def test_clearance_filter_sharp():
    """Radii tuned to +-2 ulp of the decision threshold: python and C oracles must agree."""
    rng = np.random.default_rng(3)
    pp = rng.uniform(20, 200, (8, 1000, 2))
    cand = np.stack([rng.uniform(0, 50, (8, 50)), rng.uniform(0, 50, (8, 50)),
                     rng.uniform(0, 5, (8, 50))], axis=2)
    thr_c = 1.0 / 50 * 224
    for m in range(8):
        for j in range(50):
            q = cand[m, j, :2] / 50 * 224
            d = pp[m, 1::2] - q
            mind = np.sqrt(np.min(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]))
            r_img = mind - thr_c
            if r_img > 0 and j % 2:
                r = r_img / 224 * 50
                for _ in range(int(rng.integers(0, 3))):
                    r = np.nextafter(r, np.inf if rng.random() < 0.5 else -np.inf)
                cand[m, j, 2] = r
    acc_c, out_c, cnt_c = c_oracle.clearance_filter(pp, cand, 50, 224, 1.0)
    n_acc = 0
    for m in range(8):
        acc, out = orc.clearance_filter(pp[m], cand[m], 50, 224, 1.0)
        assert np.array_equal(acc, acc_c[m].astype(bool))
        assert np.array_equal(out, out_c[m, :cnt_c[m]])
        n_acc += acc.sum()
    assert 0 < n_acc < 8 * 50


# ------------------------------------------------------------------ A16
def test_add_init_end_single_vs_reference(golden):
    g = golden("misc")
    for i in range(len(g["img_in"])):
        img = g["img_in"][i].copy()
        out = orc.add_init_end_single(img, g["init"][i], g["end"][i])
        assert np.array_equal(out, g["img_out"][i])


# ------------------------------------------------------------------ raster + DDA (new functionality)
def test_raster_and_dda_python_vs_c():
    rng = np.random.default_rng(9)
    obs = np.zeros([6, 50, 3])
    cnt = rng.integers(0, 51, 6).astype(np.int32)
    obs[..., 0] = rng.uniform(-5, 229, (6, 50))
    obs[..., 1] = rng.uniform(-5, 229, (6, 50))
    obs[..., 2] = rng.uniform(0, 22.4, (6, 50))
    bits = c_oracle.raster_circles_bits(obs, cnt, 224, inflate=2.24)
    for m in range(6):
        assert np.array_equal(orc.raster_circles_bits(obs[m, :cnt[m]], 224, 2.24), bits[m])
    s = rng.uniform(-4, 228, (3000, 2))
    e = s + rng.normal(0, 30, (3000, 2))
    segs = np.concatenate([s, e], axis=1).astype(np.float32)
    segs[::50, 2:] = segs[::50, :2]                     # degenerate
    segs[::77] = np.floor(segs[::77]) + 0.5             # half-integer ties (rint half-even)
    sm = rng.integers(0, 6, 3000).astype(np.int32)
    v, fh = c_oracle.dda_gridcheck(bits, 224, segs, sm, threads=2)
    for i in range(0, 3000, 7):
        hit, k = orc.dda_gridcheck(bits[sm[i]], 224, segs[i, :2], segs[i, 2:])
        assert hit == bool(v[i]) and k == fh[i], i
    assert 0.2 < v.mean() < 0.99


def test_space_normalization_isles_and_path_obstacles_vs_reference(golden):
    """A7 (point part), A8 and A9 against the real Path.space_normalization / search_isle / set_obstacles
    (A9 replays the torch.rand(1) values the reference consumed)."""
    g = golden("paths")
    for k in range(int(g["n_paths"])):
        pre = "p%d_" % k
        c = int(g[pre + "clearance"])
        n = orc.space_normalization_points(g[pre + "SegPoint_raw"], g[pre + "PathPoint_raw"], g[pre + "BoundaryPoint_raw"],
                                           g[pre + "hull_raw"], 50, 224)
        assert abs(n["Rotation"] - float(g[pre + "Rotation"])) < 1e-12
        for name in ("Translation", "ConvexHull", "SegPointImage", "PathPoint", "BoundaryPoint"):
            np.testing.assert_allclose(n[name], g[pre + name], rtol=0, atol=1e-10, err_msg=name)
        isles = orc.search_isle(g[pre + "PathPoint"], g[pre + "ConvexHull"], c, 50, 224)
        off = g[pre + "isle_off"]
        assert len(isles) == len(off) - 1
        for (lo, hi), i in zip(isles, range(len(off) - 1)):
            assert np.array_equal(g[pre + "PathPoint"][lo:hi], g[pre + "isle_pts"][off[i]:off[i + 1]])
        obs, used = orc.set_obstacles(g[pre + "PathPoint"], isles, c, 50, 224, g[pre + "obst_rand"])
        assert used == len(g[pre + "obst_rand"])
        np.testing.assert_allclose(np.asarray(obs).reshape(-1, 3), g[pre + "obstacles"], rtol=1e-12, atol=0)
        # the invariant A9 guarantees: every emitted circle keeps clearance c_px to all odd path points
        odd = g[pre + "PathPoint"][1::2]
        for x, y, r in obs:
            d = np.sqrt((odd[:, 0] - y) ** 2 + (odd[:, 1] - x) ** 2).min()
            assert d >= r + c / 50 * 224 - 1e-4          # radius is float32 when it was not clamped


def test_next_rows_oracle_vs_reference(golden):
    """N1 corridor placement, N2 label masks, N3 extract_path against tests/golden/masks.npz (real reference)."""
    g = golden("masks")
    for i in range(len(g["n1_angle"])):
        placed = orc.mask_rigid(g["n1_space"][g["n1_path"][i]], -g["n1_angle"][i], g["n1_translation"][i], 224)
        assert np.array_equal(placed, g["n1_placed"][i])
        assert np.array_equal(orc.gen_path_mask(g["n2_pathpoint"][i]) != 0, g["n2_gen_path"][i] != 0)
        assert np.array_equal(placed > 127, g["n2_seg_space"][i] != 0)
    for k in range(int(g["n3_count"])):
        ok, path = orc.extract_path_walk(g["n3_%d_mask" % k], g["n3_%d_init" % k], g["n3_%d_end" % k], int(g["n3_ds"]))
        assert ok == bool(g["n3_%d_ok" % k])
        assert np.array_equal(path, g["n3_%d_path" % k])


def test_planner_masks_oracle_vs_reference(golden):
    g = golden("planner_masks")
    for i in range(int(g["n"])):
        sp, pm = orc.planner_masks(g["wp"][g["off"][i]:g["off"][i + 1]])
        assert np.array_equal(sp != 0, g["mask_space"][i] != 0) and np.array_equal(pm != 0, g["mask_path"][i] != 0)
